// K0 text ingest: the bytes of a FASTA file -> the 2-bit text, the wildcard mask and the record table, on the device.
//
// Replaces the reference's two serial passes over the file, countLength and readFile (countReads.cpp:28-125), and
// the packing loops of AutoTextArray (AutoTextArray.hpp:27-61).  Their state machine: '>' starts a header that runs
// to the next '\n' (a '>' anywhere, also in the middle of a sequence line or of a header); a '\n' that ends a header
// files the record (name = the bytes behind the line's LAST '>', start = bases kept so far); outside headers A C G T N
// are kept (N as code 0 + a wildcard bit) and every other byte -- lower case included -- is dropped.
//
// The header state is a prefix composition of three functions on one bit (identity, clear = '\n', set = '>'), so the
// file is cut into tiles of FA_TILE bytes and processed in three launches:
//   k_fa_summary  per tile: the composed function and the base / record counts as far as they do not depend on
//                 the state the tile is entered in, and the part that does (everything before the first '>' or '\n')
//   k_fa_scan     one CTA: exclusive scan of the summaries -> entry state, first base index, first record index per tile
//   k_fa_pack     per tile: classify again with the known state, rank the kept bases, stage the 2-bit codes and the
//                 wildcard bits in shared memory and write whole words (the two words a tile shares with its
//                 neighbours are merged with atomicOr into the zeroed arrays); record starts and the file offsets
//                 of the header ends go to the record arrays
// Inside a thread (16 bytes = one 128-bit load) the state is a 17-bit Kogge-Stone carry chain.
// Algorithmic bytes per file byte: 1 read + (2 + 1)/8 written per kept base; the input is read twice (summary, pack).
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace realgpu
{

static const uint32_t FA_THREADS = 256;
static const uint32_t FA_TILE = FA_THREADS * 16;               // bytes of one tile

// what a piece of the file does to the scan state; counts are split at its first '>' or '\n' (the first "setter"):
// in front of it the entry state decides, behind it the piece decides by itself
template<typename C>
struct FaSum
{
        uint32_t f;             // 0 = no setter (state passes through), 1 = leaves "no header", 2 = leaves "header"
        uint32_t pre_hdr;       // 1 = the first setter is a '\n': it files a record if the piece is entered inside a header
        C pre_cnt;              // A C G T N in front of the first setter: kept if the piece is entered outside a header
        C post_cnt;             // bases kept behind the first setter
        C post_hdr;             // records filed behind the first setter
};
typedef FaSum<uint32_t> FaSum32;
typedef FaSum<uint64_t> FaSum64;

template<typename A, typename B>
__device__ __forceinline__ A fa_combine(A const & a, B const & b)         // a, then b
{
        A r;
        if ( a.f == 0 )
        {
                r.f = b.f; r.pre_hdr = b.pre_hdr;
                r.pre_cnt = a.pre_cnt + b.pre_cnt; r.post_cnt = b.post_cnt; r.post_hdr = b.post_hdr;
        }
        else
        {
                r.f = b.f ? b.f : a.f; r.pre_hdr = a.pre_hdr;
                r.pre_cnt = a.pre_cnt;
                r.post_cnt = a.post_cnt + b.post_cnt + (a.f == 1 ? b.pre_cnt : 0);
                r.post_hdr = a.post_hdr + b.post_hdr + (a.f == 2 ? b.pre_hdr : 0);
        }
        return r;
}

// byte classes: bits 0-1 = 2-bit code, bit 2 = kept outside headers, bit 3 = N, bit 4 = '>', bit 5 = '\n'
__host__ __device__ __forceinline__ uint32_t fa_class(uint32_t c)
{
        return c == 'A' ? 4u : c == 'C' ? 5u : c == 'G' ? 6u : c == 'T' ? 7u : c == 'N' ? 12u : c == '>' ? 16u : c == '\n' ? 32u : 0u;
}

struct FaMasks { uint32_t base, nb, gt, nl, codes; };           // one bit (codes: two) per byte of the 16-byte piece, byte i = bit i

__device__ __forceinline__ FaMasks fa_classify(uint4 const v, const uint8_t * lut)
{
        uint32_t const w[4] = { v.x, v.y, v.z, v.w };
        FaMasks M; M.base = M.nb = M.gt = M.nl = M.codes = 0;
        #pragma unroll
        for ( int i = 0; i < 16; ++i )
        {
                uint32_t const t = lut[(w[i >> 2] >> (8 * (i & 3))) & 0xFF];
                M.base |= ((t >> 2) & 1) << i;
                M.nb |= ((t >> 3) & 1) << i;
                M.gt |= ((t >> 4) & 1) << i;
                M.nl |= ((t >> 5) & 1) << i;
                M.codes |= (t & 3) << (2 * i);
        }
        return M;
}

// bit 0 = the state the piece is entered in, bit i+1 = the state behind byte i (1 = inside a header)
__device__ __forceinline__ uint32_t fa_states(uint32_t gt, uint32_t nl, uint32_t in)
{
        uint32_t G = (gt << 1) | in;
        uint32_t P = ~((gt | nl) << 1);
        G |= (G << 1) & P; P &= P << 1;
        G |= (G << 2) & P; P &= P << 2;
        G |= (G << 4) & P; P &= P << 4;
        G |= (G << 8) & P; P &= P << 8;
        G |= (G << 16) & P;
        return G;
}

__device__ __forceinline__ FaSum32 fa_piece_summary(FaMasks const & M)
{
        FaSum32 s;
        uint32_t const set = M.gt | M.nl;
        if ( ! set )
        {
                s.f = 0; s.pre_hdr = 0; s.pre_cnt = __popc(M.base); s.post_cnt = 0; s.post_hdr = 0;
                return s;
        }
        uint32_t const first = set & (0u - set), front = first - 1, behind = ~(front | first);
        uint32_t const st = fa_states(M.gt, M.nl, 0);
        s.f = ((M.gt >> (31 - __clz(set))) & 1) ? 2 : 1;
        s.pre_hdr = (M.nl & first) ? 1 : 0;
        s.pre_cnt = __popc(M.base & front);
        s.post_cnt = __popc(M.base & ~st & behind);
        s.post_hdr = __popc(M.nl & st & behind);
        return s;
}

__device__ __forceinline__ FaSum32 fa_shfl_down(FaSum32 const & a, int o)
{
        FaSum32 r;
        r.f = __shfl_down_sync(0xffffffffu, a.f, o); r.pre_hdr = __shfl_down_sync(0xffffffffu, a.pre_hdr, o);
        r.pre_cnt = __shfl_down_sync(0xffffffffu, a.pre_cnt, o); r.post_cnt = __shfl_down_sync(0xffffffffu, a.post_cnt, o);
        r.post_hdr = __shfl_down_sync(0xffffffffu, a.post_hdr, o);
        return r;
}

// 16 bytes at file offset off (a multiple of 16; the buffer is 16-byte aligned); bytes behind the end of the file read as 0,
// a dropped byte that leaves the state alone
__device__ __forceinline__ uint4 fa_load(const uint8_t * bytes, uint64_t nbytes, uint64_t off)
{
        if ( off + 16 <= nbytes )
                return __ldg(reinterpret_cast<const uint4 *>(bytes + off));
        uint32_t w[4] = { 0, 0, 0, 0 };
        for ( int i = 0; i < 16; ++i )
                if ( off + i < nbytes ) w[i >> 2] |= (uint32_t)bytes[off + i] << (8 * (i & 3));
        return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(FA_THREADS) k_fa_summary(const uint8_t * __restrict__ bytes, uint64_t nbytes, FaSum32 * __restrict__ sums)
{
        __shared__ uint8_t lut[256];
        __shared__ FaSum32 wsum[FA_THREADS / 32];
        lut[threadIdx.x] = (uint8_t)fa_class(threadIdx.x);
        uint4 const v = fa_load(bytes, nbytes, (uint64_t)blockIdx.x * FA_TILE + threadIdx.x * 16);
        __syncthreads();
        FaSum32 s = fa_piece_summary(fa_classify(v, lut));
        #pragma unroll
        for ( int o = 1; o < 32; o <<= 1 )
        {
                FaSum32 const t = fa_shfl_down(s, o);           // lanes whose partner lies outside the warp combine garbage nobody reads
                s = fa_combine(s, t);
        }
        if ( (threadIdx.x & 31) == 0 ) wsum[threadIdx.x >> 5] = s;
        __syncthreads();
        if ( threadIdx.x == 0 )
        {
                FaSum32 a = wsum[0];
                #pragma unroll
                for ( int w = 1; w < (int)(FA_THREADS / 32); ++w ) a = fa_combine(a, wsum[w]);
                sums[blockIdx.x] = a;
        }
}

// one CTA; tile_base[t] = (index of the tile's first kept base << 1) | entry state, tile_rec[t] = index of its first record;
// totals[0] = kept bases, totals[1] = records of the whole file
static const uint32_t FA_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(FA_SCAN_THREADS) k_fa_scan(const FaSum32 * __restrict__ sums, uint64_t ntiles, uint64_t * __restrict__ tile_base,
                                                             uint64_t * __restrict__ tile_rec, uint64_t * __restrict__ totals)
{
        __shared__ FaSum64 wtot[FA_SCAN_THREADS / 32];
        uint64_t const per = (ntiles + FA_SCAN_THREADS - 1) / FA_SCAN_THREADS;
        uint64_t const t0 = min(ntiles, threadIdx.x * per), t1 = min(ntiles, t0 + per);
        FaSum64 acc; acc.f = 0; acc.pre_hdr = 0; acc.pre_cnt = 0; acc.post_cnt = 0; acc.post_hdr = 0;
        FaSum64 const ident = acc;
        for ( uint64_t t = t0; t < t1; ++t ) acc = fa_combine(acc, sums[t]);
        // inclusive scan over the threads of a warp, then over the warps
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        FaSum64 inc = acc;
        #pragma unroll
        for ( int o = 1; o < 32; o <<= 1 )
        {
                FaSum64 p;
                p.f = __shfl_up_sync(0xffffffffu, inc.f, o); p.pre_hdr = __shfl_up_sync(0xffffffffu, inc.pre_hdr, o);
                p.pre_cnt = __shfl_up_sync(0xffffffffu, inc.pre_cnt, o); p.post_cnt = __shfl_up_sync(0xffffffffu, inc.post_cnt, o);
                p.post_hdr = __shfl_up_sync(0xffffffffu, inc.post_hdr, o);
                if ( lane >= o ) inc = fa_combine(p, inc);
        }
        if ( lane == 31 ) wtot[wid] = inc;
        FaSum64 exc;                                            // everything in front of this thread inside its warp
        exc.f = __shfl_up_sync(0xffffffffu, inc.f, 1); exc.pre_hdr = __shfl_up_sync(0xffffffffu, inc.pre_hdr, 1);
        exc.pre_cnt = __shfl_up_sync(0xffffffffu, inc.pre_cnt, 1); exc.post_cnt = __shfl_up_sync(0xffffffffu, inc.post_cnt, 1);
        exc.post_hdr = __shfl_up_sync(0xffffffffu, inc.post_hdr, 1);
        if ( lane == 0 ) exc = ident;
        __syncthreads();
        FaSum64 pre = ident;
        for ( int w = 0; w < wid; ++w ) pre = fa_combine(pre, wtot[w]);
        pre = fa_combine(pre, exc);
        // the file is entered outside a header: the "pre" parts count as kept bases and file no record
        for ( uint64_t t = t0; t < t1; ++t )
        {
                tile_base[t] = ((pre.pre_cnt + pre.post_cnt) << 1) | (pre.f == 2 ? 1u : 0u);
                tile_rec[t] = pre.post_hdr;
                pre = fa_combine(pre, sums[t]);
        }
        if ( t1 == ntiles && t0 < t1 )
        {
                totals[0] = pre.pre_cnt + pre.post_cnt;
                totals[1] = pre.post_hdr;
        }
        if ( ntiles == 0 && threadIdx.x == 0 ) { totals[0] = 0; totals[1] = 0; }
}

// text / nmask: word 0 of the (zeroed) arrays; rec_start / rec_nl: one entry per record (rec_nl = file offset of the '\n' that filed it)
__global__ void __launch_bounds__(FA_THREADS) k_fa_pack(const uint8_t * __restrict__ bytes, uint64_t nbytes, const uint64_t * __restrict__ tile_base,
                                                        const uint64_t * __restrict__ tile_rec, unsigned long long * __restrict__ text,
                                                        unsigned long long * __restrict__ nmask, uint64_t * __restrict__ rec_start, uint64_t * __restrict__ rec_nl)
{
        __shared__ uint8_t lut[256];
        __shared__ uint32_t wf[FA_THREADS / 32];
        __shared__ uint32_t cu[FA_TILE / 16 + 4];               // 2-bit codes, 16 bases per unit, most significant first; unit pairs = text words
        __shared__ uint32_t nu[FA_TILE / 32 + 4];               // wildcard bits, 32 bases per unit; unit pairs = mask words
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        lut[threadIdx.x] = (uint8_t)fa_class(threadIdx.x);
        for ( uint32_t i = threadIdx.x; i < FA_TILE / 16 + 4; i += FA_THREADS ) cu[i] = 0;
        for ( uint32_t i = threadIdx.x; i < FA_TILE / 32 + 4; i += FA_THREADS ) nu[i] = 0;
        uint64_t const byte0 = (uint64_t)blockIdx.x * FA_TILE + threadIdx.x * 16;
        uint4 const v = fa_load(bytes, nbytes, byte0);
        uint64_t const tb = tile_base[blockIdx.x];
        uint64_t const B0 = tb >> 1;
        __syncthreads();
        FaMasks const M = fa_classify(v, lut);
        // entry state of this thread: the last setter in front of it -- in its warp, else in an earlier warp, else the tile's
        uint32_t const set = M.gt | M.nl;
        uint32_t const mine = set ? (((M.gt >> (31 - __clz(set))) & 1) ? 2u : 1u) : 0u;
        uint32_t const m_set = __ballot_sync(0xffffffffu, mine != 0), m_one = __ballot_sync(0xffffffffu, mine == 2);
        if ( lane == 0 ) wf[wid] = m_set ? (((m_one >> (31 - __clz(m_set))) & 1) ? 2u : 1u) : 0u;
        __syncthreads();
        uint32_t in = (uint32_t)(tb & 1);
        for ( int w = 0; w < wid; ++w ) if ( wf[w] ) in = wf[w] == 2;
        uint32_t const lower = m_set & ((1u << lane) - 1);
        if ( lower ) in = (m_one >> (31 - __clz(lower))) & 1;
        uint32_t const st = fa_states(M.gt, M.nl, in);           // bit i = state in front of byte i
        uint32_t const kept = M.base & ~st & 0xFFFFu, ends = M.nl & st & 0xFFFFu;
        uint32_t const k = __popc(kept), nh = __popc(ends);
        uint32_t tot;
        uint32_t const ex = block_excl_scan((nh << 16) | k, &tot);
        uint32_t const K = tot & 0xFFFFu;
        uint64_t const g = B0 + (ex & 0xFFFFu);                  // index of this thread's first kept base
        if ( k )
        {
                // compact the codes and the wildcard bits of the kept bytes, first byte most significant
                uint32_t cc = 0, nn = 0;
                for ( uint32_t m = kept; m; m &= m - 1 )
                {
                        int const i = __ffs(m) - 1;
                        cc = (cc << 2) | ((M.codes >> (2 * i)) & 3);
                        nn = (nn << 1) | ((M.nb >> i) & 1);
                }
                {
                        uint32_t const val = cc << (32 - 2 * k), sh = 2 * (uint32_t)(g & 15);
                        uint32_t const u = (uint32_t)((g >> 4) - 2 * (B0 >> 5));
                        atomicOr(&cu[u], val >> sh);
                        if ( sh + 2 * k > 32 ) atomicOr(&cu[u + 1], val << (32 - sh));
                }
                if ( nn )
                {
                        uint32_t const val = nn << (32 - k), sh = (uint32_t)(g & 31);
                        uint32_t const u = (uint32_t)((g >> 5) - 2 * (B0 >> 6));
                        atomicOr(&nu[u], val >> sh);
                        if ( sh + k > 32 ) atomicOr(&nu[u + 1], val << (32 - sh));
                }
        }
        if ( nh )
        {
                uint64_t r = tile_rec[blockIdx.x] + (ex >> 16);
                for ( uint32_t m = ends; m; m &= m - 1, ++r )
                {
                        int const i = __ffs(m) - 1;
                        rec_start[r] = g + __popc(kept & ((1u << i) - 1));
                        rec_nl[r] = byte0 + i;
                }
        }
        __syncthreads();
        if ( ! K ) return;
        // whole words are stored, the first and the last word of the tile may be shared with the neighbours
        uint64_t const B1 = B0 + K;
        for ( uint64_t w = (B0 >> 5) + threadIdx.x; w * 32 < B1; w += FA_THREADS )
        {
                uint32_t const j = (uint32_t)(w - (B0 >> 5));
                unsigned long long const val = ((unsigned long long)cu[2 * j] << 32) | cu[2 * j + 1];
                if ( w * 32 >= B0 && w * 32 + 32 <= B1 ) text[w] = val;
                else if ( val ) atomicOr(&text[w], val);
        }
        for ( uint64_t w = (B0 >> 6) + threadIdx.x; w * 64 < B1; w += FA_THREADS )
        {
                uint32_t const j = (uint32_t)(w - (B0 >> 6));
                unsigned long long const val = ((unsigned long long)nu[2 * j] << 32) | nu[2 * j + 1];
                if ( w * 64 >= B0 && w * 64 + 64 <= B1 ) nmask[w] = val;
                else if ( val ) atomicOr(&nmask[w], val);
        }
}

} // namespace realgpu
