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
// A thread owns 64 consecutive bytes (four 128-bit loads) as two pieces of 32; a byte is classified by ONE 8-byte
// shared-memory lookup whose words are accumulated with one multiply-add each (four flag masks in the byte lanes of one
// register, the 2-bit codes most significant first in another); inside a piece the state is a 32-bit carry chain
// (skipped when the piece holds no '>').  The first version (one table lookup and five shift/mask/or steps per byte,
// 16 bytes per thread, ranking every kept byte one by one) was instruction bound at 280 GB/s of file bytes.
// Algorithmic bytes per file byte: 1 read + (2 + 1)/8 written per kept base; the input is read twice (summary, pack).
#pragma once
#include "common.cuh"
#include "prims.cuh"

namespace realgpu
{

static const uint32_t FA_THREADS = 256;
static const uint32_t FA_PIECE = 32;                           // bytes of one piece (one bit per byte in a 32-bit mask)
static const uint32_t FA_PER_THREAD = 2 * FA_PIECE;
static const uint32_t FA_TILE = FA_THREADS * FA_PER_THREAD;    // bytes of one tile

// what a stretch of the file does to the scan state; counts are split at its first '>' or '\n' (the first "setter"):
// in front of it the entry state decides, behind it the stretch decides by itself
template<typename C>
struct FaSum
{
        uint32_t f;             // 0 = no setter (state passes through), 1 = leaves "no header", 2 = leaves "header"
        uint32_t pre_hdr;       // 1 = the first setter is a '\n': it files a record if the stretch is entered inside a header
        C pre_cnt;              // A C G T N in front of the first setter: kept if the stretch is entered outside a header
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

template<typename S> __device__ __forceinline__ S fa_shfl_down(S const & a, int o)
{
        S r;
        r.f = __shfl_down_sync(0xffffffffu, a.f, o); r.pre_hdr = __shfl_down_sync(0xffffffffu, a.pre_hdr, o);
        r.pre_cnt = __shfl_down_sync(0xffffffffu, a.pre_cnt, o); r.post_cnt = __shfl_down_sync(0xffffffffu, a.post_cnt, o);
        r.post_hdr = __shfl_down_sync(0xffffffffu, a.post_hdr, o);
        return r;
}
template<typename S> __device__ __forceinline__ S fa_shfl_up(S const & a, int o)
{
        S r;
        r.f = __shfl_up_sync(0xffffffffu, a.f, o); r.pre_hdr = __shfl_up_sync(0xffffffffu, a.pre_hdr, o);
        r.pre_cnt = __shfl_up_sync(0xffffffffu, a.pre_cnt, o); r.post_cnt = __shfl_up_sync(0xffffffffu, a.post_cnt, o);
        r.post_hdr = __shfl_up_sync(0xffffffffu, a.post_hdr, o);
        return r;
}

// table entry of a byte: x = flags in the low bit of the four byte lanes (kept outside headers, N, '>', '\n'), y = 2-bit code.
// mode 0 = text file (countReads.cpp:28-125): A C G T N are kept, everything else is dropped.
// mode 1 = pattern file (FastAReader::getNextPatternUnlocked, FastAReader.hpp:107-138): every byte that is not white space is a
// base of the read (SpaceTable.hpp:12-19); A C G T map to 0..3, anything else -- lower case included -- is a wildcard
// (Pattern::computeMapped, Pattern.hpp:105-128).
__device__ __forceinline__ uint2 fa_entry(uint32_t c, uint32_t mode)
{
        uint32_t const acgt = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? 1u : 0u;
        bool const space = c == ' ' || (c >= 9 && c <= 13);
        uint32_t const base = mode ? ((! space && c != '>') ? 1u : 0u) : ((acgt || c == 'N') ? 1u : 0u);
        uint32_t const wild = mode ? (base & (acgt ^ 1u)) : (c == 'N' ? 1u : 0u);
        uint32_t const code = c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 0u;
        return make_uint2(base | (wild << 8) | ((c == '>' ? 1u : 0u) << 16) | ((c == '\n' ? 1u : 0u) << 24), code);
}

// one bit per byte of a 32-byte piece (byte i = bit i); codes of the bytes 0..15 / 16..31, byte 0 (16) in the two top bits
struct FaMasks { uint32_t base, nb, gt, nl, c0, c1; };

__device__ __forceinline__ uint32_t fa_lut_flags(uint2 const & e) { return e.x; }
__device__ __forceinline__ uint32_t fa_lut_flags(uint32_t const & e) { return e; }
__device__ __forceinline__ uint32_t fa_lut_code(uint2 const & e) { return e.y; }
__device__ __forceinline__ uint32_t fa_lut_code(uint32_t const &) { return 0; }

// LUT = uint2 (flags + code) or uint32_t (flags only: the summary pass needs no codes)
template<typename LUT>
__device__ __forceinline__ FaMasks fa_classify(uint4 const v0, uint4 const v1, const LUT * lut)
{
        uint32_t const w[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
        uint32_t acc[4] = { 0, 0, 0, 0 }, c[2] = { 0, 0 };
        #pragma unroll
        for ( int i = 0; i < 32; ++i )
        {
                LUT const e = lut[(w[i >> 2] >> (8 * (i & 3))) & 0xFF];
                acc[i >> 3] += fa_lut_flags(e) << (i & 7);
                c[i >> 4] += fa_lut_code(e) << (30 - 2 * (i & 15));
        }
        FaMasks M;
        M.base = __byte_perm(__byte_perm(acc[0], acc[1], 0x0040), __byte_perm(acc[2], acc[3], 0x0040), 0x5410);
        M.nb   = __byte_perm(__byte_perm(acc[0], acc[1], 0x0051), __byte_perm(acc[2], acc[3], 0x0051), 0x5410);
        M.gt   = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
        M.nl   = __byte_perm(__byte_perm(acc[0], acc[1], 0x0073), __byte_perm(acc[2], acc[3], 0x0073), 0x5410);
        M.c0 = c[0]; M.c1 = c[1];
        return M;
}

// bit i = the state in front of byte i (1 = inside a header) of a piece entered in state `in`; *out = the state behind it
__device__ __forceinline__ uint32_t fa_states(uint32_t gt, uint32_t nl, uint32_t in, uint32_t * out)
{
        if ( gt == 0 )
        {
                // no '>': the entry state lasts up to and including the first '\n'
                *out = nl ? 0u : in;
                return in ? (((nl & (0u - nl)) << 1) - 1u) : 0u;
        }
        uint32_t const set = gt | nl;
        uint32_t G = gt, P = ~set;
        G |= (G << 1) & P; P &= P << 1;
        G |= (G << 2) & P; P &= P << 2;
        G |= (G << 4) & P; P &= P << 4;
        G |= (G << 8) & P; P &= P << 8;
        G |= (G << 16) & P;
        uint32_t const after = G | (in ? ((set & (0u - set)) - 1u) : 0u);       // bit i = state behind byte i
        *out = after >> 31;
        return (after << 1) | in;
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
        uint32_t o;
        uint32_t const st = fa_states(M.gt, M.nl, 0, &o);
        s.f = ((M.gt >> (31 - __clz(set))) & 1) ? 2 : 1;
        s.pre_hdr = (M.nl & first) ? 1 : 0;
        s.pre_cnt = __popc(M.base & front);
        s.post_cnt = __popc(M.base & ~st & behind);
        s.post_hdr = __popc(M.nl & st & behind);
        return s;
}

// 16 bytes at file offset off (a multiple of 16; the buffer is 16-byte aligned); bytes behind the end of the file read as blanks,
// dropped bytes that leave the state alone in both modes
__device__ __forceinline__ uint4 fa_load(const uint8_t * bytes, uint64_t nbytes, uint64_t off)
{
        if ( off + 16 <= nbytes )
                return __ldg(reinterpret_cast<const uint4 *>(bytes + off));
        uint32_t w[4] = { 0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u };
        for ( int i = 0; i < 16; ++i )
                if ( off + i < nbytes ) w[i >> 2] = (w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)bytes[off + i] << (8 * (i & 3)));
        return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(FA_THREADS) k_fa_summary(const uint8_t * __restrict__ bytes, uint64_t nbytes, FaSum32 * __restrict__ sums, uint32_t mode)
{
        __shared__ uint32_t lut[256];
        __shared__ FaSum32 wsum[FA_THREADS / 32];
        lut[threadIdx.x] = fa_entry(threadIdx.x, mode).x;
        uint64_t const byte0 = (uint64_t)blockIdx.x * FA_TILE + threadIdx.x * FA_PER_THREAD;
        uint4 v[4];
        #pragma unroll
        for ( int j = 0; j < 4; ++j ) v[j] = fa_load(bytes, nbytes, byte0 + 16 * j);
        __syncthreads();
        FaSum32 s = fa_piece_summary(fa_classify(v[0], v[1], lut));
        s = fa_combine(s, fa_piece_summary(fa_classify(v[2], v[3], lut)));
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

// one CTA, FA_SCAN_THREADS tiles at a time; tile_base[t] = (index of the tile's first kept base << 1) | entry state,
// tile_rec[t] = index of its first record; totals[0] = kept bases, totals[1] = records of the whole file
static const uint32_t FA_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(FA_SCAN_THREADS) k_fa_scan(const FaSum32 * __restrict__ sums, uint64_t ntiles, uint64_t * __restrict__ tile_base,
                                                             uint64_t * __restrict__ tile_rec, uint64_t * __restrict__ totals)
{
        __shared__ FaSum64 wtot[FA_SCAN_THREADS / 32];          // per warp: its total, then everything in front of it in the chunk
        __shared__ FaSum64 carry_s;
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        FaSum64 ident; ident.f = 0; ident.pre_hdr = 0; ident.pre_cnt = 0; ident.post_cnt = 0; ident.post_hdr = 0;
        FaSum64 carry = ident;                                  // everything in front of the chunk
        for ( uint64_t c0 = 0; c0 < ntiles; c0 += FA_SCAN_THREADS )
        {
                uint64_t const t = c0 + threadIdx.x;
                FaSum64 own = ident;
                if ( t < ntiles ) own = fa_combine(ident, sums[t]);
                FaSum64 inc = own;
                #pragma unroll
                for ( int o = 1; o < 32; o <<= 1 )
                {
                        FaSum64 const p = fa_shfl_up(inc, o);
                        if ( lane >= o ) inc = fa_combine(p, inc);
                }
                FaSum64 exc = fa_shfl_up(inc, 1);               // everything in front of this thread inside its warp
                if ( lane == 0 ) exc = ident;
                if ( lane == 31 ) wtot[wid] = inc;
                __syncthreads();
                if ( wid == 0 )
                {
                        FaSum64 const mine = wtot[lane];
                        FaSum64 winc = mine;
                        #pragma unroll
                        for ( int o = 1; o < 32; o <<= 1 )
                        {
                                FaSum64 const p = fa_shfl_up(winc, o);
                                if ( lane >= o ) winc = fa_combine(p, winc);
                        }
                        FaSum64 wexc = fa_shfl_up(winc, 1);
                        if ( lane == 0 ) wexc = ident;
                        wtot[lane] = wexc;
                        if ( lane == 31 ) carry_s = fa_combine(carry, winc);
                }
                __syncthreads();
                FaSum64 const pre = fa_combine(carry, fa_combine(wtot[wid], exc));
                if ( t < ntiles )
                {
                        // the file is entered outside a header: the "pre" parts count as kept bases and file no record
                        tile_base[t] = ((pre.pre_cnt + pre.post_cnt) << 1) | (pre.f == 2 ? 1u : 0u);
                        tile_rec[t] = pre.post_hdr;
                }
                carry = carry_s;
                __syncthreads();
        }
        if ( threadIdx.x == 0 )
        {
                totals[0] = carry.pre_cnt + carry.post_cnt;
                totals[1] = carry.post_hdr;
        }
}

// removes the 2-bit fields of the bytes whose bit is set in drop (16 bytes, byte j = bits 31-2j..30-2j); the rest closes up towards the top
__device__ __forceinline__ uint32_t fa_squeeze(uint32_t c, uint32_t drop)
{
        while ( drop )
        {
                int const p = 31 - __clz(drop);                 // highest byte first: the fields in front of it stay where they are
                drop ^= 1u << p;
                uint32_t const below = (1u << (30 - 2 * p)) - 1u;
                c = (c & ~(below * 4u + 3u)) | ((c & below) << 2);
        }
        return c;
}

static const uint32_t FA_CU = FA_TILE / 16 + 4, FA_NU = FA_TILE / 32 + 4;

// text / nmask: word 0 of the (zeroed) arrays; rec_start / rec_nl: one entry per record (rec_nl = file offset of the '\n' that filed it)
__global__ void __launch_bounds__(FA_THREADS) k_fa_pack(const uint8_t * __restrict__ bytes, uint64_t nbytes, const uint64_t * __restrict__ tile_base,
                                                        const uint64_t * __restrict__ tile_rec, unsigned long long * __restrict__ text,
                                                        unsigned long long * __restrict__ nmask, uint64_t * __restrict__ rec_start, uint64_t * __restrict__ rec_nl,
                                                        uint64_t * __restrict__ rec_open, uint32_t mode)
{
        __shared__ uint2 lut[256];
        __shared__ uint32_t wf[FA_THREADS / 32];
        __shared__ uint32_t cu[FA_CU];                          // 2-bit codes, 16 bases per unit, most significant first; unit pairs = text words
        __shared__ uint32_t nu[FA_NU];                          // wildcard bits, 32 bases per unit; unit pairs = mask words
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        lut[threadIdx.x] = fa_entry(threadIdx.x, mode);
        for ( uint32_t i = threadIdx.x; i < FA_CU; i += FA_THREADS ) cu[i] = 0;
        for ( uint32_t i = threadIdx.x; i < FA_NU; i += FA_THREADS ) nu[i] = 0;
        uint64_t const byte0 = (uint64_t)blockIdx.x * FA_TILE + threadIdx.x * FA_PER_THREAD;
        uint4 v[4];
        #pragma unroll
        for ( int j = 0; j < 4; ++j ) v[j] = fa_load(bytes, nbytes, byte0 + 16 * j);
        uint64_t const tb = tile_base[blockIdx.x];
        uint64_t const B0 = tb >> 1;
        __syncthreads();
        FaMasks M[2];
        M[0] = fa_classify(v[0], v[1], lut);
        M[1] = fa_classify(v[2], v[3], lut);
        // entry state of this thread: the last setter in front of it -- in its warp, else in an earlier warp, else the tile's
        uint32_t mine = 0;
        #pragma unroll
        for ( int s = 0; s < 2; ++s )
        {
                uint32_t const set = M[s].gt | M[s].nl;
                if ( set ) mine = ((M[s].gt >> (31 - __clz(set))) & 1) ? 2u : 1u;
        }
        uint32_t const m_set = __ballot_sync(0xffffffffu, mine != 0), m_one = __ballot_sync(0xffffffffu, mine == 2);
        if ( lane == 0 ) wf[wid] = m_set ? (((m_one >> (31 - __clz(m_set))) & 1) ? 2u : 1u) : 0u;
        __syncthreads();
        uint32_t in = (uint32_t)(tb & 1);
        for ( int w = 0; w < wid; ++w ) if ( wf[w] ) in = wf[w] == 2;
        uint32_t const lower = m_set & ((1u << lane) - 1);
        if ( lower ) in = (m_one >> (31 - __clz(lower))) & 1;
        uint32_t kept[2], ends[2], opens[2];
        #pragma unroll
        for ( int s = 0; s < 2; ++s )
        {
                uint32_t out;
                uint32_t const st = fa_states(M[s].gt, M[s].nl, in, &out);
                kept[s] = M[s].base & ~st; ends[s] = M[s].nl & st; opens[s] = M[s].gt & ~st;
                in = out;
        }
        uint32_t const k = __popc(kept[0]) + __popc(kept[1]), nh = __popc(ends[0]) + __popc(ends[1]);
        uint32_t tot;
        uint32_t const ex = block_excl_scan((nh << 16) | k, &tot);
        uint32_t const K = tot & 0xFFFFu;
        uint64_t g = B0 + (ex & 0xFFFFu);                        // index of this thread's next kept base
        uint64_t r = (nh || (rec_open && (opens[0] | opens[1]))) ? tile_rec[blockIdx.x] + (ex >> 16) : 0;
        if ( rec_open )
        {
                // the '>' that opens header number k (pattern files: the id starts behind it): every header opened earlier has been
                // closed when another one opens, so k = the records filed in front of the '>'
                uint64_t rr = r;
                #pragma unroll
                for ( int s = 0; s < 2; ++s )
                {
                        for ( uint32_t m = opens[s]; m; m &= m - 1 )
                        {
                                int const i = __ffs(m) - 1;
                                rec_open[rr + __popc(ends[s] & ((1u << i) - 1))] = byte0 + 32 * s + i;
                        }
                        rr += __popc(ends[s]);
                }
        }
        #pragma unroll
        for ( int s = 0; s < 2; ++s )
        {
                if ( ends[s] )
                        for ( uint32_t m = ends[s]; m; m &= m - 1, ++r )
                        {
                                int const i = __ffs(m) - 1;
                                rec_start[r] = g + __popc(kept[s] & ((1u << i) - 1));
                                rec_nl[r] = byte0 + 32 * s + i;
                        }
                uint32_t const nk = M[s].nb & kept[s];
                if ( nk )
                        for ( uint32_t m = nk; m; m &= m - 1 )
                        {
                                uint64_t const x = g + __popc(kept[s] & ((1u << (__ffs(m) - 1)) - 1));
                                atomicOr(&nu[(uint32_t)((x >> 5) - 2 * (B0 >> 6))], 0x80000000u >> (x & 31));
                        }
                #pragma unroll
                for ( int hh = 0; hh < 2; ++hh )
                {
                        uint32_t const k16 = (kept[s] >> (16 * hh)) & 0xFFFFu;
                        if ( ! k16 ) continue;
                        uint32_t const kk = __popc(k16);
                        uint32_t const val = fa_squeeze(hh ? M[s].c1 : M[s].c0, k16 ^ 0xFFFFu);
                        uint32_t const sh = 2 * (uint32_t)(g & 15);
                        uint32_t const u = (uint32_t)((g >> 4) - 2 * (B0 >> 5));
                        atomicOr(&cu[u], val >> sh);
                        if ( sh + 2 * kk > 32 ) atomicOr(&cu[u + 1], val << (32 - sh));
                        g += kk;
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

// ---- K0 for pattern files: FASTA reads on the device ----------------------------------------------------------------
// k_fa_summary / k_fa_scan / k_fa_pack in mode 1 leave: the bases of all reads as ONE 2-bit stream (wildcards as code 0 + a
// mask bit), read_start[r] = index of read r's first base (read_start[nreads] = all bases), and per read the file offsets of
// the '>' that opens its id line and of the '\n' that closes it.  Bases in front of the first '>' belong to no read
// (FastAReader::findFirstMarker); an id line the file ends in without a '\n' opens no read (FastAReader.hpp:120-121).
// The kernels below turn that into what real_gpu_set_reads_packed takes -- every read on a byte boundary, 4 bases per byte --
// in the order perm gives (the rewritten order of the reference, or file order), plus the ids for K8.

// per read: length, wildcard flag, packed bytes, id bytes
__global__ void __launch_bounds__(256) k_rd_table(const uint64_t * __restrict__ read_start, const uint64_t * __restrict__ nmask, const uint64_t * __restrict__ rec_open,
                                                  const uint64_t * __restrict__ rec_nl, uint64_t nreads, uint32_t * __restrict__ len, uint32_t * __restrict__ wild,
                                                  uint32_t * __restrict__ idlen, unsigned int * __restrict__ maxlen /* [3]: max length, min key, max key */)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t L = 0;
        if ( r < nreads )
        {
                uint64_t const b = read_start[r], e = read_start[r+1];
                L = (uint32_t)min((unsigned long long)(e - b), 0xFFFFFFFFull);
                len[r] = L;
                bool w = false;
                if ( L )
                {
                        uint64_t const first = b >> 6, last = (e - 1) >> 6;
                        for ( uint64_t x = first; x <= last && ! w; ++x )
                        {
                                uint64_t m = nmask[x];
                                if ( x == first ) m &= (~0ULL) >> (b & 63);
                                if ( x == last ) m &= (~0ULL) << (63 - ((e - 1) & 63));
                                w = m != 0;
                        }
                }
                wild[r] = w ? 1u : 0u;
                idlen[r] = (uint32_t)(rec_nl[r] - rec_open[r] - 1);
        }
        // key of the rewritten order: (length, has a wildcard); when all reads share one key the order is the file order
        uint32_t mx = L, kmin = 0xFFFFFFFFu, kmax = 0;
        if ( r < nreads ) { uint32_t const key = (min(L, 0x7FFFFFFFu) << 1) | wild[r]; kmin = key; kmax = key; }
        #pragma unroll
        for ( int o = 16; o > 0; o >>= 1 )
        {
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if ( (threadIdx.x & 31) == 0 )
        {
                if ( mx ) atomicMax(maxlen, mx);
                atomicMin(maxlen + 1, kmin);
                atomicMax(maxlen + 2, kmax);
        }
}

// 64-bit sums of two u32 arrays (the totals their 32-bit exclusive scans must stay below)
__global__ void __launch_bounds__(256) k_rd_sums(const uint32_t * __restrict__ a, const uint32_t * __restrict__ b, uint64_t n, unsigned long long * __restrict__ out)
{
        unsigned long long sa = 0, sb = 0;
        for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x ) { sa += a[i]; sb += b[i]; }
        #pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
        if ( (threadIdx.x & 31) == 0 ) { if ( sa ) atomicAdd(out, sa); if ( sb ) atomicAdd(out + 1, sb); }
}

// slot j of the output order holds read perm[j] (perm == nullptr: file order): lengths, flags, sizes in output order
__global__ void __launch_bounds__(256) k_rd_permute(const uint32_t * __restrict__ perm, uint64_t nreads, const uint32_t * __restrict__ len, const uint32_t * __restrict__ wild,
                                                    const uint32_t * __restrict__ idlen, uint32_t * __restrict__ olen, uint8_t * __restrict__ oflag,
                                                    uint32_t * __restrict__ obytes, uint32_t * __restrict__ oidlen)
{
        uint64_t const j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( j >= nreads ) return;
        uint32_t const r = perm ? perm[j] : (uint32_t)j;
        uint32_t const L = len[r];
        olen[j] = L; oflag[j] = (uint8_t)wild[r]; obytes[j] = (L + 3) >> 2; oidlen[j] = idlen[r];
}

// the packed bytes and the id of output slot j; offsets = exclusive scans of obytes / oidlen (u32: the totals are checked by the host)
__global__ void __launch_bounds__(256) k_rd_repack(const uint32_t * __restrict__ perm, uint64_t nreads, const uint64_t * __restrict__ stream,
                                                   const uint64_t * __restrict__ read_start, const uint32_t * __restrict__ boff32, uint8_t * __restrict__ packed,
                                                   uint64_t * __restrict__ boff64, const uint8_t * __restrict__ file, const uint64_t * __restrict__ rec_open,
                                                   const uint32_t * __restrict__ oidlen, const uint32_t * __restrict__ ioff32, char * __restrict__ ids, uint64_t * __restrict__ ioff64,
                                                   uint64_t total_bytes, uint64_t total_id)
{
        uint64_t const j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( j > nreads ) return;
        if ( j == nreads ) { boff64[j] = total_bytes; ioff64[j] = total_id; return; }
        uint32_t const r = perm ? perm[j] : (uint32_t)j;
        uint64_t const b = read_start[r];
        uint32_t const L = (uint32_t)(read_start[r+1] - b);
        uint8_t * o = packed + boff32[j];
        boff64[j] = boff32[j];
        for ( uint32_t w = 0; w * 32 < L; ++w )
        {
                uint32_t const n = min(32u, L - 32 * w);
                uint64_t const v = text_word(stream, b + 32 * w, n) << (64 - 2 * n);        // left aligned, zero filled
                uint32_t const nb = (n + 3) >> 2;
                #pragma unroll
                for ( uint32_t k = 0; k < 8; ++k )
                        if ( k < nb ) o[8 * w + k] = (uint8_t)(v >> (56 - 8 * k));
        }
        char * io = ids + ioff32[j];
        ioff64[j] = ioff32[j];
        const uint8_t * src = file + rec_open[r] + 1;
        uint32_t const n = oidlen[j];
        for ( uint32_t k = 0; k < n; ++k ) io[k] = (char)src[k];
}

} // namespace realgpu
