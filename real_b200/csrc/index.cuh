// K1 (read packing, seed extraction) and K2 (read-side signature index) kernels.
//
// K1 replaces, per read, Pattern::computeMapped (Pattern.hpp:105-128),
// SignatureConstruction::signatureMapped / reverseMappedSignature
// (SignatureConstruction.hpp:218-280, 347-410) and RestWordBuffer::setupStraight/setupReverse
// (RestWordBuffer.hpp:56-78): both strands of every read are packed 2 bit/base, MSB first, into
// W words, and the 4-fragment seed of each strand is kept as one word.
//
// K2 replaces ListSet::sort + getLookupTable (ListSet.hpp:41-63, getLookupTable.hpp:25-51), on the
// READ side (BASELINE.json north_star).  The six pair lists collapse into three tables because a
// text position x sees the same 2-fragment key for several lists:
//   table A (adjacent fragments)  key(x) = bases [x,x+2F)               lists 0,3,5 at window x, x-F, x-2F
//   table B (one fragment apart)  key(x) = [x,x+F) ++ [x+2F,x+3F)       lists 1,4   at window x, x-F
//   table C (two apart)           key(x) = [x,x+F) ++ [x+3F,x+4F)       list  2     at window x
// Each table is a presence array over signature slots -- one 8-byte SlotWord {32 presence bits, rank of the
// word's first slot} per 32 slots, so that one 8-byte probe yields the bit and the entry index -- plus an
// entry array addressed by rank.
#pragma once

#include "common.cuh"
#include "prims.cuh"

namespace realgpu
{

// list id of (table, fragment offset t) and the fragment pair it keys on
__host__ __device__ __forceinline__ int list_of(int table, int t)
{
        // A: (0,1)->0 (1,2)->3 (2,3)->5 ; B: (0,2)->1 (1,3)->4 ; C: (0,3)->2
        return table == 0 ? (t == 0 ? 0 : (t == 1 ? 3 : 5)) : (table == 1 ? (t == 0 ? 1 : 4) : 2);
}
__host__ __device__ __forceinline__ int pair_second(int table, int t) { return t + 1 + table; }

// which (table,t) entries exist for a given seed error budget: a seed with at most s mismatching
// bases has at least 4-s exact fragments, and a match is reported through the pair made of its two
// LOWEST exact fragments only, so pairs that can never be that pair are not indexed.
__host__ __device__ __forceinline__ int table_lists(int table, uint32_t seedkmax)
{
        if ( seedkmax >= 2 ) return table == 0 ? 3 : (table == 1 ? 2 : 1);
        if ( seedkmax == 1 ) return table == 0 ? 2 : (table == 1 ? 1 : 0);
        return table == 0 ? 1 : 0;
}

// ---- K1 --------------------------------------------------------------------------------------

// reverse complement of the `len` bases held right aligned in x
__device__ __forceinline__ uint64_t revcomp_word(uint64_t x, uint32_t len)
{
        uint64_t y = __brevll(~x);                                                     // complement, reverse all bits
        y = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);   // put the two bits of every base back in order
        return y >> (64 - 2*len);
}

// `len` (1..32) mapped bytes starting at p, packed 2 bit/base MSB first and left aligned; *bad is set when a
// byte is not 0..3.  The bytes are fetched as aligned 32-bit words and realigned with funnel shifts; four
// bases are packed at a time with one multiply: for x = b0 | b1<<8 | b2<<16 | b3<<24 (each 0..3) the top byte
// of x * 0x40100401 is b0<<6 | b1<<4 | b2<<2 | b3 (the partial products land on disjoint bits).
__device__ __forceinline__ uint64_t pack_bytes(const uint8_t * __restrict__ p, uint32_t len, uint32_t & bad)
{
        uintptr_t const addr = reinterpret_cast<uintptr_t>(p);
        const uint32_t * q = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        uint32_t const sh = (uint32_t)(addr & 3) * 8;
        uint32_t const nq = (uint32_t)((addr & 3) + len + 3) >> 2;        // aligned words that hold the bytes
        uint32_t w[9];
        #pragma unroll
        for ( uint32_t i = 0; i < 9; ++i )
                w[i] = (i < nq) ? __ldg(q + i) : 0u;
        uint64_t word = 0;
        #pragma unroll
        for ( uint32_t i = 0; i < 8; ++i )
        {
                uint32_t x = __funnelshift_r(w[i], w[i+1], sh);
                int const rem = (int)len - 4 * (int)i;                      // bytes of this group that belong to the read
                if ( rem < 4 ) x &= (rem <= 0) ? 0u : (0xFFFFFFFFu >> (8 * (4 - rem)));
                bad |= x & 0xFCFCFCFCu;
                uint32_t const packed = ((x & 0x03030303u) * 0x40100401u) >> 24;
                word |= (uint64_t)packed << (56 - 8 * i);
        }
        return word;
}

// one thread per (read, strand, word): word w of the '+' strand is read[32w, 32w+32); word w of the '-' strand
// is the reverse complement of read[L-32w-len, L-32w)
__global__ void __launch_bounds__(256) k_pack_reads(const uint8_t * __restrict__ mapped, const uint64_t * __restrict__ offsets,
                                                  uint64_t nreads, uint32_t W, uint64_t * __restrict__ rpack, uint32_t * __restrict__ bad)
{
        uint64_t const gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( gid >= nreads * 2 * W ) return;
        uint32_t const w = (uint32_t)(gid % W);
        uint64_t const rs = gid / W;
        uint32_t const s = (uint32_t)(rs & 1);
        uint64_t const r = rs >> 1;
        uint64_t const o = offsets[r];
        uint32_t const L = (uint32_t)(offsets[r+1] - o);
        uint64_t word = 0;
        if ( 32 * w < L )
        {
                uint32_t const len = min(32u, L - 32 * w);
                uint32_t anybad = 0;
                if ( s == 0 )
                {
                        word = pack_bytes(mapped + o + 32 * w, len, anybad);
                        if ( anybad ) bad[r] = 1;
                }
                else
                {
                        uint64_t const fwd = pack_bytes(mapped + o + (L - 32 * w - len), len, anybad);
                        word = revcomp_word(fwd >> (64 - 2 * len), len) << (64 - 2 * len);
                }
        }
        rpack[gid] = word;
}

// Reads that arrive 2 bit/base, 4 bases per byte MSB first, every read starting on a byte boundary -- the layout
// of the reference's rewritten pattern file (TemporaryFile.hpp:231-268, writePatternDontCareFree).  `len` bases
// starting at base `first` (a multiple of 32) of the read whose packed bytes start at p, left aligned in a word.
__device__ __forceinline__ uint64_t load_packed_word(const uint8_t * __restrict__ p, uint32_t first, uint32_t len)
{
        uintptr_t const addr = reinterpret_cast<uintptr_t>(p) + (first >> 2);
        const uint32_t * q = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
        uint32_t const sh = (uint32_t)(addr & 3) * 8;
        uint32_t const nbytes = (len + 3) >> 2;
        uint32_t const nq = (uint32_t)((addr & 3) + nbytes + 3) >> 2;
        uint32_t const w0 = __ldg(q), w1 = nq > 1 ? __ldg(q + 1) : 0u, w2 = nq > 2 ? __ldg(q + 2) : 0u;
        // bytes in memory order -> most significant byte first
        uint32_t const hi = __byte_perm(__funnelshift_r(w0, w1, sh), 0, 0x0123);
        uint32_t const lo = __byte_perm(__funnelshift_r(w1, w2, sh), 0, 0x0123);
        uint64_t const v = ((uint64_t)hi << 32) | lo;
        return len == 32 ? v : (v & (~0ULL << (64 - 2 * len)));
}

// one thread per (read, strand, word), packed input; offsets/lengths null = all reads have `uniform` bases
__global__ void __launch_bounds__(256) k_pack_reads_packed(const uint8_t * __restrict__ packed, const uint64_t * __restrict__ byte_offsets,
                                                         const uint64_t * __restrict__ offsets, uint32_t uniform, uint64_t nreads, uint32_t W,
                                                         uint64_t * __restrict__ rpack)
{
        uint64_t const gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( gid >= nreads * 2 * W ) return;
        uint32_t const w = (uint32_t)(gid % W);
        uint64_t const rs = gid / W;
        uint32_t const s = (uint32_t)(rs & 1);
        uint64_t const r = rs >> 1;
        uint32_t const L = offsets ? (uint32_t)(offsets[r+1] - offsets[r]) : uniform;
        const uint8_t * p = packed + (byte_offsets ? byte_offsets[r] : r * (uint64_t)((uniform + 3) >> 2));
        uint64_t word = 0;
        if ( 32 * w < L )
        {
                uint32_t const len = min(32u, L - 32 * w);
                if ( s == 0 )
                        word = load_packed_word(p, 32 * w, len);
                else
                {
                        // reverse complement of read[L-32w-len, L-32w): fetch the 64 bases around it and cut them out
                        uint32_t const b0 = L - 32 * w - len;                     // first base of the stretch
                        uint32_t const a0 = b0 & ~31u;                            // word-aligned base in front of it
                        uint64_t const x0 = load_packed_word(p, a0, min(32u, L - a0));
                        uint64_t const x1 = (a0 + 32 < L) ? load_packed_word(p, a0 + 32, min(32u, L - a0 - 32)) : 0ULL;
                        uint32_t const o = 2 * (b0 - a0);
                        uint64_t const fwd = (o ? ((x0 << o) | (x1 >> (64 - o))) : x0) >> (64 - 2 * len);
                        word = revcomp_word(fwd, len) << (64 - 2 * len);
                }
        }
        rpack[gid] = word;
}

__global__ void __launch_bounds__(256) k_uniform_offsets(uint64_t * __restrict__ offsets, uint64_t nreads, uint32_t L)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i <= nreads ) offsets[i] = i * L;
}

__global__ void __launch_bounds__(256) k_flags_to_bad(const uint8_t * __restrict__ flags, uint64_t nreads, uint32_t * __restrict__ bad)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i < nreads ) bad[i] = flags[i] ? 1u : 0u;
}

// one thread per read: usable length and the two strand seeds.
// seed word = seedl bases right aligned, fragment 0 in the top bits (what getTextWord(p,seedl) yields for the
// text window the strand is laid over); the '-' seed is the LAST seedl bases of the reverse complement strand
// (RestMatch.hpp:84-89) = the reverse complement of the first seedl bases of the read.
__global__ void __launch_bounds__(256) k_read_seeds(const uint64_t * __restrict__ offsets, uint64_t nreads, uint32_t W, uint32_t seedl,
                                                  const uint64_t * __restrict__ rpack, const uint32_t * __restrict__ bad,
                                                  uint32_t * __restrict__ rlen, uint64_t * __restrict__ seeds, uint32_t * __restrict__ usable)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r >= nreads ) return;
        uint32_t const L = (uint32_t)(offsets[r+1] - offsets[r]);
        bool const ok = (L >= seedl) && ! bad[r];
        rlen[r] = ok ? L : 0;
        usable[r] = ok ? 1 : 0;
        uint64_t const sf = rpack[(2*r) * W] >> (64 - 2*seedl);
        seeds[2*r] = ok ? sf : 0;
        seeds[2*r+1] = ok ? revcomp_word(sf, seedl) : 0;
}

// ---- K2 --------------------------------------------------------------------------------------

struct TableGeom
{
        uint32_t F;          // bases per fragment
        uint32_t keybits;    // 4F
        uint32_t hb;         // log2 slots
        uint32_t nlists;     // fragment offsets indexed in this table
        int table;           // 0,1,2
};

__device__ __forceinline__ uint64_t pair_key(uint64_t seed, uint32_t F, int a, int b)
{
        uint64_t const fm = (F == 32) ? ~0ULL : ((1ULL << (2*F)) - 1);
        uint64_t const ma = (seed >> (2*F*(3-a))) & fm;
        uint64_t const mb = (seed >> (2*F*(3-b))) & fm;
        return (ma << (2*F)) | mb;
}

// The table build does not sort.  (1) The entries of a table -- (strand seed, strand id, fragment offset),
// nlists per usable read strand -- are grouped by the top bits of their slot with one staged 256-way
// partition pass (histogram, then a scatter that lays the records out bucket by bucket in shared memory
// and writes whole runs).  (2..4) Walking the grouped entries in order then only ever touches one
// 1/256 slice of the table at a time, which stays in L2, so the table can be built with plain L2
// atomics: set the presence bits, rank them (sector popcounts + scan), and claim E[rank] with a
// compare-and-swap; entries that lose the claim (same slot) go to an overflow area behind the distinct
// entries and are pushed on the head's chain with an exchange.
static const int EP_IDS_PER_THREAD = 4;
static const int EP_TILE_IDS = 256 * EP_IDS_PER_THREAD;       // strand ids per tile
static const int EP_TILE_ENTRIES = EP_TILE_IDS * 3;
static const int EP_MAX_BUCKETS = 256;
static const int EP_CURSOR_STRIDE = 32;

struct EntryPartParams
{
        const uint64_t * seeds;       // 2*nreads strand seeds
        const uint32_t * usable;      // nreads
        uint64_t nids;                // 2*nreads
        TableGeom G;
        uint32_t ebits;               // bucket = slot >> (hb - ebits)
        uint64_t * ent_seed;          // grouped output
        uint32_t * ent_val;
        uint32_t * bucket_count;      // [256]
        uint32_t * bucket_start;      // [257]
        uint32_t * bucket_cursor;     // [256 * EP_CURSOR_STRIDE]
};

__device__ __forceinline__ uint32_t entry_slot(uint64_t seed, TableGeom const & G, uint32_t t)
{
        return slot_of(pair_key(seed, G.F, (int)t, pair_second(G.table, (int)t)), G.keybits, G.hb);
}

__global__ void __launch_bounds__(256) k_ent_hist(EntryPartParams P)
{
        __shared__ uint32_t cnt[EP_MAX_BUCKETS];
        cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t const sh = P.G.hb - P.ebits;
        for ( uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; id < P.nids; id += (uint64_t)gridDim.x * blockDim.x )
        {
                if ( ! P.usable[id >> 1] ) continue;
                uint64_t const seed = P.seeds[id];
                for ( uint32_t t = 0; t < P.G.nlists; ++t )
                        atomicAdd(&cnt[P.ebits ? (entry_slot(seed, P.G, t) >> sh) : 0u], 1u);
        }
        __syncthreads();
        if ( cnt[threadIdx.x] ) atomicAdd(P.bucket_count + threadIdx.x, cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(EP_MAX_BUCKETS) k_ent_offsets(EntryPartParams P, uint32_t * total)
{
        __shared__ uint32_t sc[EP_MAX_BUCKETS];
        sc[threadIdx.x] = P.bucket_count[threadIdx.x];
        __syncthreads();
        if ( threadIdx.x == 0 )
        {
                uint32_t a = 0;
                for ( int b = 0; b < EP_MAX_BUCKETS; ++b ) { P.bucket_start[b] = a; a += sc[b]; }
                P.bucket_start[EP_MAX_BUCKETS] = a;
                *total = a;
        }
        P.bucket_cursor[threadIdx.x * EP_CURSOR_STRIDE] = 0;
}

struct EntryPartSmem
{
        uint64_t stage_seed[EP_TILE_ENTRIES];
        uint32_t stage_val[EP_TILE_ENTRIES];
        uint32_t wcnt[8][EP_MAX_BUCKETS];
        uint32_t loc[EP_MAX_BUCKETS + 1];
        uint32_t base[EP_MAX_BUCKETS];
        uint8_t stage_b[EP_TILE_ENTRIES];
};

__global__ void __launch_bounds__(256) k_ent_scatter(EntryPartParams P)
{
        extern __shared__ __align__(16) unsigned char ep_smem[];
        EntryPartSmem & S = *reinterpret_cast<EntryPartSmem *>(ep_smem);
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t const lt = (1u << lane) - 1;
        uint32_t const sh = P.G.hb - P.ebits;
        uint32_t const nl = P.G.nlists;
        #pragma unroll
        for ( int w = 0; w < 8; ++w ) S.wcnt[w][threadIdx.x] = 0;
        __syncthreads();

        uint64_t const ntiles = (P.nids + EP_TILE_IDS - 1) / EP_TILE_IDS;
        for ( uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
        {
                uint64_t seed[EP_IDS_PER_THREAD];
                bool ok[EP_IDS_PER_THREAD];
                #pragma unroll
                for ( int k = 0; k < EP_IDS_PER_THREAD; ++k )
                {
                        uint64_t const id = tile * EP_TILE_IDS + (uint64_t)k * 256 + threadIdx.x;
                        ok[k] = (id < P.nids) && P.usable[id >> 1];
                        seed[k] = ok[k] ? P.seeds[id] : 0;
                }
                // (1) per-warp bucket counts
                #pragma unroll
                for ( int k = 0; k < EP_IDS_PER_THREAD; ++k )
                        if ( ok[k] )
                                for ( uint32_t t = 0; t < nl; ++t )
                                        atomicAdd(&S.wcnt[wid][P.ebits ? (entry_slot(seed[k], P.G, t) >> sh) : 0u], 1u);
                __syncthreads();
                // (2) staging layout, global run reservation, per-warp running slots
                {
                        uint32_t tot = 0;
                        #pragma unroll
                        for ( int w = 0; w < 8; ++w ) tot += S.wcnt[w][threadIdx.x];
                        uint32_t blocktot;
                        uint32_t const ex = block_excl_scan(tot, &blocktot);
                        S.loc[threadIdx.x] = ex;
                        if ( threadIdx.x == EP_MAX_BUCKETS - 1 ) S.loc[EP_MAX_BUCKETS] = blocktot;
                        S.base[threadIdx.x] = tot ? (P.bucket_start[threadIdx.x] + atomicAdd(P.bucket_cursor + threadIdx.x * EP_CURSOR_STRIDE, tot)) : 0;
                        uint32_t run = ex;
                        #pragma unroll
                        for ( int w = 0; w < 8; ++w )
                        {
                                uint32_t const c = S.wcnt[w][threadIdx.x];
                                S.wcnt[w][threadIdx.x] = run;
                                run += c;
                        }
                }
                __syncthreads();
                // (3) warp multisplit into the staging area
                #pragma unroll
                for ( int k = 0; k < EP_IDS_PER_THREAD; ++k )
                {
                        uint64_t const id = tile * EP_TILE_IDS + (uint64_t)k * 256 + threadIdx.x;
                        for ( uint32_t t = 0; t < nl; ++t )
                        {
                                uint32_t const b = (ok[k] && P.ebits) ? (entry_slot(seed[k], P.G, t) >> sh) : 0u;
                                uint32_t const peers = __match_any_sync(0xffffffffu, ok[k] ? b : 0x100u);
                                uint32_t const below = __popc(peers & lt);
                                uint32_t pre = 0;
                                if ( ok[k] ) pre = S.wcnt[wid][b];
                                __syncwarp();
                                if ( ok[k] && below == 0 ) S.wcnt[wid][b] = pre + __popc(peers);
                                __syncwarp();
                                if ( ok[k] )
                                {
                                        uint32_t const slot = pre + below;
                                        S.stage_seed[slot] = seed[k];
                                        S.stage_val[slot] = (uint32_t)(id << 2) | t;
                                        S.stage_b[slot] = (uint8_t)b;
                                }
                        }
                }
                __syncthreads();
                // (4) copy out run by run
                {
                        uint32_t const n = S.loc[EP_MAX_BUCKETS];
                        for ( uint32_t i = threadIdx.x; i < n; i += 256 )
                        {
                                uint32_t const b = S.stage_b[i];
                                uint32_t const o = S.base[b] + (i - S.loc[b]);
                                P.ent_seed[o] = S.stage_seed[i];
                                P.ent_val[o] = S.stage_val[i];
                        }
                }
                __syncthreads();
                #pragma unroll
                for ( int w = 0; w < 8; ++w ) S.wcnt[w][threadIdx.x] = 0;
                __syncthreads();
        }
}

// presence bits of the grouped entries (reductions into the L2-resident slice)
__global__ void __launch_bounds__(256) k_build_bits(const uint64_t * __restrict__ ent_seed, const uint32_t * __restrict__ ent_val, const uint32_t * __restrict__ total,
                                                  TableGeom G, SlotWord * __restrict__ slots)
{
        uint32_t const n = *total;
        for ( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
        {
                uint32_t const h = entry_slot(__ldcs(ent_seed + i), G, __ldcs(ent_val + i) & 3);
                atomicOr(&slots[h >> 5].bits, 1u << (h & 31));
        }
}

// E[rank(slot)] is claimed by the first entry that gets there; an entry that finds its slot taken is stored
// at E[ovf_base + i] (i = its index in the grouped array: no allocation counter, and still inside the
// bucket's address range) and linked in front of the head's chain.  A single global overflow counter was
// measured at ~4 ns per (same-address) atomic: 75 ms for the 19 M warps of a 600 M entry build.
//
// The claims of a bucket land all over its slice of E, and each one first pulls its (freshly memset) sector out
// of DRAM -- at the random-sector rate that was most of this kernel's time.  So while the grid sweeps bucket b it
// also streams bucket b+1's slice of E into L2 with line prefetches, one 128-byte line per 8 entries processed
// (the grid walks the grouped entries front to back, so "the entry at fraction f of bucket b" prefetches "the line
// at fraction f of bucket b+1's slice").  bucket_start = the partition's 257 bucket offsets; bucket_shift =
// hb - log2(buckets), 0 = no prefetching.
__global__ void __launch_bounds__(256) k_build_entries(const uint64_t * __restrict__ ent_seed, const uint32_t * __restrict__ ent_val, const uint32_t * __restrict__ total,
                                                     TableGeom G, const SlotWord * __restrict__ slots, uint32_t ovf_base, Entry * __restrict__ E,
                                                     const uint32_t * __restrict__ bucket_start, uint32_t nbuckets, uint32_t bucket_shift)
{
        __shared__ uint32_t bs[EP_MAX_BUCKETS + 1];       // first grouped entry of a bucket
        __shared__ uint32_t rs[EP_MAX_BUCKETS + 1];       // first rank (= index into E) of a bucket
        bool const pf = bucket_shift >= 5 && nbuckets > 1;
        if ( pf )
        {
                for ( uint32_t b = threadIdx.x; b <= nbuckets; b += blockDim.x )
                {
                        bs[b] = bucket_start[b];
                        rs[b] = (b < nbuckets) ? slots[((uint64_t)b << bucket_shift) >> 5].rank : 0xFFFFFFFFu;
                }
                __syncthreads();
        }
        uint32_t const n = *total;
        uint64_t const pol = policy_evict_last();
        for ( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
        {
                uint64_t const seed = __ldcs(ent_seed + i);
                uint32_t const val = __ldcs(ent_val + i);
                uint32_t const h = entry_slot(seed, G, val & 3);
                if ( pf && (threadIdx.x & 7) == 0 )
                {
                        uint32_t const b = h >> bucket_shift;
                        if ( b + 1 < nbuckets )
                        {
                                uint32_t const cnt = bs[b+1] - bs[b], f = i - bs[b];
                                uint32_t const r0 = rs[b+1], r1 = (b + 2 < nbuckets) ? rs[b+2] : r0 + (bs[b+2] - bs[b+1]);
                                uint32_t const lines = (r1 - r0 + 7) / 8 + 1;
                                uint32_t const line = (uint32_t)(((uint64_t)(f >> 3) * lines) / ((cnt >> 3) + 1));
                                if ( line < lines )
                                        asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(reinterpret_cast<const char *>(E + r0) + (uint64_t)line * 128));
                        }
                }
                SlotWord const sw = ld_hot_slotword(slots + (h >> 5), pol);
                uint32_t const rank = sw.rank + __popc(sw.bits & ((1u << (h & 31)) - 1));
                uint32_t const old = atomicCAS(&E[rank].val, ENTRY_NONE, val);
                if ( old == ENTRY_NONE )
                        E[rank].seed = seed;                    // .next stays ENTRY_NONE until somebody links behind it
                else
                {
                        uint32_t const idx = ovf_base + i;
                        Entry en; en.seed = seed; en.val = val;
                        en.next = atomicExch(&E[rank].next, idx);
                        E[idx] = en;
                }
        }
}

// Ranks: every block owns RANK_BLOCK_WORDS consecutive slot words (8 per thread).  First pass: presence bits per
// block; after an exclusive scan of the block sums the second pass writes the rank of every word's first slot.
__device__ __forceinline__ uint32_t load_block_words(const SlotWord * __restrict__ slots, SlotWord (&w)[8])
{
        const uint4 * p = reinterpret_cast<const uint4 *>(slots + ((uint64_t)blockIdx.x * RANK_BLOCK_WORDS + (uint64_t)threadIdx.x * 8));
        uint32_t c = 0;
        #pragma unroll
        for ( int i = 0; i < 4; ++i )
        {
                uint4 const v = p[i];
                w[2*i].bits = v.x; w[2*i].rank = v.y; w[2*i+1].bits = v.z; w[2*i+1].rank = v.w;
                c += __popc(v.x) + __popc(v.z);
        }
        return c;
}

__global__ void __launch_bounds__(256) k_word_sums(const SlotWord * __restrict__ slots, uint32_t * __restrict__ block_sums)
{
        SlotWord w[8];
        uint32_t const c = load_block_words(slots, w);
        uint32_t total;
        block_excl_scan(c, &total);
        if ( threadIdx.x == 0 ) block_sums[blockIdx.x] = total;
}

// the last block also publishes the number of distinct slots
__global__ void __launch_bounds__(256) k_word_ranks(SlotWord * __restrict__ slots, const uint32_t * __restrict__ block_offsets, uint32_t * __restrict__ ndistinct)
{
        SlotWord w[8];
        uint32_t const c = load_block_words(slots, w);
        uint32_t total;
        uint32_t run = block_excl_scan(c, &total) + block_offsets[blockIdx.x];
        if ( blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 ) *ndistinct = block_offsets[blockIdx.x] + total;
        uint4 * p = reinterpret_cast<uint4 *>(slots + ((uint64_t)blockIdx.x * RANK_BLOCK_WORDS + (uint64_t)threadIdx.x * 8));
        #pragma unroll
        for ( int i = 0; i < 4; ++i )
        {
                uint32_t const r0 = run, r1 = run + __popc(w[2*i].bits);
                run = r1 + __popc(w[2*i+1].bits);
                p[i] = make_uint4(w[2*i].bits, r0, w[2*i+1].bits, r1);
        }
}

} // namespace realgpu
