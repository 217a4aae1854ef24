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
// Each table is a presence bitmap over signature slots, cut into 32-byte sectors of
// {u32 rank of the first slot, 224 slot bits}, plus an entry array addressed by rank.
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

// one thread per (read, strand, word)
__global__ void __launch_bounds__(256) k_pack_reads(const uint8_t * __restrict__ mapped, const uint64_t * __restrict__ offsets,
                                                  uint64_t nreads, uint32_t W, uint64_t * __restrict__ rpack, uint32_t * __restrict__ bad)
{
        uint64_t const gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint64_t const total = nreads * 2 * W;
        if ( gid >= total ) return;
        uint32_t const w = (uint32_t)(gid % W);
        uint64_t const rs = gid / W;
        uint32_t const s = (uint32_t)(rs & 1);
        uint64_t const r = rs >> 1;
        uint64_t const o = offsets[r];
        uint32_t const L = (uint32_t)(offsets[r+1] - o);
        const uint8_t * p = mapped + o;
        uint64_t word = 0;
        uint32_t anybad = 0;
        uint32_t const j0 = w * 32;
        #pragma unroll 8
        for ( uint32_t j = 0; j < 32; ++j )
        {
                uint32_t const b = j0 + j;
                uint32_t sym = 0;
                if ( b < L )
                {
                        uint32_t const c = s ? p[L - 1 - b] : p[b];
                        anybad |= (c > 3);
                        sym = s ? (3 - (c & 3)) : (c & 3);
                }
                word = (word << 2) | sym;
        }
        rpack[gid] = word;
        if ( anybad ) bad[r] = 1;
}

// one thread per read: usable length and the two strand seeds
// seed word = seedl bases right aligned, fragment 0 in the top bits (what getTextWord(p,seedl) yields
// for the text window the strand is laid over)
__global__ void __launch_bounds__(256) k_read_seeds(const uint64_t * __restrict__ offsets, uint64_t nreads, uint32_t W, uint32_t seedl,
                                                  const uint64_t * __restrict__ rpack, const uint32_t * __restrict__ bad,
                                                  uint32_t * __restrict__ rlen, uint64_t * __restrict__ seeds, uint32_t * __restrict__ usable)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r >= nreads ) return;
        uint32_t const L = (uint32_t)(offsets[r+1] - offsets[r]);
        bool const ok = (L >= seedl) && !bad[r];
        rlen[r] = ok ? L : 0;
        usable[r] = ok ? 1 : 0;
        if ( ! ok ) { seeds[2*r] = 0; seeds[2*r+1] = 0; return; }
        // '+' : read[0..seedl)
        seeds[2*r] = rpack[(2*r) * W] >> (64 - 2*seedl);
        // '-' : the LAST seedl bases of the reverse complement strand (RestMatch.hpp:84-89)
        uint32_t const i = L - seedl;
        const uint64_t * rc = rpack + (2*r+1) * W;
        uint32_t const wi = i >> 5, sh = (i & 31) << 1;
        uint64_t const a = rc[wi];
        uint64_t const b = (wi + 1 < W) ? rc[wi+1] : 0;
        uint64_t const v = sh ? ((a << sh) | (b >> (64 - sh))) : a;
        seeds[2*r+1] = v >> (64 - 2*seedl);
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

// one thread per usable read strand: emits nlists (slot, val) pairs
__global__ void __launch_bounds__(256) k_gen_entries(const uint64_t * __restrict__ seeds, const uint32_t * __restrict__ usable_rank,
                                                   const uint32_t * __restrict__ usable, uint64_t nreads, TableGeom G,
                                                   uint32_t * __restrict__ keys, uint32_t * __restrict__ vals)
{
        uint64_t const id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // read*2 + strand
        if ( id >= 2 * nreads ) return;
        uint64_t const r = id >> 1;
        if ( ! usable[r] ) return;
        uint64_t const e0 = ((uint64_t)usable_rank[r] * 2 + (id & 1)) * G.nlists;
        uint64_t const seed = seeds[id];
        for ( uint32_t t = 0; t < G.nlists; ++t )
        {
                uint64_t const key = pair_key(seed, G.F, (int)t, pair_second(G.table, (int)t));
                keys[e0 + t] = slot_of(key, G.keybits, G.hb);
                vals[e0 + t] = (uint32_t)(id << 2) | t;
        }
}

__global__ void __launch_bounds__(256) k_mark_heads(const uint32_t * __restrict__ keys, uint32_t n, uint32_t * __restrict__ flags)
{
        uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        flags[i] = (i == 0 || keys[i] != keys[i-1]) ? 1u : 0u;
}

// heads go to E[rank], the other members of a slot group are chained behind them in E[ndistinct..n)
__global__ void __launch_bounds__(256) k_place_entries(const uint32_t * __restrict__ keys, const uint32_t * __restrict__ vals,
                                                     const uint32_t * __restrict__ headscan, uint32_t n, uint32_t ndistinct,
                                                     const uint64_t * __restrict__ seeds, Entry * __restrict__ E, uint32_t * __restrict__ bitmap)
{
        uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        uint32_t const key = keys[i];
        bool const head = (i == 0) || (keys[i-1] != key);
        uint32_t const g = headscan[i] + (head ? 1u : 0u) - 1u;         // rank of this slot group
        bool const more = (i + 1 < n) && (keys[i+1] == key);
        uint32_t const val = vals[i];
        Entry en;
        en.seed = seeds[val >> 2];
        en.val = val;
        en.next = more ? (ndistinct + (i - g)) : ENTRY_NONE;
        if ( head )
        {
                E[g] = en;
                uint32_t const sector = key / SECTOR_SLOTS, slot = key % SECTOR_SLOTS;
                atomicOr(&bitmap[(uint64_t)sector * SECTOR_WORDS + 1 + (slot >> 5)], 1u << (slot & 31));
        }
        else
                E[ndistinct + (i - g - 1)] = en;
}

__global__ void __launch_bounds__(256) k_sector_counts(const uint32_t * __restrict__ bitmap, uint32_t nsectors, uint32_t * __restrict__ counts)
{
        uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
        if ( s >= nsectors ) return;
        const uint4 * p = reinterpret_cast<const uint4 *>(bitmap + (uint64_t)s * SECTOR_WORDS);
        uint4 const a = p[0], b = p[1];
        counts[s] = __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
}

__global__ void __launch_bounds__(256) k_sector_headers(uint32_t * __restrict__ bitmap, uint32_t nsectors, const uint32_t * __restrict__ ranks)
{
        uint32_t const s = blockIdx.x * blockDim.x + threadIdx.x;
        if ( s >= nsectors ) return;
        bitmap[(uint64_t)s * SECTOR_WORDS] = ranks[s];
}

} // namespace realgpu
