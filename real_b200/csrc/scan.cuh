// K3: the text scan -- probe + verify + report, the dominant kernel of the path.
//
// Replaces, for every seed window of the text and every read at once, the reference's per-read
// ::match (match.hpp:335-416): directory lookup + equal_range on the signature, seed error count
// against the complementary signature (:386-388), position / record / wildcard predicates
// (:391-398, RangeVector.hpp:59-66, AutoTextArray.hpp:167-172), rest-of-read Hamming distance
// (RestMatch.hpp:39-81) and the updater call (:410).
//
// Shape: persistent CTAs walk tiles of 8192 text positions.  A tile of 2-bit text (+2 words of halo
// on either side) is staged in shared memory by one bulk asynchronous copy (cp.async.bulk, TMA unit)
// into a double buffer guarded by mbarriers while the previous tile is being probed.  Each thread
// owns one text word = 32 consecutive window starts, slides the 4-fragment window through registers
// with funnel shifts, and for every position issues one 4-byte load into each presence table
// (3 independent random sector reads per position, issued in batches of 8 positions for
// memory-level parallelism).  Only positions whose slot bit is set (a few %) take the second level:
// sector rank -> entry -> seed test -> canonical-list rule -> verification against the packed read.
#pragma once

#include "common.cuh"
#include "index.cuh"

namespace realgpu
{

static const int SC_THREADS = 256;
static const int SC_WPT = 2;                                 // text words per thread and tile
static const int SC_TILE_WORDS = SC_THREADS * SC_WPT;
static const int SC_TILE_POS = SC_TILE_WORDS * 32;           // 16384 window starts per tile
static const int SC_HALO = 2;                                // words of halo in front and behind
static const int SC_SMEM_WORDS = SC_TILE_WORDS + 2 * SC_HALO; // 516 words = 4128 bytes (multiple of 16)

struct TableDev
{
        const uint32_t * bitmap;
        const Entry * E;
        uint32_t hb;
        uint32_t nlists;
};

struct ScanParams
{
        const uint64_t * text;        // local word 0; valid from -TEXT_PAD_WORDS
        const uint64_t * nmask;       // local mask word 0
        uint64_t shard_begin;         // global position of local base 0
        uint64_t x_begin, x_end;      // local text positions whose keys are probed
        uint64_t win_begin, win_end;  // local seed-window starts this shard evaluates
        uint64_t own_begin, own_end;  // GLOBAL hit start positions this shard reports
        TableDev tab[3];
        uint32_t seedl, F, keybits, seedkmax, totalkmax;
        const uint64_t * rpack;
        uint32_t W;
        const uint32_t * rlen;
        const uint64_t * rec;         // nrec+1 global record starts
        uint32_t nrec;
        uint32_t fileid;
        uint32_t pass_bits, pass_id;  // this launch handles the windows whose first bases spell pass_id (2^pass_bits launches)
        int mode;                     // 0 = report all hits, 1 = fold into the unique state, 2 = gapped seed candidates
        RawHit * hits;
        unsigned long long hit_cap;
        unsigned long long * hit_count;
        unsigned long long * info;
        unsigned long long * stats;   // [0] candidates  [1] seed-pass  [2] hits
};

__device__ __forceinline__ uint32_t ld_probe(const uint32_t * p)
{
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
        uint32_t done = 0;
        while ( ! done )
        {
                asm volatile(
                        "{\n\t"
                        ".reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t"
                        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        }
}
// 1-D bulk asynchronous copy global -> shared, completion signalled on the mbarrier (TMA unit)
__device__ __forceinline__ void bulk_load(void * dst, const void * src, uint32_t bytes, uint64_t * bar)
{
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- reporting -------------------------------------------------------------------------------

// UpdateUniqueInfo<false>::update (matchUniqueImplementation.cpp:97-159) as an order-independent
// compare-and-swap: lowest error count wins, a second distinct position at that count makes the
// read NonUnique, '+' wins when both strands hit the same position (the reference probes the
// straight lists first).  The position kept in a NonUnique word is the smallest one seen, which
// makes the word deterministic (the reference keeps whichever came first).
__device__ __forceinline__ void unique_update(unsigned long long * slot, uint32_t inverted, uint32_t file, uint64_t pos, uint32_t k, uint32_t frag)
{
        unsigned long long old = *slot;
        while ( true )
        {
                uint32_t const st = umi_state(old);
                unsigned long long const take = umi_make(inverted ? ST_REVERSE : ST_STRAIGHT, file, pos, k, frag);
                unsigned long long neu = old;
                if ( st == ST_NOMATCH || st == ST_GAPPED )
                        neu = take;
                else if ( k < umi_err(old) )
                        neu = take;
                else if ( k == umi_err(old) )
                {
                        bool const same = (pos == umi_pos(old)) && (file == umi_file(old)) && (frag == umi_frag(old));
                        if ( st == ST_NONUNIQUE )
                        {
                                if ( file < umi_file(old) || (file == umi_file(old) && pos < umi_pos(old)) )
                                        neu = umi_with_state(take, ST_NONUNIQUE);
                        }
                        else if ( ! same )
                        {
                                bool const smaller = file < umi_file(old) || (file == umi_file(old) && pos < umi_pos(old));
                                neu = umi_with_state(smaller ? take : old, ST_NONUNIQUE);
                        }
                        else if ( ! inverted && st == ST_REVERSE )
                                neu = take;
                }
                if ( neu == old )
                        return;
                unsigned long long const prev = atomicCAS(slot, old, neu);
                if ( prev == old )
                        return;
                old = prev;
        }
}

// RangeVector::positionToRange (RangeVector.hpp:59-62): last record start <= pos
__device__ __forceinline__ uint32_t record_of(const uint64_t * __restrict__ rec, uint32_t nrec, uint64_t pos)
{
        uint32_t lo = 0, hi = nrec + 1;
        while ( lo < hi )
        {
                uint32_t const mid = (lo + hi) >> 1;
                if ( __ldg(rec + mid) <= pos ) lo = mid + 1; else hi = mid;
        }
        return lo - 1;
}

// AutoTextArray::isDontCareFree (AutoTextArray.hpp:167-172) on the raw mask bits
__device__ __forceinline__ bool wildcard_free(const uint64_t * __restrict__ nmask, uint64_t lpos, uint32_t len)
{
        uint64_t const first = lpos >> 6, last = (lpos + len - 1) >> 6;
        for ( uint64_t w = first; w <= last; ++w )
        {
                uint64_t m = __ldg(nmask + w);
                if ( m )
                {
                        if ( w == first ) m &= (~0ULL) >> (lpos & 63);
                        if ( w == last ) m &= (~0ULL) << (63 - ((lpos + len - 1) & 63));
                        if ( m ) return false;
                }
        }
        return true;
}

// one signature-equal entry at seed window lp (local): seed test, canonical-list rule, verification
__device__ __forceinline__ void examine_candidate(ScanParams const & P, int table, Entry const & en, uint64_t lx, unsigned long long * lstats)
{
        uint32_t const t = en.val & 3;
        uint32_t const id = en.val >> 2;
        if ( lx < (uint64_t)t * P.F ) return;
        uint64_t const lp = lx - (uint64_t)t * P.F;                      // seed window start (local)
        if ( lp < P.win_begin || lp >= P.win_end ) return;
        lstats[0] += 1;

        // seed errors: the whole seed of the read strand against the text window (match.hpp:386-388)
        uint64_t const win = text_word(P.text, lp, P.seedl);
        uint64_t x = en.seed ^ win;
        x = ((x >> 1) | x) & 0x5555555555555555ULL;
        uint32_t const seedk = (uint32_t)__popcll(x);
        if ( seedk > P.seedkmax ) return;
        // canonical list: the pair made of the two lowest exact fragments reports the match, so a
        // position reached through several lists is reported once (replaces unifyMatches' dedup)
        uint64_t const fm = (1ULL << (2*P.F)) - 1;
        int first = -1, second = -1;
        #pragma unroll
        for ( int f = 0; f < 4; ++f )
        {
                bool const exact = ((x >> (2*P.F*(3-f))) & fm) == 0;
                if ( exact ) { if ( first < 0 ) first = f; else if ( second < 0 ) second = f; }
        }
        if ( first != (int)t || second != pair_second(table, (int)t) ) return;
        lstats[1] += 1;

        uint32_t const strand = id & 1;
        uint32_t const read = id >> 1;
        uint32_t const L = __ldg(P.rlen + read);
        uint32_t const matchoffset = strand ? (L - P.seedl) : 0;          // RestMatch.hpp:84-89
        uint64_t const gp = P.shard_begin + lp;
        if ( gp < matchoffset ) return;                                   // match.hpp:393
        uint64_t const gpos = gp - matchoffset;
        if ( gpos < P.own_begin || gpos >= P.own_end ) return;
        uint64_t const lpos = gpos - P.shard_begin;

        // RangeVector::isPositionValid && AutoTextArray::isDontCareFree (match.hpp:398)
        uint32_t const frag = record_of(P.rec, P.nrec, gpos);
        if ( frag >= P.nrec || gpos + L > __ldg(P.rec + frag + 1) ) return;
        if ( ! wildcard_free(P.nmask, lpos, L) ) return;

        // whole-read Hamming distance = seedk + restk (match.hpp:400-405)
        const uint64_t * rp = P.rpack + (uint64_t)id * P.W;
        uint32_t k = 0;
        uint32_t const nw = (L + 31) >> 5;
        for ( uint32_t w = 0; w < nw; ++w )
        {
                uint32_t const len = (w + 1 == nw) ? (L - 32*w) : 32;
                uint64_t const rw = __ldg(rp + w) >> (64 - 2*len);
                k += diffcount64(rw, text_word(P.text, lpos + 32*w, len));
                if ( k > P.totalkmax ) return;
        }

        lstats[2] += 1;
        if ( P.mode == 1 )
                unique_update(P.info + read, strand, P.fileid, gpos, k, frag);
        else
        {
                unsigned long long const slot = atomicAdd(P.hit_count, 1ULL);
                if ( slot < P.hit_cap )
                {
                        RawHit h;
                        h.pm = rawhit_pack(gpos, k, strand, frag);
                        h.read = read;
                        h.score = 1.0f;
                        P.hits[slot] = h;
                }
        }
}

// second level of a probe whose slot bit is set: rank inside the sector, then the entry chain
__device__ __forceinline__ void follow_slot(ScanParams const & P, int table, uint32_t h, uint64_t lx, unsigned long long * lstats)
{
        uint32_t const sector = h / SECTOR_SLOTS, slot = h - sector * SECTOR_SLOTS;
        const uint4 * sp = reinterpret_cast<const uint4 *>(P.tab[table].bitmap + (uint64_t)sector * SECTOR_WORDS);
        uint4 const a = __ldg(sp), b = __ldg(sp + 1);
        uint32_t const wv[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
        uint32_t rank = wv[0];
        uint32_t const wi = slot >> 5;
        #pragma unroll
        for ( uint32_t w = 0; w < 7; ++w )
        {
                if ( w < wi ) rank += __popc(wv[1+w]);
                else if ( w == wi ) rank += __popc(wv[1+w] & ((1u << (slot & 31)) - 1));
        }
        uint32_t e = rank;
        while ( e != ENTRY_NONE )
        {
                uint4 const raw = __ldg(reinterpret_cast<const uint4 *>(P.tab[table].E + e));
                Entry en;
                en.seed = ((uint64_t)raw.y << 32) | raw.x;
                en.val = raw.z;
                en.next = raw.w;
                examine_candidate(P, table, en, lx, lstats);
                e = en.next;
        }
}

// positions (as bits 62-2j of a word, j = base index) whose base equals `code`; with `half` only the
// high bit of the base is compared
__device__ __forceinline__ uint64_t base_eq_mask(uint64_t w, uint32_t code, bool half)
{
        uint64_t const y = w ^ (0x5555555555555555ULL * code);
        uint64_t const z = half ? (y >> 1) : (y | (y >> 1));
        return ~z & 0x5555555555555555ULL;
}

// The scan is run as 2^pass_bits launches.  Launch `pass_id` handles the text positions whose window
// starts with the pass's base prefix, i.e. whose three table keys (they all start with fragment 0)
// carry pass_id in their top pass_bits bits -- so one launch only ever touches 1/2^pass_bits of every
// presence table, a slice that stays resident in L2 (measured on B200: random 32-byte sector reads run
// at ~290 G/s out of L2 against ~40 G/s out of HBM, tools/gather_bench.cu).
//
// Per tile of 16384 positions (TMA-staged, double buffered): (1) every thread derives, with a few
// 64-bit mask operations, which of the 64 positions of its two text words belong to the pass;
// (2) the selected positions are compacted into a shared-memory queue; (3) the queue is probed with
// all lanes busy: 3 independent 4-byte sector reads per position; (4) probes that found a set slot bit
// are compacted into a second queue and followed up (rank, entry chain, seed test, verification)
// in batches.
static const int SC_Q2_CAP = 2048;
static const int SC_Q2_DRAIN = 512;

struct ScanSmem
{
        uint64_t tile[2][SC_SMEM_WORDS];          // 2 x 4128 bytes, 16-byte aligned for the bulk copies
        unsigned long long q2[SC_Q2_CAP];
        uint64_t bar[2];
        uint16_t q1[SC_TILE_POS];
        uint32_t q1n, q2n;
};

__device__ __forceinline__ void drain_candidates(ScanParams const & P, unsigned long long * q2, uint32_t n2, unsigned long long * lstats)
{
        for ( uint32_t i = threadIdx.x; i < n2; i += SC_THREADS )
        {
                unsigned long long const it = q2[i];
                int const table = (int)(it & 3);
                uint64_t const lx = it >> 2;
                uint64_t const win = text_word(P.text, lx, P.seedl);
                uint32_t const F = P.F;
                uint64_t const fm = (1ULL << (2*F)) - 1;
                uint64_t const m0 = (win >> (6*F)) & fm;
                uint64_t const mo = (win >> (2*F*(2 - table))) & fm;
                uint32_t const h = slot_of((m0 << (2*F)) | mo, P.keybits, P.tab[table].hb);
                follow_slot(P, table, h, lx, lstats);
        }
}

__global__ void __launch_bounds__(SC_THREADS) k_text_scan(ScanParams P)
{
        extern __shared__ __align__(128) unsigned char sc_smem[];
        ScanSmem & S = *reinterpret_cast<ScanSmem *>(sc_smem);
        uint64_t (& tile)[2][SC_SMEM_WORDS] = S.tile;
        uint64_t (& bar)[2] = S.bar;
        uint16_t (& q1)[SC_TILE_POS] = S.q1;
        unsigned long long (& q2)[SC_Q2_CAP] = S.q2;
        uint32_t & q1n = S.q1n;
        uint32_t & q2n = S.q2n;

        uint64_t const first_tile = P.x_begin / SC_TILE_POS;
        uint64_t const end_tile = (P.x_end + SC_TILE_POS - 1) / SC_TILE_POS;
        unsigned long long lstats[3] = {0, 0, 0};
        int const lane = threadIdx.x & 31;

        if ( threadIdx.x == 0 )
        {
                mbar_init(&bar[0], 1);
                mbar_init(&bar[1], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                q1n = 0; q2n = 0;
        }
        __syncthreads();

        uint64_t tile_id = first_tile + blockIdx.x;
        if ( threadIdx.x == 0 && tile_id < end_tile )
        {
                mbar_expect_tx(&bar[0], SC_SMEM_WORDS * 8);
                bulk_load(&tile[0][0], P.text + (int64_t)tile_id * SC_TILE_WORDS - SC_HALO, SC_SMEM_WORDS * 8, &bar[0]);
        }

        uint32_t const F = P.F;
        uint32_t const kb = P.keybits;
        uint32_t const fsh = 64 - 2 * P.seedl;          // window is kept left aligned in 64 bits
        uint64_t const fm = (1ULL << (2*F)) - 1;
        bool const nlA = P.tab[0].nlists != 0, nlB = P.tab[1].nlists != 0, nlC = P.tab[2].nlists != 0;
        uint32_t const pbits = P.pass_bits, pid = P.pass_id;
        uint32_t const nb = (pbits + 1) >> 1;           // bases that decide the pass

        for ( uint32_t it = 0; tile_id < end_tile; tile_id += gridDim.x, ++it )
        {
                uint32_t const buf = it & 1;
                uint64_t const next_tile = tile_id + gridDim.x;
                if ( threadIdx.x == 0 && next_tile < end_tile )
                {
                        mbar_expect_tx(&bar[buf ^ 1], SC_SMEM_WORDS * 8);
                        bulk_load(&tile[buf ^ 1][0], P.text + (int64_t)next_tile * SC_TILE_WORDS - SC_HALO, SC_SMEM_WORDS * 8, &bar[buf ^ 1]);
                }
                mbar_wait(&bar[buf], (it >> 1) & 1);
                const uint64_t * tw = &tile[buf][SC_HALO];
                uint64_t const tile_x0 = tile_id * SC_TILE_POS;

                // (1)+(2) select the pass's positions and queue them
                #pragma unroll
                for ( int k = 0; k < SC_WPT; ++k )
                {
                        uint32_t const wi = threadIdx.x + k * SC_THREADS;
                        uint64_t const w0 = tw[wi], w1 = tw[wi + 1];
                        uint64_t sel = 0x5555555555555555ULL;
                        for ( uint32_t t = 0; t < nb; ++t )
                        {
                                bool const half = (2*(t+1) > pbits);
                                uint32_t const code = half ? ((pid & 1) << 1) : ((pid >> (pbits - 2*(t+1))) & 3);
                                uint64_t const e0 = base_eq_mask(w0, code, half), e1 = base_eq_mask(w1, code, half);
                                sel &= t ? ((e0 << (2*t)) | (e1 >> (64 - 2*t))) : e0;
                        }
                        // clip to [x_begin, x_end)
                        uint64_t const lx0 = tile_x0 + (uint64_t)wi * 32;
                        if ( lx0 + 32 <= P.x_begin || lx0 >= P.x_end ) sel = 0;
                        else
                        {
                                if ( lx0 < P.x_begin ) sel &= (~0ULL) >> (2 * (P.x_begin - lx0));
                                if ( lx0 + 32 > P.x_end ) sel &= (~0ULL) << (2 * (lx0 + 32 - P.x_end));
                        }
                        uint32_t const c = (uint32_t)__popcll(sel);
                        uint32_t const incl = warp_incl_scan(c, lane);
                        uint32_t base = 0;
                        if ( lane == 31 && incl ) base = atomicAdd(&q1n, incl);
                        base = __shfl_sync(0xffffffffu, base, 31);
                        uint32_t o = base + incl - c;
                        while ( sel )
                        {
                                uint32_t const b = 63 - __clzll(sel);
                                sel ^= 1ULL << b;
                                q1[o++] = (uint16_t)(wi * 32 + ((62 - b) >> 1));
                        }
                }
                __syncthreads();
                uint32_t const n1 = q1n;

                // (3) probe the queue, (4) queue the candidates
                for ( uint32_t i0 = 0; i0 < n1; i0 += SC_THREADS )
                {
                        uint32_t const i = i0 + threadIdx.x;
                        uint32_t cand = 0;
                        uint64_t lx = 0;
                        if ( i < n1 )
                        {
                                uint32_t const lp = q1[i];
                                uint32_t const wi = lp >> 5, j = lp & 31;
                                uint64_t const w0 = tw[wi], w1 = tw[wi + 1];
                                uint64_t const win = (j ? ((w0 << (2*j)) | (w1 >> (64 - 2*j))) : w0) >> fsh;
                                uint64_t const m0 = (win >> (6*F)) & fm, m1 = (win >> (4*F)) & fm, m2 = (win >> (2*F)) & fm, m3 = win & fm;
                                lx = tile_x0 + lp;
                                uint32_t vA = 0, vB = 0, vC = 0, rA = 0, rB = 0, rC = 0;
                                if ( nlA )
                                {
                                        uint32_t const h = slot_of((m0 << (2*F)) | m1, kb, P.tab[0].hb);
                                        uint32_t const sc = h / SECTOR_SLOTS; rA = h - sc * SECTOR_SLOTS;
                                        vA = ld_probe(P.tab[0].bitmap + (uint64_t)sc * SECTOR_WORDS + 1 + (rA >> 5));
                                }
                                if ( nlB )
                                {
                                        uint32_t const h = slot_of((m0 << (2*F)) | m2, kb, P.tab[1].hb);
                                        uint32_t const sc = h / SECTOR_SLOTS; rB = h - sc * SECTOR_SLOTS;
                                        vB = ld_probe(P.tab[1].bitmap + (uint64_t)sc * SECTOR_WORDS + 1 + (rB >> 5));
                                }
                                if ( nlC )
                                {
                                        uint32_t const h = slot_of((m0 << (2*F)) | m3, kb, P.tab[2].hb);
                                        uint32_t const sc = h / SECTOR_SLOTS; rC = h - sc * SECTOR_SLOTS;
                                        vC = ld_probe(P.tab[2].bitmap + (uint64_t)sc * SECTOR_WORDS + 1 + (rC >> 5));
                                }
                                cand = ((vA >> (rA & 31)) & 1) | (((vB >> (rB & 31)) & 1) << 1) | (((vC >> (rC & 31)) & 1) << 2);
                        }
                        uint32_t const c = __popc(cand);
                        uint32_t const incl = warp_incl_scan(c, lane);
                        uint32_t base = 0;
                        if ( lane == 31 && incl ) base = atomicAdd(&q2n, incl);
                        base = __shfl_sync(0xffffffffu, base, 31);
                        uint32_t o = base + incl - c;
                        #pragma unroll
                        for ( int table = 0; table < 3; ++table )
                                if ( (cand >> table) & 1 )
                                        q2[o++] = ((unsigned long long)lx << 2) | (unsigned long long)table;
                        __syncthreads();
                        uint32_t const n2 = q2n;
                        if ( n2 >= SC_Q2_DRAIN )
                        {
                                drain_candidates(P, q2, n2, lstats);
                                __syncthreads();
                                if ( threadIdx.x == 0 ) q2n = 0;
                                __syncthreads();
                        }
                }
                __syncthreads();   // everyone is done with tile[buf] and q1 before they are refilled
                if ( threadIdx.x == 0 ) q1n = 0;
                __syncthreads();
        }
        {
                uint32_t const n2 = q2n;
                if ( n2 ) drain_candidates(P, q2, n2, lstats);
        }

        // statistics: one atomic per warp and counter
        #pragma unroll
        for ( int s = 0; s < 3; ++s )
        {
                unsigned long long v = lstats[s];
                #pragma unroll
                for ( int o = 16; o > 0; o >>= 1 )
                        v += __shfl_xor_sync(0xffffffffu, v, o);
                if ( (threadIdx.x & 31) == 0 && v )
                        atomicAdd(P.stats + s, v);
        }
}

} // namespace realgpu
