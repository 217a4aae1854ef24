// K4 (quality-aware score), K6 (per-read ordering of matchAll hits) and the key transforms of the
// cross-shard unique reduction (K5).
#pragma once

#include "common.cuh"
#include "../../include/real_gpu.h"

namespace realgpu
{

// ---- K4 : ComputeScore<...,true>::computeScore (ComputeScore.hpp:50-190) ----------------------
// float(1.0 + sum_i LL[ref_i][read_i][q_i]), double accumulation in index order; the '-' strand is
// scored with the reverse complement and the qualities back to front (ComputeScore.hpp:79).
// One thread per hit: the additions are a dependent chain by definition of the result -- but nothing else is.  The bases
// and qualities of 32 positions are fetched together (the qualities as aligned 32-bit words: a byte load per position
// and thread made the kernel wait for one round trip per position, 4.7 ms for C4's 4 M hits), then the chain runs over
// registers.
// The 32 quality bytes of q[first .. first+32), byte j in out[j/4] bits 8*(j%4)..; only [first, first+len) is wanted,
// and only the aligned words that hold a wanted byte are read.
__device__ __forceinline__ void quality_run(const uint8_t * __restrict__ q, int64_t first, uint32_t want_from, uint32_t want_to, uint32_t (&out)[8])
{
        intptr_t const a = reinterpret_cast<intptr_t>(q) + first;
        const uint32_t * p = reinterpret_cast<const uint32_t *>(a & ~(intptr_t)3);
        uint32_t const mis = (uint32_t)(a & 3), sh = mis * 8;
        uint32_t w[9];
        #pragma unroll
        for ( int k = 0; k < 9; ++k )
        {
                // word k holds the bytes [4k - mis, 4k + 4 - mis) of the run
                int const b0 = 4 * k - (int)mis, b1 = b0 + 4;
                w[k] = (b1 > (int)want_from && b0 < (int)want_to) ? __ldg(p + k) : 0u;
        }
        #pragma unroll
        for ( int k = 0; k < 8; ++k ) out[k] = __funnelshift_r(w[k], w[k+1], sh);
}

// n = bases scored (the whole read, or its seed for the gapped pass), Lr = length of the read
__device__ __forceinline__ float score_hit(const double * __restrict__ sll, const uint64_t * __restrict__ text, uint64_t lpos,
                                           ReadSrc const & rs, uint32_t id, uint32_t Lr, const uint8_t * __restrict__ q, uint32_t L, uint32_t strand)
{
        double raw = 1.0;
        for ( uint32_t w = 0; w * 32 < L; ++w )
        {
                uint32_t const len = (L - 32*w < 32) ? (L - 32*w) : 32;
                uint64_t const tw = text_word(text, lpos + 32*w, len) << (64 - 2*len);   // left aligned
                uint64_t const rw = strand_bases(rs, id, Lr, 32*w, len);
                uint32_t qv[8];
                if ( q )
                {
                        if ( ! strand )
                                quality_run(q, (int64_t)32*w, 0, len, qv);
                        else
                        {
                                // position i of the strand carries quality L-1-i: the 32 bytes that END at L - 32w, back to front
                                uint32_t t[8];
                                quality_run(q, (int64_t)L - 32*(int64_t)w - 32, 32 - len, 32, t);
                                #pragma unroll
                                for ( int k = 0; k < 8; ++k ) qv[k] = __byte_perm(t[7-k], 0, 0x0123);
                        }
                }
                #pragma unroll
                for ( uint32_t j = 0; j < 32; ++j )
                        if ( j < len )
                        {
                                uint32_t const refb = (uint32_t)(tw >> (62 - 2*j)) & 3;
                                uint32_t const readb = (uint32_t)(rw >> (62 - 2*j)) & 3;
                                uint32_t const qq = q ? ((qv[j >> 2] >> (8 * (j & 3))) & 0xFFu) : 30u;
                                raw = __dadd_rn(raw, sll[((refb << 8) | (readb << 6) | qq) & 1023]);
                        }
        }
        return (float)raw;
}

__global__ void __launch_bounds__(256) k_score_hits(RawHit * __restrict__ hits, uint64_t nhits, const double * __restrict__ ll,
                                                  const uint64_t * __restrict__ text, uint64_t shard_begin,
                                                  ReadSrc rs, const uint32_t * __restrict__ rlen,
                                                  const uint8_t * __restrict__ quality, const uint64_t * __restrict__ offsets)
{
        __shared__ double sll[1024];
        for ( int i = threadIdx.x; i < 1024; i += blockDim.x ) sll[i] = ll[i];
        __syncthreads();
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= nhits ) return;
        RawHit h = hits[i];
        uint32_t const strand = rawhit_strand(h.pm);
        uint32_t const read = rawhit_read(h.read);
        uint32_t const L = rlen[read];
        const uint8_t * q = quality ? (quality + offsets[read]) : nullptr;
        h.score = score_hit(sll, text, rawhit_pos(h.pm) - shard_begin, rs, read * 2 + strand, L, q, L, strand);
        hits[i] = h;
}

// ---- K6 : unifyMatches order (matchAllImplementation.cpp:122-161) -----------------------------

__global__ void __launch_bounds__(256) k_hit_count(const RawHit * __restrict__ hits, uint64_t nhits, uint32_t * __restrict__ counts)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i < nhits ) atomicAdd(&counts[rawhit_read(hits[i].read)], 1u);
}

__global__ void __launch_bounds__(256) k_hit_scatter(const RawHit * __restrict__ hits, uint64_t nhits, const uint32_t * __restrict__ starts,
                                                   uint32_t * __restrict__ cursor, RawHit * __restrict__ out)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= nhits ) return;
        RawHit const h = hits[i];
        uint32_t const o = starts[rawhit_read(h.read)] + atomicAdd(&cursor[rawhit_read(h.read)], 1u);
        out[o] = h;
}

// (k, pos, file, frag, score, inverted); file is constant within one call and frag is a function of pos
__device__ __forceinline__ bool hit_before(RawHit const & a, RawHit const & b)
{
        uint32_t const ka = rawhit_k(a.pm), kb = rawhit_k(b.pm);
        if ( ka != kb ) return ka < kb;
        uint64_t const pa = rawhit_pos(a.pm), pb = rawhit_pos(b.pm);
        if ( pa != pb ) return pa < pb;
        if ( a.score != b.score ) return a.score < b.score;
        return rawhit_strand(a.pm) < rawhit_strand(b.pm);
}

// ---- segments too long for one thread --------------------------------------------------------------
// The per-read kernels below order a read's hits with an insertion sort run by one thread: fine for the handful of
// hits a read normally has, quadratic for a repeat-rich read with 10^4..10^5 of them (the reference uses std::sort in
// unifyMatches).  Segments of more than SEG_SMALL hits are therefore sorted beforehand, one CTA per segment, by a
// bottom-up merge sort (runs of SEG_SMALL by insertion, then merge passes in which every thread places one element by a
// binary search into the partner run); the scratch half of every pass is the same range of the raw hit buffer, which is
// free once the hits have been grouped by read.
static const uint32_t SEG_SMALL = 32;

__global__ void __launch_bounds__(256) k_mark_large(const uint32_t * __restrict__ counts, uint64_t nreads, uint32_t * __restrict__ list, uint32_t * __restrict__ nlarge)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r < nreads && counts[r] > SEG_SMALL ) list[atomicAdd(nlarge, 1u)] = (uint32_t)r;
}

template<typename Less>
__device__ __forceinline__ void seg_insertion_sort(RawHit * s, uint32_t n, Less less)
{
        for ( uint32_t i = 1; i < n; ++i )
        {
                RawHit const x = s[i];
                uint32_t j = i;
                while ( j > 0 && less(x, s[j-1]) ) { s[j] = s[j-1]; --j; }
                s[j] = x;
        }
}

// all threads of the CTA; s and tmp hold n elements each; the sorted segment ends up in s
template<typename Less>
__device__ void cta_merge_sort(RawHit * s, RawHit * tmp, uint32_t n, Less less)
{
        for ( uint32_t run = threadIdx.x; run * SEG_SMALL < n; run += blockDim.x )
                seg_insertion_sort(s + run * SEG_SMALL, min(SEG_SMALL, n - run * SEG_SMALL), less);
        __syncthreads();
        RawHit * src = s, * dst = tmp;
        for ( uint32_t w = SEG_SMALL; w < n; w <<= 1 )
        {
                for ( uint32_t i = threadIdx.x; i < n; i += blockDim.x )
                {
                        uint32_t const p0 = (i / (2 * w)) * (2 * w), mid = min(p0 + w, n), end = min(p0 + 2 * w, n);
                        RawHit const x = src[i];
                        uint32_t pos;
                        if ( i < mid )
                        {
                                uint32_t lo = mid, hi = end;            // elements of the right run that come before x
                                while ( lo < hi ) { uint32_t const m = (lo + hi) >> 1; if ( less(src[m], x) ) lo = m + 1; else hi = m; }
                                pos = i + (lo - mid);
                        }
                        else
                        {
                                uint32_t lo = p0, hi = mid;             // elements of the left run that do not come after x
                                while ( lo < hi ) { uint32_t const m = (lo + hi) >> 1; if ( ! less(x, src[m]) ) lo = m + 1; else hi = m; }
                                pos = lo + (i - mid);
                        }
                        dst[pos] = x;
                }
                __syncthreads();
                RawHit * t = src; src = dst; dst = t;
        }
        if ( src != s )
        {
                for ( uint32_t i = threadIdx.x; i < n; i += blockDim.x ) s[i] = src[i];
                __syncthreads();
        }
}

struct LessHit { __device__ bool operator()(RawHit const & a, RawHit const & b) const { return hit_before(a, b); } };
struct LessPos { __device__ bool operator()(RawHit const & a, RawHit const & b) const { return rawhit_pos(a.pm) < rawhit_pos(b.pm); } };

struct SortLargeParams
{
        RawHit * seg; RawHit * tmp;
        const uint32_t * starts; const uint32_t * counts;
        const uint32_t * list; const uint32_t * nlarge;
        // order-faithful replay: (text block, strand, seed window position)
        const uint32_t * rlen; uint32_t seedl;
        const uint64_t * bounds; uint32_t nblocks;
};

__device__ __forceinline__ uint32_t block_index(const uint64_t * __restrict__ bounds, uint32_t nblocks, uint64_t rpos)
{
        if ( ! bounds ) return 0;
        uint32_t lo = 0, hi = nblocks;
        while ( hi - lo > 1 )
        {
                uint32_t const mid = (lo + hi) >> 1;
                if ( bounds[mid] <= rpos ) lo = mid; else hi = mid;
        }
        return lo;
}

struct LessReplay
{
        const uint64_t * bounds; uint32_t nblocks, moff;
        __device__ bool operator()(RawHit const & a, RawHit const & b) const
        {
                uint32_t const sa = rawhit_strand(a.pm), sb = rawhit_strand(b.pm);
                uint64_t const ra = rawhit_pos(a.pm) + (sa ? moff : 0), rb = rawhit_pos(b.pm) + (sb ? moff : 0);
                uint32_t const ba = block_index(bounds, nblocks, ra), bb = block_index(bounds, nblocks, rb);
                if ( ba != bb ) return ba < bb;
                if ( sa != sb ) return sa < sb;
                return ra < rb;
        }
};

// MODE 0: unifyMatches order, 1: replay order of matchUnique with scores, 2: ascending position (gapped pass)
template<int MODE>
__global__ void __launch_bounds__(256) k_sort_large(SortLargeParams P)
{
        uint32_t const nl = *P.nlarge;
        for ( uint32_t li = blockIdx.x; li < nl; li += gridDim.x )
        {
                uint32_t const r = P.list[li];
                uint32_t const o = P.starts[r], n = P.counts[r];
                if ( MODE == 0 ) cta_merge_sort(P.seg + o, P.tmp + o, n, LessHit());
                else if ( MODE == 1 )
                {
                        LessReplay L; L.bounds = P.bounds; L.nblocks = P.nblocks; L.moff = P.rlen[r] - P.seedl;
                        cta_merge_sort(P.seg + o, P.tmp + o, n, L);
                }
                else cta_merge_sort(P.seg + o, P.tmp + o, n, LessPos());
                __syncthreads();
        }
}

// one thread per read: insertion sort of its (short) segment, then expansion to the ABI record
template<typename HitOut>
__global__ void __launch_bounds__(128) k_hit_order(RawHit * __restrict__ seg, const uint32_t * __restrict__ starts, const uint32_t * __restrict__ counts,
                                                 uint64_t nreads, uint32_t fileid, HitOut * __restrict__ out, real_gpu_hit16 * __restrict__ out16)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r >= nreads ) return;
        uint32_t const n = counts[r];
        if ( ! n ) return;
        RawHit * s = seg + starts[r];
        if ( n <= SEG_SMALL ) seg_insertion_sort(s, n, LessHit());          // longer segments: k_sort_large<0> has run
        HitOut * o = out + starts[r];
        real_gpu_hit16 * o16 = out16 ? out16 + starts[r] : nullptr;
        for ( uint32_t i = 0; i < n; ++i )
        {
                if ( o16 )
                {
                        // the compact row: position, error count, strand and record in one word (the layout of RawHit::pm)
                        real_gpu_hit16 C;
                        C.pos_k_inv_frag = s[i].pm; C.patid = (uint32_t)r; C.score = s[i].score;
                        o16[i] = C;
                }
                HitOut H;
                H.patid = r;
                H.pos = rawhit_pos(s[i].pm);
                H.file = fileid;
                H.frag = rawhit_frag(s[i].pm);
                H.k = rawhit_k(s[i].pm);
                H.inverted = rawhit_strand(s[i].pm);
                H.score = s[i].score;
                H.reserved = 0;
                o[i] = H;
        }
}

// ---- order-faithful matchUnique with scores ---------------------------------------------------
// UpdateUniqueInfo<true>::update (matchUniqueImplementation.cpp:179-248) compares float scores with a
// tolerance epsilon = (float)(filter_mult * patl); the comparisons are not transitive, so the result
// depends on the order in which the reference visits the hits of a read: text block, strand ('+' first),
// list 0..5, ascending seed-window position.  The scan reports every hit once together with
// the set of exact seed fragments; the lists that see a hit are the fragment pairs that are both exact.

// valid seed-window starts of one 64-position group: no wildcard in [i, i+seedl) and i + seedl <= n
__global__ void __launch_bounds__(256) k_window_counts(const uint64_t * __restrict__ nmask, uint64_t n, uint32_t seedl, uint64_t ngroups,
                                                     uint64_t * __restrict__ valid, uint32_t * __restrict__ counts)
{
        uint64_t const g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( g >= ngroups ) return;
        // bad(i) = OR of the wildcard bits i .. i+seedl-1 ; positions are MSB first, so "later" is toward the low bits
        uint64_t hi = nmask[g], lo = nmask[g + 1];
        uint32_t covered = 1;
        while ( covered < seedl )
        {
                uint32_t const step = min(covered, seedl - covered);
                uint64_t const nhi = hi | ((hi << step) | (lo >> (64 - step)));
                uint64_t const nlo = lo | (lo << step);
                hi = nhi; lo = nlo;
                covered += step;
        }
        uint64_t v = ~hi;
        uint64_t const first = g * 64;
        uint64_t const last_start = (n >= seedl) ? (n - seedl) : 0;              // inclusive
        if ( n < seedl || first > last_start ) v = 0;
        else if ( first + 63 > last_start ) v &= (~0ULL) << (63 - (last_start - first));
        valid[g] = v;
        counts[g] = (uint32_t)__popcll(v);
}

// position of the first window of block b (b >= 1): the (b * n_list)-th valid window start
__global__ void __launch_bounds__(64) k_block_bounds(const uint64_t * __restrict__ valid, const uint32_t * __restrict__ prefix, uint64_t ngroups,
                                                   uint64_t n_list, uint32_t nblocks, uint64_t * __restrict__ bounds)
{
        uint32_t const b = blockIdx.x * blockDim.x + threadIdx.x;
        if ( b >= nblocks ) return;
        if ( b == 0 ) { bounds[0] = 0; return; }
        uint64_t const target = (uint64_t)b * n_list;
        uint64_t lo = 0, hi = ngroups;           // last group with prefix[g] <= target
        while ( hi - lo > 1 )
        {
                uint64_t const mid = (lo + hi) >> 1;
                if ( prefix[mid] <= target ) lo = mid; else hi = mid;
        }
        uint64_t v = valid[lo];
        uint32_t skip = (uint32_t)(target - prefix[lo]);
        while ( skip-- ) v &= ~(1ULL << (63 - __clzll(v)));
        bounds[b] = lo * 64 + __clzll(v);
}

struct ReplayParams
{
        RawHit * seg;                   // hits grouped by read
        const uint32_t * starts;
        const uint32_t * counts;
        const uint32_t * rlen;
        uint64_t nreads;
        uint32_t seedl, fileid;
        double filter_mult;
        const uint64_t * bounds;        // first window of every block, or null for a single block
        uint32_t nblocks;
        unsigned long long * info;
        float * score;
};

__device__ __forceinline__ uint32_t block_of(ReplayParams const & P, uint64_t rpos)
{
        if ( ! P.bounds ) return 0;
        uint32_t lo = 0, hi = P.nblocks;
        while ( hi - lo > 1 )
        {
                uint32_t const mid = (lo + hi) >> 1;
                if ( P.bounds[mid] <= rpos ) lo = mid; else hi = mid;
        }
        return lo;
}

// one thread per read
__global__ void __launch_bounds__(128) k_unique_replay(ReplayParams P)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r >= P.nreads ) return;
        uint32_t const n = P.counts[r];
        if ( ! n ) return;
        RawHit * s = P.seg + P.starts[r];
        uint32_t const L = P.rlen[r];
        float const epsilon = (float)(P.filter_mult * (double)L);
        uint32_t const moff = L - P.seedl;
        // order by (block, strand, seed window position); longer segments: k_sort_large<1> has run
        if ( n <= SEG_SMALL )
        {
                LessReplay LR; LR.bounds = P.bounds; LR.nblocks = P.nblocks; LR.moff = moff;
                seg_insertion_sort(s, n, LR);
        }
        unsigned long long d = P.info[r];
        float sc = P.score[r];
        uint32_t i = 0;
        while ( i < n )
        {
                uint32_t const st = rawhit_strand(s[i].pm);
                uint32_t const bl = block_of(P, rawhit_pos(s[i].pm) + (st ? moff : 0));
                uint32_t j = i + 1;
                while ( j < n && rawhit_strand(s[j].pm) == st && block_of(P, rawhit_pos(s[j].pm) + (st ? moff : 0)) == bl ) ++j;
                for ( uint32_t l = 0; l < 6; ++l )
                {
                        // list l = fragment pair (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
                        uint32_t const fa = l < 3 ? 0u : (l < 5 ? 1u : 2u);
                        uint32_t const fb = l < 3 ? l + 1 : (l < 5 ? l - 1 : 3u);
                        uint32_t const need = (1u << fa) | (1u << fb);
                        for ( uint32_t t = i; t < j; ++t )
                        {
                                if ( (rawhit_exact(s[t].read) & need) != need ) continue;
                                uint64_t const pos = rawhit_pos(s[t].pm);
                                uint32_t const k = rawhit_k(s[t].pm), frag = rawhit_frag(s[t].pm);
                                float const score = s[t].score;
                                uint32_t const state = umi_state(d);
                                bool const same = (pos == umi_pos(d)) && (P.fileid == umi_file(d)) && (frag == umi_frag(d));
                                if ( state == ST_NOMATCH || state == ST_GAPPED )
                                {
                                        d = umi_make(st ? ST_REVERSE : ST_STRAIGHT, P.fileid, pos, k, frag); sc = score;
                                }
                                else if ( score > sc + epsilon )
                                {
                                        d = umi_make(st ? ST_REVERSE : ST_STRAIGHT, P.fileid, pos, k, frag); sc = score;
                                }
                                else if ( state != ST_NONUNIQUE && (score > sc - epsilon) && ! same )
                                        d = umi_with_state(d, ST_NONUNIQUE);
                        }
                }
                i = j;
        }
        P.info[r] = d;
        P.score[r] = sc;
}

// ---- gapped extension (K7): ::matchGaps, match.hpp:428-602 --------------------------------------
// Per seed candidate of a read that is still NoMatch/Gapped: banded DP `agm` (match.hpp:267-332) of the
// read rest (m = L - seedl bases) against up to n = min(record end - seedl - rpos, 2L) text bases behind the
// seed, band |i-j| <= 3, fp64 scores from the LL table:
//     G[i][j] = max( G[i-1][j-1] + LL[text_i][read_j][q_j] ,  G[d][d] with d = min(i,j) )      (off the main diagonal)
//     H[i][j] = |i-j| when the second term is strictly larger, else 0
// then `opt_solution` (:159-260) over the end cells of the seven diagonals and `backtracing` (:115-152) along the
// winning diagonal.  Only seven running diagonal values, the last four main-diagonal values and the row of the
// last gap per diagonal are needed; every addition happens in the reference's order, so the doubles are identical.
struct GapRes { double complete; uint32_t mingap, where, start, gap_pos; };

struct GapParams
{
        const RawHit * seg; uint64_t ncand;
        const double * ll;
        const uint64_t * text; const uint64_t * nmask; uint64_t shard_begin;
        const uint64_t * rec; uint32_t nrec;
        ReadSrc rs; const uint32_t * rlen;
        const uint8_t * quality; const uint64_t * offsets;
        uint32_t seedl, scores;
        GapRes * res;
};

__device__ __forceinline__ double gap_total_scoring(uint32_t gap, double cur)
{
        // total_scoring(gap, cur, open = -1, extend = -1, offset = -1), match.hpp:16-29
        if ( gap % 3 == 0 ) return cur + ((double)gap * -1.0) + -1.0 + -1.0;
        return cur + ((double)gap * -1.0) + -1.0;
}

__global__ void __launch_bounds__(128) k_gap_dp(GapParams P)
{
        __shared__ double sll[1024];
        for ( int i = threadIdx.x; i < 1024; i += blockDim.x ) sll[i] = P.ll[i];
        __syncthreads();
        uint64_t const c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( c >= P.ncand ) return;
        RawHit const h = P.seg[c];
        uint32_t const read = rawhit_read(h.read);
        uint64_t const rpos = rawhit_pos(h.pm);
        uint32_t const frag = rawhit_frag(h.pm);
        uint32_t const L = P.rlen[read];
        uint32_t const seedl = P.seedl;
        GapRes R; R.complete = 0; R.mingap = 0; R.where = 0; R.start = 0; R.gap_pos = 0;
        uint64_t const phigh = P.rec[frag + 1];
        uint64_t n64 = phigh - seedl - rpos;
        if ( n64 > 2ULL * L ) n64 = 2ULL * L;
        uint32_t const n = (uint32_t)n64, m = L - seedl;
        uint64_t const lbase = rpos - P.shard_begin;
        if ( ! (n && m && wildcard_free(P.nmask, lbase + seedl, n)) ) { P.res[c] = R; return; }

        // '+' strand: packed words, or the caller's 2 bit/base bytes (4 bases per byte, first base in bits 7..6)
        const uint64_t * rp = P.rs.rpack ? P.rs.rpack + (uint64_t)read * 2 * P.rs.W : nullptr;
        const uint8_t * pp = P.rs.rpack ? nullptr : packed_read(P.rs, read);
        const uint8_t * q = P.quality ? (P.quality + P.offsets[read]) : nullptr;
        double const seedscore = P.scores ? (double)score_hit(sll, P.text, lbase, P.rs, read * 2, L, q, seedl, 0) : (double)1.0f;

        int const MAXgap = 3;
        double const MINscore = -100.0;
        double g[7];                 // running value of diagonal delta = i - j, index delta + 3
        uint32_t lastgap[7];         // row of the last cell of the diagonal that took the gap branch (0 = none)
        #pragma unroll
        for ( int d = 0; d < 7; ++d ) { g[d] = 0.0; lastgap[d] = 0; }
        double mainh[4] = {0.0, 0.0, 0.0, 0.0};    // G[i-1][i-1], G[i-2][i-2], G[i-3][i-3] at [1..3]; [0] = G[i][i]

        // The substitution term of cell (i, j) depends on the text base of row i and on (read base, quality) of column j; a
        // row touches the columns i-3 .. i+3, so the columns slide through a seven-entry window and every row fetches ONE
        // new column (win[k] = (base << 6 | quality) of column i - 3 + k) -- the cells used to fetch their column each,
        // fourteen loads per row, and the kernel waited on them (L1TEX 78 % busy, 5.3 ms for C4's candidates).
        auto column = [&](uint32_t j) -> uint32_t
        {
                uint32_t const rpos_read = j + seedl - 1;
                uint32_t const rb = rp ? ((uint32_t)(__ldg(rp + (rpos_read >> 5)) >> (62 - 2 * (rpos_read & 31))) & 3)
                                       : (((uint32_t)__ldg(pp + (rpos_read >> 2)) >> (6 - 2 * (rpos_read & 3))) & 3);
                uint32_t const qq = q ? (uint32_t)__ldg(q + rpos_read) : 30u;
                return (rb << 6) | qq;
        };
        uint32_t win[7] = {0, 0, 0, 0, 0, 0, 0};
        #pragma unroll
        for ( uint32_t j = 1; j <= 3; ++j ) if ( j <= m ) win[3 + j] = column(j);      // becomes win[2 + j] of row 1
        uint64_t tword = 0;
        for ( uint32_t i = 1; i <= n; ++i )
        {
                int const left = ((int)i - MAXgap > 0) ? ((int)i - MAXgap) : 1;
                int const right = (i + MAXgap > m) ? (int)m : (int)(i + MAXgap);
                if ( left > right ) break;                                   // the band has left the matrix: nothing below is ever read
                #pragma unroll
                for ( int k = 0; k < 6; ++k ) win[k] = win[k+1];
                win[6] = (i + 3 <= m) ? column(i + 3) : 0u;
                uint64_t const tpos = lbase + seedl - 1 + i;
                if ( i == 1 || (tpos & 31) == 0 ) tword = __ldg(P.text + (tpos >> 5));
                uint32_t const tb = (uint32_t)(tword >> (62 - 2 * (tpos & 31))) & 3;
                // main diagonal history: shift before the row is computed
                mainh[3] = mainh[2]; mainh[2] = mainh[1]; mainh[1] = mainh[0];
                #pragma unroll
                for ( int dd = 3; dd >= -3; --dd )                            // j = i - dd ascending
                {
                        int const j = (int)i - dd;
                        if ( j < left || j > right ) continue;
                        double const sub = sll[((tb << 8) | win[3 - dd]) & 1023];
                        double const mis = __dadd_rn(g[dd + 3], sub);
                        if ( dd == 0 )
                        {
                                g[3] = mis;
                                mainh[0] = mis;
                        }
                        else
                        {
                                // G[d][d], d = min(i,j): for j < i the main diagonal value of dd rows ago, for j > i this row's
                                double const gp = (dd > 0) ? mainh[dd] : mainh[0];
                                g[dd + 3] = (mis < gp) ? gp : mis;
                                if ( gp > mis ) lastgap[dd + 3] = i;
                        }
                }
        }

        // opt_solution: end cells (i, m) for i in [up, down], then (n, j) for j in [left, right)
        double score = MINscore;
        int const up = ((int)m - MAXgap < 0) ? 0 : ((int)m - MAXgap);
        int const down = (m + MAXgap > n) ? (int)n : (int)(m + MAXgap);
        double maxscore = 0;
        for ( int i = up; i <= down; ++i )
        {
                int const dd = i - (int)m;
                double const gv = (i >= 1) ? g[dd + 3] : 0.0;               // row 0 is never written: calloc'ed zero
                if ( gv >= MINscore )
                {
                        uint32_t const gap = (uint32_t)(dd < 0 ? -dd : dd);
                        double const t = gap_total_scoring(gap, gv);
                        if ( t > score )
                        {
                                score = t; maxscore = t; R.mingap = gap;
                                R.where = dd < 0 ? 1u : (dd > 0 ? 2u : 0u);
                                R.start = dd == 0 ? m : (uint32_t)i;
                        }
                }
        }
        if ( m + MAXgap > n )
        {
                int const left = ((int)n - MAXgap > 0) ? ((int)n - MAXgap) : 1;
                int const right = (n + MAXgap > m) ? (int)m : (int)(n + MAXgap);
                for ( int j = left; j < right; ++j )
                {
                        int const dd = (int)n - j;
                        double const gv = g[dd + 3];
                        if ( gv >= MINscore && dd <= MAXgap )
                        {
                                double const t = gap_total_scoring((uint32_t)dd, gv);
                                if ( t > score ) { score = t; maxscore = t; R.mingap = (uint32_t)dd; R.where = 3; R.start = (uint32_t)j; }
                        }
                }
        }
        // backtracing along the winning diagonal: first cell from the end whose H is set
        {
                int bi, bj;
                if ( R.where == 1 || R.where == 2 ) { bi = (int)R.start; bj = (int)m; }
                else { bi = (int)n; bj = (int)R.start; }
                int const dd = bi - bj;
                R.gap_pos = 0;
                if ( dd >= -3 && dd <= 3 && dd != 0 )
                {
                        uint32_t const lg = lastgap[dd + 3];
                        if ( lg && (int)lg <= bi )
                        {
                                int const gi = (int)lg, gj = gi - dd;
                                R.gap_pos = (uint32_t)(gi > gj ? gj : gi);
                        }
                        // else the walk ends on the matrix border (H[i][0] = i, H[0][j] = j): min(i,j) = 0
                }
        }
        R.complete = seedscore + maxscore;
        P.res[c] = R;
}

struct GapReplayParams
{
        RawHit * seg; GapRes * res;
        const uint32_t * starts; const uint32_t * counts;
        uint64_t nreads;
        uint32_t scores;
        const uint64_t * bounds; uint32_t nblocks;
        unsigned long long * info; float * score; real_gpu_gapinfo * gaps;
};

// one thread per read: the reference's visiting order (block, list 0..5, window position), '+' strand only
__global__ void __launch_bounds__(128) k_gap_replay(GapReplayParams P)
{
        uint64_t const r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( r >= P.nreads ) return;
        uint32_t const n = P.counts[r];
        if ( ! n ) return;
        RawHit * s = P.seg + P.starts[r];
        GapRes * gr = P.res + P.starts[r];
        auto blk = [&](uint64_t rpos) -> uint32_t
        {
                if ( ! P.bounds ) return 0;
                uint32_t lo = 0, hi = P.nblocks;
                while ( hi - lo > 1 ) { uint32_t const mid = (lo + hi) >> 1; if ( P.bounds[mid] <= rpos ) lo = mid; else hi = mid; }
                return lo;
        };
        // ascending window position (which also orders the blocks); longer segments were sorted before k_gap_dp ran
        // (k_sort_large<2>), so their results already stand in that order
        for ( uint32_t i = 1; i < n && n <= SEG_SMALL; ++i )
        {
                RawHit const x = s[i]; GapRes const y = gr[i];
                uint32_t j = i;
                while ( j > 0 && rawhit_pos(x.pm) < rawhit_pos(s[j-1].pm) ) { s[j] = s[j-1]; gr[j] = gr[j-1]; --j; }
                s[j] = x; gr[j] = y;
        }
        unsigned long long d = P.info[r];
        float sc = P.scores ? P.score[r] : 0.0f;
        real_gpu_gapinfo gi = P.gaps[r];
        uint32_t i = 0;
        while ( i < n )
        {
                uint32_t const bl = blk(rawhit_pos(s[i].pm));
                uint32_t j = i + 1;
                while ( j < n && blk(rawhit_pos(s[j].pm)) == bl ) ++j;
                for ( uint32_t l = 0; l < 6; ++l )
                {
                        uint32_t const fa = l < 3 ? 0u : (l < 5 ? 1u : 2u);
                        uint32_t const fb = l < 3 ? l + 1 : (l < 5 ? l - 1 : 3u);
                        uint32_t const need = (1u << fa) | (1u << fb);
                        for ( uint32_t t = i; t < j; ++t )
                        {
                                if ( (rawhit_exact(s[t].read) & need) != need ) continue;
                                GapRes const & G = gr[t];
                                if ( ! G.mingap ) continue;                                  // match.hpp:540
                                uint64_t const rpos = rawhit_pos(s[t].pm);
                                float const stored = P.scores ? sc : 0.0f;                  // UniqueMatchInfo.hpp:178-185
                                uint32_t const st = umi_state(d);
                                bool take = false;
                                if ( st == ST_NOMATCH )
                                {
                                        d = umi_with_state(d, ST_GAPPED);
                                        take = true;
                                }
                                else if ( st == ST_GAPPED )
                                {
                                        if ( G.complete > (double)stored + 1e-6 ) take = true;
                                        else if ( G.complete < (double)stored - 1e-6 ) {}
                                        else gi.present = 0;
                                }
                                if ( take )
                                {
                                        if ( P.scores ) sc = (float)G.complete;
                                        d = (d & ~UMI_POSMASK) | rpos;
                                        gi.patid = (uint32_t)r; gi.mingap = G.mingap; gi.where = G.where; gi.start = G.start; gi.gap_pos = G.gap_pos; gi.present = 1;
                                }
                        }
                }
                i = j;
        }
        P.info[r] = d;
        if ( P.scores ) P.score[r] = sc;
        P.gaps[r] = gi;
}

__global__ void __launch_bounds__(256) k_fill_f32(float * __restrict__ p, uint64_t n, float v)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i < n ) p[i] = v;
}

// ---- K5 : cross-shard reduction keys --------------------------------------------
// key, most significant first: err(4) | unique flag (0 = NonUnique, sorts first) | file(6) | pos(35) | strand(1) | frag(16)
// (63 bits, so the key orders the same as u64 and as i64 -- torch/NCCL reduce it as int64).
// "no hit" = UNIQUE_KEY_NONE = 2^63-1.  MIN over shards picks the lowest error count, prefers an already-NonUnique
// shard, then the smallest (file,pos), then '+'.
#define UNIQUE_KEY_NONE 0x7FFFFFFFFFFFFFFFULL
__device__ __forceinline__ uint64_t unique_key(uint64_t d)
{
        uint32_t const st = umi_state(d);
        if ( st == ST_NOMATCH || st == ST_GAPPED ) return UNIQUE_KEY_NONE;
        uint64_t const uniq = (st == ST_NONUNIQUE) ? 0 : 1;
        uint64_t const strand = (st == ST_REVERSE) ? 1 : 0;
        return ((uint64_t)umi_err(d) << 59) | (uniq << 58) | ((uint64_t)umi_file(d) << 52) | (umi_pos(d) << 17) | (strand << 16) | umi_frag(d);
}

__global__ void __launch_bounds__(256) k_unique_export(const unsigned long long * __restrict__ info, uint64_t n, uint64_t * __restrict__ keys)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i < n ) keys[i] = unique_key(info[i]);
}

__global__ void __launch_bounds__(256) k_unique_ties(const unsigned long long * __restrict__ info, uint64_t n, const uint64_t * __restrict__ minkeys, uint8_t * __restrict__ ties)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        uint64_t const mine = unique_key(info[i]);
        uint64_t const win = minkeys[i];
        // a DIFFERENT (file,pos) at the winning error count; shards that merely carry the same merged
        // state from an earlier file do not count
        bool const samepos = (((mine ^ win) >> 17) & ((1ULL << 41) - 1)) == 0;
        ties[i] = (mine != UNIQUE_KEY_NONE && (mine >> 59) == (win >> 59) && ! samepos) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_unique_import(unsigned long long * __restrict__ info, uint64_t n, const uint64_t * __restrict__ minkeys, const uint8_t * __restrict__ tiesum)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        uint64_t const key = minkeys[i];
        if ( key == UNIQUE_KEY_NONE )
                return;    // no shard has a hit: keep the local NoMatch/Gapped word
        uint32_t const err = (uint32_t)(key >> 59);
        bool const uniq = ((key >> 58) & 1) && (tiesum[i] == 0);
        uint32_t const file = (uint32_t)((key >> 52) & 63);
        uint64_t const pos = (key >> 17) & UMI_POSMASK;
        uint32_t const strand = (uint32_t)((key >> 16) & 1);
        uint32_t const frag = (uint32_t)(key & 0xFFFF);
        info[i] = umi_make(uniq ? (strand ? ST_REVERSE : ST_STRAIGHT) : ST_NONUNIQUE, file, pos, err, frag);
}

// ---- K5 over peer memory: the fold as ONE exchange (real_gpu_fold_*) ---------------------------------
// Reduce-scatter form of the reduction above, without a collective library: rank r owns the reads
// [R r / N, R (r+1) / N).  k_fold_push stores this rank's state words of every owner's reads straight into the owner's
// staging area (peer memory over NVLink; 8 bytes per read and peer, plain coalesced stores), the ranks hand over with
// the release/acquire flags of scan.cuh, and k_fold_merge folds the N words of each own read -- lowest error count,
// NonUnique when a second distinct position holds it, smallest (file, position), '+' before '-': the rule of
// unique_key / k_unique_ties / k_unique_import in one pass.  UpdateUniqueInfo<false>::update,
// matchUniqueImplementation.cpp:97-159, folded over disjoint sets of hits.
struct FoldPeers { unsigned long long * stage[8]; };       // staging area of every rank for this exchange (own included)

// grid.y = destination rank
__global__ void __launch_bounds__(256) k_fold_push(const unsigned long long * __restrict__ info, uint64_t nreads, uint32_t nranks, uint32_t rank,
                                                 uint64_t seg, FoldPeers peers)
{
        uint32_t const j = blockIdx.y;
        uint64_t const b = (nreads * j) / nranks, e = (nreads * (j + 1)) / nranks;
        unsigned long long * dst = peers.stage[j] + (uint64_t)rank * seg;
        for ( uint64_t i = b + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (uint64_t)gridDim.x * blockDim.x )
                dst[i - b] = info[i];
}

// stage = this rank's staging area: the words of source s at stage + s * seg; info_own = &info[first own read]
__global__ void __launch_bounds__(256) k_fold_merge(const unsigned long long * __restrict__ stage, uint32_t nranks, uint64_t seg, uint64_t n,
                                                  unsigned long long * __restrict__ info_own)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        uint64_t key[8];
        uint64_t win = UNIQUE_KEY_NONE;
        #pragma unroll
        for ( uint32_t s = 0; s < 8; ++s )
        {
                key[s] = (s < nranks) ? unique_key(__ldcs(stage + (uint64_t)s * seg + i)) : UNIQUE_KEY_NONE;
                win = key[s] < win ? key[s] : win;
        }
        if ( win == UNIQUE_KEY_NONE )
                return;            // no rank has a hit: the local NoMatch/Gapped word stays
        uint32_t ties = 0;
        #pragma unroll
        for ( uint32_t s = 0; s < 8; ++s )
        {
                bool const samepos = (((key[s] ^ win) >> 17) & ((1ULL << 41) - 1)) == 0;
                ties += (key[s] != UNIQUE_KEY_NONE && (key[s] >> 59) == (win >> 59) && ! samepos) ? 1u : 0u;
        }
        uint32_t const err = (uint32_t)(win >> 59);
        bool const uniq = ((win >> 58) & 1) && ties == 0;
        uint32_t const strand = (uint32_t)((win >> 16) & 1);
        info_own[i] = umi_make(uniq ? (strand ? ST_REVERSE : ST_STRAIGHT) : ST_NONUNIQUE, (uint32_t)((win >> 52) & 63), (win >> 17) & UMI_POSMASK, err, (uint32_t)(win & 0xFFFF));
}

// ---- state checksum (bench / tests) -----------------------------------------------------------------
// Order independent digest of the unique state of the reads [first, first + n): sum over the reads of
// splitmix64(canonical word ^ splitmix64(read ordinal)) mod 2^64, where the canonical word keeps what the reference
// defines about a read independently of its visiting order (matcher.canonical_unique: everything for Straight / Reverse,
// state and error count for NonUnique, nothing but the state for NoMatch).  Sums of disjoint ranges add up, so the
// ranks of a multi-GPU job digest their own reads and the digests are added.
__global__ void __launch_bounds__(256) k_unique_checksum(const unsigned long long * __restrict__ info, uint64_t first, uint64_t n, unsigned long long * __restrict__ out)
{
        unsigned long long acc = 0;
        for ( uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x )
        {
                unsigned long long d = info[first + i];
                uint32_t const st = umi_state(d);
                if ( st == ST_NONUNIQUE ) d &= (7ULL << UMI_STATESHIFT) | (15ULL << UMI_ERRSHIFT);
                else if ( st == ST_NOMATCH ) d = 0;
                acc += splitmix64(d ^ splitmix64(first + i));
        }
        #pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ( (threadIdx.x & 31) == 0 && acc ) atomicAdd(out, acc);
}

// ---- synthetic inputs (same formulas as real_b200/synth.py) -------------------------------------

__device__ __forceinline__ uint64_t synth_stream(uint64_t seed, uint64_t tag) { return splitmix64(splitmix64(seed) + tag); }

__device__ __forceinline__ bool synth_nhit(uint64_t seed_stream2, uint64_t maskword, uint32_t n_per_million)
{
        return n_per_million && (splitmix64(maskword ^ seed_stream2) % 1000000ULL) < n_per_million;
}

// text word w = splitmix64(stream(seed,1) ^ w); words under an all-N mask word are stored as A
__global__ void __launch_bounds__(256) k_synth_text(uint64_t seed, uint64_t first_word, uint64_t nwords, uint32_t n_per_million,
                                                  uint64_t * __restrict__ words)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= nwords ) return;
        uint64_t const w = first_word + i;
        uint64_t v = splitmix64(w ^ synth_stream(seed, 1));
        if ( synth_nhit(synth_stream(seed, 2), w >> 1, n_per_million) ) v = 0;
        words[i] = v;
}

// N mask: each 64-base mask word is all-N with probability n_per_million / 1e6
__global__ void __launch_bounds__(256) k_synth_nmask(uint64_t seed, uint64_t first_mask_word, uint64_t nmaskwords, uint32_t n_per_million,
                                                   uint64_t * __restrict__ nmask)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= nmaskwords ) return;
        nmask[i] = synth_nhit(synth_stream(seed, 2), first_mask_word + i, n_per_million) ? ~0ULL : 0ULL;
}

// one thread per read; writes `length` mapped bytes (+ qualities).  Text must be resident (global coordinates
// from word 0).  Reads over wildcards get the code 4 like genpat prints 'N'.
__global__ void __launch_bounds__(128) k_synth_reads(uint64_t seed, const uint64_t * __restrict__ words, const uint64_t * __restrict__ nmask,
                                                   uint64_t text_n, uint64_t total, uint64_t first, uint64_t count, uint32_t L, uint32_t thr,
                                                   uint8_t * __restrict__ mapped, uint8_t * __restrict__ quality)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= count ) return;
        uint64_t const r = first + i;
        uint64_t const span = text_n - L + 1;
        // stratified start: (r*span)/total without overflow for span < 2^36, total < 2^28
        uint64_t const lo = (uint64_t)(((unsigned __int128)r * span) / total);
        uint64_t const hi = (uint64_t)(((unsigned __int128)(r + 1) * span) / total);
        uint64_t const width = (hi > lo) ? (hi - lo) : 1;
        uint64_t pos = lo + splitmix64(r ^ synth_stream(seed, 4)) % width;
        if ( pos > span - 1 ) pos = span - 1;
        uint32_t const strand = (uint32_t)(splitmix64(r ^ synth_stream(seed, 5)) >> 63);
        uint64_t const s6 = synth_stream(seed, 6);
        uint8_t * out = mapped + i * L;
        uint8_t * qo = quality ? (quality + i * L) : nullptr;
        uint64_t hc = 0;
        for ( uint32_t j = 0; j < L; ++j )
        {
                if ( (j & 3) == 0 ) hc = splitmix64((r * 64 + (j >> 2)) ^ s6);
                uint32_t const u16 = (uint32_t)(hc >> (16 * (j & 3))) & 0xFFFF;
                uint64_t const tp = strand ? (pos + L - 1 - j) : (pos + j);
                uint32_t b = (uint32_t)(words[tp >> 5] >> (62 - 2*(tp & 31))) & 3;
                bool const isn = nmask && ((nmask[tp >> 6] >> (63 - (tp & 63))) & 1);
                if ( strand ) b = 3 - b;
                bool changed = false;
                if ( isn ) b = 4;
                else if ( (u16 >> 2) < thr ) { b = (b + 1 + (u16 & 3) % 3) & 3; changed = true; }
                out[j] = (uint8_t)b;
                if ( qo ) qo[j] = changed ? (42 - 33) : (68 - 33);
        }
}

} // namespace realgpu
