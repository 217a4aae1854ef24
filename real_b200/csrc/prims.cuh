// Hand-written device-wide primitives used by the index build and the hit ordering: warp / block scans
// and an exclusive prefix sum over u32.  (The reference's RadixSort32/64, ParallelRadixSort.hpp:29-484,
// has no counterpart here: the tables are built by partitioning, not by sorting -- see index.cuh.)
#pragma once

#include "common.cuh"

namespace realgpu
{

// ------------------------------------------------------------------------------------------
// exclusive scan, u32, n < 2^32
// ------------------------------------------------------------------------------------------

static const int SCAN_THREADS = 256;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
        #pragma unroll
        for ( int o = 1; o < 32; o <<= 1 )
        {
                uint32_t const t = __shfl_up_sync(0xffffffffu, v, o);
                if ( lane >= o ) v += t;
        }
        return v;
}

// one bit of the multisplit: the ballot of bit `mask` of v and the word that turns it into "lanes that agree with me"
// (0 where the bit is set, ~0 where it is clear).  Written in PTX: the compiler's own rendering of the C version
// spends six instructions per bit (shift, and, compare, vote, select, combine), this one four.
__device__ __forceinline__ void ballot_bit(uint32_t v, uint32_t mask, uint32_t & bal, uint32_t & flip)
{
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     ".reg .b32 t;\n\t"
                     "and.b32 t, %2, %3;\n\t"
                     "setp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                     "selp.b32 %1, 0, 0xffffffff, p;\n\t"
                     "}" : "=r"(bal), "=r"(flip) : "r"(v), "r"(mask));
}

// Lanes of the warp that hold the same 8-bit value as this lane (among the lanes with ok set; lanes without ok get
// an unspecified mask).  One ballot per bit: measured on B200 (tools/match_bench.cu) MATCH.ANY costs ~2 cycles per
// DISTINCT value on a unit shared by the whole SM -- 58 cycles per warp for random bytes, which made the multisplit
// ranking the bottleneck of every partition kernel -- against ~32 cycles for the eight ballots.
__device__ __forceinline__ uint32_t peers_u8(uint32_t v, bool ok)
{
        uint32_t peers = __ballot_sync(0xffffffffu, ok);
        #pragma unroll
        for ( int b = 0; b < 8; ++b )
        {
                uint32_t bal, flip;
                ballot_bit(v, 1u << b, bal, flip);
                peers &= bal ^ flip;
        }
        return peers;
}

// the same when only the low nbits bits of v differ between the lanes (warp uniform nbits)
__device__ __forceinline__ uint32_t peers_low(uint32_t v, bool ok, uint32_t nbits)
{
        uint32_t peers = __ballot_sync(0xffffffffu, ok);
        #pragma unroll
        for ( int b = 0; b < 8; ++b )
                if ( (uint32_t)b < nbits )
                {
                        uint32_t bal, flip;
                        ballot_bit(v, 1u << b, bal, flip);
                        peers &= bal ^ flip;
                }
        return peers;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t * total)
{
        __shared__ uint32_t warp_sums[32];
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t const incl = warp_incl_scan(v, lane);
        if ( lane == 31 ) warp_sums[wid] = incl;
        __syncthreads();
        if ( wid == 0 )
        {
                uint32_t s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
                s = warp_incl_scan(s, lane);
                warp_sums[lane] = s;
        }
        __syncthreads();
        uint32_t const base = wid ? warp_sums[wid-1] : 0;
        *total = warp_sums[(blockDim.x >> 5) - 1];
        __syncthreads();
        return base + incl - v;
}

// The same for a block of 256 threads with ONE barrier: scratch = eight words of the caller's shared memory that no thread
// writes again before every thread has passed another block barrier (the staged partition kernels have several per tile).
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t * total, uint32_t * scratch)
{
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t const incl = warp_incl_scan(v, lane);
        if ( lane == 31 ) scratch[wid] = incl;
        __syncthreads();
        uint32_t base = 0, all = 0;
        #pragma unroll
        for ( int w = 0; w < 8; ++w )
        {
                uint32_t const x = scratch[w];
                if ( w < wid ) base += x;
                all += x;
        }
        *total = all;
        return base + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t * __restrict__ in, uint32_t * __restrict__ sums, uint64_t n)
{
        uint64_t const base = (uint64_t)blockIdx.x * SCAN_TILE;
        uint32_t acc = 0;
        #pragma unroll
        for ( int i = 0; i < SCAN_ITEMS; ++i )
        {
                uint64_t const idx = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
                if ( idx < n ) acc += in[idx];
        }
        uint32_t total;
        block_excl_scan(acc, &total);
        if ( threadIdx.x == 0 ) sums[blockIdx.x] = total;
}

// in-place capable: out may alias in
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t * __restrict__ in, uint32_t * __restrict__ out, const uint32_t * __restrict__ offsets, uint64_t n)
{
        uint64_t const base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t acc = 0;
        #pragma unroll
        for ( int i = 0; i < SCAN_ITEMS; ++i )
        {
                uint64_t const idx = base + i;
                v[i] = (idx < n) ? in[idx] : 0;
                acc += v[i];
        }
        uint32_t total;
        uint32_t run = block_excl_scan(acc, &total) + (offsets ? offsets[blockIdx.x] : 0);
        #pragma unroll
        for ( int i = 0; i < SCAN_ITEMS; ++i )
        {
                uint64_t const idx = base + i;
                if ( idx < n ) out[idx] = run;
                run += v[i];
        }
}

inline uint64_t scan_temp_elems(uint64_t n)
{
        uint64_t total = 0;
        while ( n > 1 )
        {
                uint64_t const nb = (n + SCAN_TILE - 1) / SCAN_TILE;
                total += nb;
                if ( nb == 1 ) break;
                n = nb;
        }
        return total + 1;
}

// out[i] = sum of in[0..i); temp holds scan_temp_elems(n) u32
inline uint32_t exclusive_scan_u32(const uint32_t * in, uint32_t * out, uint64_t n, uint32_t * temp, cudaStream_t st, uint32_t * launches)
{
        if ( n == 0 ) return 0;
        uint64_t const nb = (n + SCAN_TILE - 1) / SCAN_TILE;
        if ( nb == 1 )
        {
                k_scan_apply<<<1, SCAN_THREADS, 0, st>>>(in, out, nullptr, n);
                RG_KERNEL_CHECK();
                if ( launches ) *launches += 1;
                return 1;
        }
        k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, temp, n);
        RG_KERNEL_CHECK();
        if ( launches ) *launches += 1;
        exclusive_scan_u32(temp, temp, nb, temp + nb, st, launches);
        k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, temp, n);
        RG_KERNEL_CHECK();
        if ( launches ) *launches += 1;
        return 0;
}

} // namespace realgpu
