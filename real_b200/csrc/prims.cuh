// Hand-written device-wide primitives used by the index build and the hit ordering:
// exclusive prefix sum over u32 and a stable LSD radix sort of (u32 key, u32 value) pairs.
//
// The radix sort is this design's counterpart of the reference's RadixSort32
// (ParallelRadixSort.hpp:29-214, u_sort.hpp:105-131): stable, least-significant-digit first.
// 8-bit digits, one upsweep (per-tile digit histogram), one scan and one downsweep
// (warp-level multi-split ranking with __match_any_sync, then scatter) per pass.
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

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (key,val) u32 pairs, 8-bit digits
// ------------------------------------------------------------------------------------------

static const int RS_THREADS = 256;
static const int RS_WARPS = RS_THREADS / 32;
static const int RS_ITEMS = 16;                         // keys per thread
static const int RS_TILE = RS_THREADS * RS_ITEMS;       // 4096 keys per block
static const int RS_BINS = 256;

// histogram of one digit per tile, stored bin-major: hist[bin * nblk + blk]
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t * __restrict__ keys, uint32_t * __restrict__ hist, uint32_t n, uint32_t nblk, uint32_t shift)
{
        __shared__ uint32_t h[RS_BINS];
        h[threadIdx.x] = 0;
        __syncthreads();
        uint32_t const base = blockIdx.x * RS_TILE;
        #pragma unroll
        for ( int i = 0; i < RS_ITEMS; ++i )
        {
                uint32_t const idx = base + i * RS_THREADS + threadIdx.x;
                if ( idx < n )
                        atomicAdd(&h[(keys[idx] >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// stable scatter.  Tile order is: warp w owns keys [w*512, (w+1)*512) of the tile, item j of lane l
// is key w*512 + j*32 + l, so ranking by (warp, j, lane) preserves input order.
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint32_t * __restrict__ keys, const uint32_t * __restrict__ vals,
                                                         uint32_t * __restrict__ okeys, uint32_t * __restrict__ ovals,
                                                         const uint32_t * __restrict__ hist_scanned, uint32_t n, uint32_t nblk, uint32_t shift)
{
        __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];     // per-warp digit counts, then per-warp start offsets
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for ( int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS )
                (&wcnt[0][0])[i] = 0;
        __syncthreads();

        uint32_t const wbase = blockIdx.x * RS_TILE + wid * (32 * RS_ITEMS);
        uint32_t k[RS_ITEMS], v[RS_ITEMS], off[RS_ITEMS];
        #pragma unroll
        for ( int j = 0; j < RS_ITEMS; ++j )
        {
                uint32_t const idx = wbase + j * 32 + lane;
                bool const ok = idx < n;
                k[j] = ok ? keys[idx] : 0xFFFFFFFFu;
                v[j] = ok ? vals[idx] : 0;
        }
        #pragma unroll
        for ( int j = 0; j < RS_ITEMS; ++j )
        {
                uint32_t const idx = wbase + j * 32 + lane;
                bool const ok = idx < n;
                uint32_t const d = (k[j] >> shift) & 0xFF;
                // lanes holding the same digit (invalid lanes grouped apart via bit 8)
                uint32_t const peers = __match_any_sync(0xffffffffu, ok ? d : (0x100u | d));
                uint32_t const below = __popc(peers & ((1u << lane) - 1));
                uint32_t pre = 0;
                if ( ok ) pre = wcnt[wid][d];
                __syncwarp();
                if ( ok && below == 0 ) wcnt[wid][d] = pre + __popc(peers);
                __syncwarp();
                off[j] = pre + below;
        }
        __syncthreads();
        // digit d (thread d): turn per-warp counts into start offsets in the output
        {
                uint32_t run = hist_scanned[threadIdx.x * nblk + blockIdx.x];
                #pragma unroll
                for ( int w = 0; w < RS_WARPS; ++w )
                {
                        uint32_t const c = wcnt[w][threadIdx.x];
                        wcnt[w][threadIdx.x] = run;
                        run += c;
                }
        }
        __syncthreads();
        #pragma unroll
        for ( int j = 0; j < RS_ITEMS; ++j )
        {
                uint32_t const idx = wbase + j * 32 + lane;
                if ( idx < n )
                {
                        uint32_t const d = (k[j] >> shift) & 0xFF;
                        uint32_t const dst = wcnt[wid][d] + off[j];
                        okeys[dst] = k[j];
                        ovals[dst] = v[j];
                }
        }
}

struct RadixSortTemp
{
        uint32_t * hist;       // 256 * nblk
        uint32_t * scan_tmp;   // scan_temp_elems(256 * nblk)
};

inline uint64_t rs_num_blocks(uint64_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// sorts by the low `bits` bits of the key.  Data ping-pongs between (k0,v0) and (k1,v1); returns 0 if
// the result is in (k0,v0), 1 if in (k1,v1).
inline int radix_sort_pairs(uint32_t * k0, uint32_t * v0, uint32_t * k1, uint32_t * v1, uint64_t n, uint32_t bits,
                            RadixSortTemp const & T, cudaStream_t st, uint32_t * launches)
{
        if ( n == 0 ) return 0;
        uint32_t const nblk = (uint32_t)rs_num_blocks(n);
        int cur = 0;
        for ( uint32_t shift = 0; shift < bits; shift += 8 )
        {
                uint32_t * ik = cur ? k1 : k0; uint32_t * iv = cur ? v1 : v0;
                uint32_t * ok = cur ? k0 : k1; uint32_t * ov = cur ? v0 : v1;
                k_rs_hist<<<nblk, RS_THREADS, 0, st>>>(ik, T.hist, (uint32_t)n, nblk, shift);
                RG_KERNEL_CHECK();
                if ( launches ) *launches += 1;
                exclusive_scan_u32(T.hist, T.hist, (uint64_t)RS_BINS * nblk, T.scan_tmp, st, launches);
                k_rs_scatter<<<nblk, RS_THREADS, 0, st>>>(ik, iv, ok, ov, T.hist, (uint32_t)n, nblk, shift);
                RG_KERNEL_CHECK();
                if ( launches ) *launches += 1;
                cur ^= 1;
        }
        return cur;
}

} // namespace realgpu
