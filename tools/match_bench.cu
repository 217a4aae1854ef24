// development microbenchmark: throughput of __match_any_sync on 8-bit values vs the same peer mask from 8 ballots
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template<int MODE>
__global__ void k(uint32_t * out, int iters, uint32_t bits)
{
        uint32_t acc = 0, s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
        uint32_t const mask = (1u << bits) - 1;
        for ( int i = 0; i < iters; ++i )
        {
                s = s * 1664525u + 1013904223u;
                uint32_t const v = (s >> 11) & mask;
                uint32_t peers;
                if ( MODE == 0 )
                        peers = __match_any_sync(0xffffffffu, v);
                else
                {
                        peers = 0xffffffffu;
                        #pragma unroll
                        for ( int b = 0; b < 8; ++b )
                        {
                                if ( (uint32_t)b >= bits ) break;
                                uint32_t const bal = __ballot_sync(0xffffffffu, (v >> b) & 1);
                                peers &= ((v >> b) & 1) ? bal : ~bal;
                        }
                }
                acc += __popc(peers) + (peers & 1);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
        uint32_t * d; cudaMalloc(&d, 148 * 8 * 256 * 4);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        int const iters = 20000;
        for ( int bits : { 8, 4 } )
        for ( int mode = 0; mode < 2; ++mode )
        for ( int bps : { 1, 4, 8 } )
        {
                for ( int rep = 0; rep < 2; ++rep )
                {
                        cudaEventRecord(a);
                        if ( mode == 0 ) k<0><<<148 * bps, 256>>>(d, iters, bits); else k<1><<<148 * bps, 256>>>(d, iters, bits);
                        cudaEventRecord(b); cudaEventSynchronize(b);
                }
                float ms; cudaEventElapsedTime(&ms, a, b);
                double const warp_ops = 148.0 * bps * 8 * iters;
                printf("%s bits=%d blocks/SM=%d: %.3f ms, %.2f cycles per warp-op per SM (at 1.9 GHz)\n", mode ? "ballot8" : "match  ", bits, bps, ms, ms * 1e-3 * 1.9e9 / (warp_ops / 148.0));
        }
        return 0;
}
