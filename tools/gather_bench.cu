// Microbenchmark (development tool, not part of the product): random 32-byte-sector gather
// throughput on B200 for the access shapes the text-scan kernel can use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if ( e != cudaSuccess ) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
        uint64_t z = x + 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
}

template<int MODE> __device__ __forceinline__ uint32_t ld(const uint32_t * p)
{
        uint32_t v;
        if ( MODE == 0 ) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        else if ( MODE == 1 ) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
        else if ( MODE == 2 ) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
        else asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
}

// each thread: iters rounds of U independent random 4-byte loads
template<int U, int MODE>
__global__ void __launch_bounds__(256) k_ldg(const uint32_t * __restrict__ tab, uint64_t mask_words, int iters, uint32_t * out)
{
        uint64_t const tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t acc = 0;
        uint64_t s = mix(tid);
        for ( int it = 0; it < iters; ++it )
        {
                uint32_t v[U];
                #pragma unroll
                for ( int u = 0; u < U; ++u )
                {
                        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                        v[u] = ld<MODE>(tab + ((s >> 20) & mask_words));
                }
                #pragma unroll
                for ( int u = 0; u < U; ++u ) acc ^= v[u];
        }
        if ( acc == 0x12345678 ) out[0] = acc;
}

// cp.async 4B gathers into shared memory, U per thread per round, one round in flight while the previous is consumed
template<int U>
__global__ void __launch_bounds__(256) k_cpasync(const uint32_t * __restrict__ tab, uint64_t mask_words, int iters, uint32_t * out)
{
        __shared__ uint32_t buf[2][U][256];
        uint64_t const tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t acc = 0;
        uint64_t s = mix(tid);
        auto issue = [&](int b)
        {
                #pragma unroll
                for ( int u = 0; u < U; ++u )
                {
                        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                        uint32_t const dst = (uint32_t)__cvta_generic_to_shared(&buf[b][u][threadIdx.x]);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(tab + ((s >> 20) & mask_words)) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
        };
        issue(0);
        for ( int it = 0; it < iters; ++it )
        {
                int const b = it & 1;
                if ( it + 1 < iters ) { issue(b ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                #pragma unroll
                for ( int u = 0; u < U; ++u ) acc ^= buf[b][u][threadIdx.x];
        }
        if ( acc == 0x12345678 ) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cp.async.bulk (TMA unit) gathers of BYTES (16 or 32) per request, U requests per thread per round
template<int U, int BYTES>
__global__ void __launch_bounds__(256) k_bulk(const uint32_t * __restrict__ tab, uint64_t mask_units, int iters, uint32_t * out)
{
        extern __shared__ __align__(128) unsigned char dyn[];
        __shared__ __align__(8) uint64_t bar[2];
        unsigned char * bufs[2] = { dyn, dyn + (size_t)U * 256 * BYTES };
        uint64_t const tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( threadIdx.x == 0 )
        {
                for ( int b = 0; b < 2; ++b )
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[b])), "r"(256));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        uint32_t acc = 0;
        uint64_t s = mix(tid);
        auto issue = [&](int b)
        {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[b])), "r"(U * BYTES) : "memory");
                #pragma unroll
                for ( int u = 0; u < U; ++u )
                {
                        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
                        const unsigned char * src = reinterpret_cast<const unsigned char *>(tab) + ((s >> 20) & mask_units) * BYTES;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(smem_u32(bufs[b] + ((size_t)u * 256 + threadIdx.x) * BYTES)), "l"(src), "r"(BYTES), "r"(smem_u32(&bar[b])) : "memory");
                }
        };
        auto wait = [&](int b, uint32_t parity)
        {
                uint32_t done = 0;
                while ( ! done )
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(done) : "r"(smem_u32(&bar[b])), "r"(parity) : "memory");
        };
        issue(0);
        for ( int it = 0; it < iters; ++it )
        {
                int const b = it & 1;
                if ( it + 1 < iters ) issue(b ^ 1);
                wait(b, (it >> 1) & 1);
                #pragma unroll
                for ( int u = 0; u < U; ++u ) acc ^= *reinterpret_cast<uint32_t *>(bufs[b] + ((size_t)u * 256 + threadIdx.x) * BYTES);
                __syncthreads();
        }
        if ( acc == 0x12345678 ) out[0] = acc;
}

template<typename F> float timeit(F f)
{
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        f(); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        CK(cudaGetLastError());
        return ms;
}

int main(int argc, char ** argv)
{
        int sms = 148;
        if ( argc > 1 )
        {
                CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])));
        }
        size_t gran = 0; CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
        printf("L2 fetch granularity limit: %zu\n", gran);
        bool const quick = argc > 2;
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); sms = prop.multiProcessorCount;
        uint32_t * out; CK(cudaMalloc(&out, 64));
        size_t const sizes[] = { (size_t)64 << 20, (size_t)2 << 30, (size_t)8 << 30 };
        for ( size_t sz : sizes )
        {
                uint32_t * tab; CK(cudaMalloc(&tab, sz)); CK(cudaMemset(tab, 1, sz));
                uint64_t const mask_words = sz / 4 - 1;
                printf("== table %zu MB\n", sz >> 20);
                int const iters = 64;
                #define RUN_LDG(U, MODE, BPS) { int blocks = sms * BPS; float ms = timeit([&]{ k_ldg<U,MODE><<<blocks,256>>>(tab, mask_words, iters, out); }); \
                        double n = (double)blocks * 256 * iters * U; printf("ldg mode=%d U=%2d blocks/SM=%d : %7.2f G sectors/s  (%6.0f GB/s sector traffic)\n", MODE, U, BPS, n / ms / 1e6, n * 32 / ms / 1e6); }
                RUN_LDG(8,0,8) RUN_LDG(16,0,8) RUN_LDG(32,0,8)
                if ( quick ) { CK(cudaFree(tab)); continue; }
                RUN_LDG(1,0,8) RUN_LDG(4,0,8) RUN_LDG(8,0,4) RUN_LDG(16,0,4) RUN_LDG(24,0,4) RUN_LDG(32,0,4)
                RUN_LDG(16,1,8) RUN_LDG(16,2,8) RUN_LDG(16,3,8) RUN_LDG(32,1,8) RUN_LDG(32,3,8)
                #define RUN_CPA(U, BPS) { int blocks = sms * BPS; float ms = timeit([&]{ k_cpasync<U><<<blocks,256>>>(tab, mask_words, iters, out); }); \
                        double n = (double)blocks * 256 * iters * U; printf("cp.async4 U=%2d blocks/SM=%d : %7.2f G sectors/s\n", U, BPS, n / ms / 1e6); }
                RUN_CPA(4,8) RUN_CPA(8,4) RUN_CPA(8,8) RUN_CPA(16,4)
                #define RUN_BULK(U, BYTES, BPS) { int blocks = sms * BPS; size_t sm = (size_t)2 * U * 256 * BYTES; \
                        CK(cudaFuncSetAttribute(k_bulk<U,BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
                        float ms = timeit([&]{ k_bulk<U,BYTES><<<blocks,256,sm>>>(tab, sz / BYTES - 1, iters, out); }); \
                        double n = (double)blocks * 256 * iters * U; printf("bulk%d U=%2d blocks/SM=%d : %7.2f G requests/s\n", BYTES, U, BPS, n / ms / 1e6); }
                RUN_BULK(1,16,4) RUN_BULK(2,16,4) RUN_BULK(4,16,4) RUN_BULK(4,32,2) RUN_BULK(2,32,4)
                CK(cudaFree(tab));
        }
        return 0;
}
