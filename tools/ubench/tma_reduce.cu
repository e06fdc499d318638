// Micro-benchmark: is the TMA bulk-reduce path (cp.reduce.async.bulk ... .add.f32, SASS UBLKRED) an
// independent way out of the SM, or does it share the L1TEX->XBAR port that bounds REDG.128?
// Every warp pushes 128-byte rows to random rows of a global table: MODE 0 = REDG.128 only,
// MODE 1 = TMA bulk reduce only (deep ring of staging buffers per warp), MODE 2 = half and half.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_reduce tma_reduce.cu && ./tma_reduce
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NST = 8;            // staging ring depth per warp (4 rows of 128 B per stage)
constexpr int ITER = 2048;
constexpr uint32_t ROWS = 1u << 17;   // 16 MB table: L2-resident, so the SM-side port is what is measured

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *gout, long long *cycles)
{
    __shared__ __align__(128) float stage[8 * NST * 128];          // 8 warps x NST x 4 rows x 32 floats
    const int lane = threadIdx.x & 31, gl = lane & 7, grp = lane >> 3, warp = threadIdx.x >> 5;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x & ~7)) * 2654435761u + 777u;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        const uint32_t row = lcg(seed) % ROWS;
        float *g = gout + (size_t)row * 32;
        const bool tma = MODE == 1 || (MODE == 2 && grp < 2);
        if (MODE != 0) {
            float *st = stage + ((warp * NST + (it % NST)) * 4 + grp) * 32;
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NST - 1) : "memory");
            __syncwarp();
            if (tma) *reinterpret_cast<float4 *>(st + gl * 4) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (tma && gl == 0) {
                unsigned saddr = (unsigned)__cvta_generic_to_shared(st);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(g), "r"(saddr) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (!tma)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(g + gl * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    if (MODE != 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, float *gout, long long *cyc, int nsm, int cps)
{
    const int ctas = nsm * cps;
    k<MODE><<<ctas, 256>>>(gout, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(gout, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)ctas * 8 * 4 * ITER;
    printf("%-26s %d CTAs/SM  %s  %.3f ms  %.1f G rows/s  -> %.2f cycles per 128-B row per SM @1.965 GHz\n", name, cps,
           cudaGetErrorString(err), ms, rows / ms / 1e6, 1.965e9 * nsm / (rows / (ms * 1e-3)));
}

int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *gout; long long *cyc;
    cudaMalloc(&gout, (size_t)ROWS * 128); cudaMemset(gout, 0, (size_t)ROWS * 128);
    cudaMalloc(&cyc, sizeof(long long) * 8192);
    for (int cps : {2, 4}) {
        run<0>("REDG.128 only", gout, cyc, nsm, cps);
        run<1>("TMA bulk reduce only", gout, cyc, nsm, cps);
        run<2>("half TMA + half REDG", gout, cyc, nsm, cps);
    }
    return 0;
}
