// Micro-benchmark: attainable rate of the MSDA gather pattern on B200 -- every 8-lane group loads a random
// 128-byte row (LDG.E.128 per lane, 4 rows per warp instruction) from a table of R rows, UNROLL independent
// loads in flight per lane.  Prints cycles per row per SM and the ms this rate implies for the 83.6 M row
// gathers of the headline forward.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_rows gather_rows.cu && ./gather_rows
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 4; }

template <int UNROLL>
__global__ void __launch_bounds__(256) gather(const float4 *__restrict__ table, uint32_t rows, int iters, float *out)
{
    const int lane = threadIdx.x & 31, gl = lane & 7;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x & ~7)) * 2654435761u + 99u;      // same stream per 8-lane group
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t r = lcg(seed) % rows;
            v[u] = __ldg(table + (size_t)r * 8 + gl);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x == 123.f) out[0] = acc.y + acc.z + acc.w;
}

template <int UNROLL>
void run(const float4 *table, uint32_t rows, int ctas_per_sm, int nsm, float *out, const char *label)
{
    const int iters = 4096 / UNROLL;
    const int grid = nsm * ctas_per_sm;
    gather<UNROLL><<<grid, 256>>>(table, rows, iters, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    gather<UNROLL><<<grid, 256>>>(table, rows, iters, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows_total = (double)grid * 8 * 4 * iters * UNROLL;            // 8 warps x 4 groups
    const double rows_per_s = rows_total / (ms * 1e-3);
    printf("%-22s unroll %2d, %d CTAs/SM: %.3f ms, %.1f G rows/s = %.2f TB/s, %.3f cycles/row/SM @1.965GHz -> 83.6M rows in %.3f ms\n",
           label, UNROLL, ctas_per_sm, ms, rows_per_s / 1e9, rows_per_s * 128 / 1e12,
           1.965e9 * nsm / rows_per_s, 83.6e6 / rows_per_s * 1e3);
}

int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t big = (size_t)167 << 20, small = (size_t)10 << 20;
    float4 *table; float *out;
    cudaMalloc(&table, big); cudaMemset(table, 0, big); cudaMalloc(&out, 16);
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t rows = (uint32_t)((pass == 0 ? small : big) / 128);
        const char *label = pass == 0 ? "10 MB table (L2 hit)" : "167 MB table";
        run<2>(table, rows, 6, nsm, out, label);
        run<4>(table, rows, 6, nsm, out, label);
        run<4>(table, rows, 8, nsm, out, label);
        run<8>(table, rows, 4, nsm, out, label);
        run<8>(table, rows, 8, nsm, out, label);
        run<16>(table, rows, 4, nsm, out, label);
    }
    return 0;
}
