// Micro-benchmark (results: profiles/r01b_ubench_lines_per_instruction.txt): what does a gathered 128-byte row cost as a function
// of HOW MANY DIFFERENT LINES one warp instruction touches, and where the row lives?
//
// The record kernels read a row with 8 lanes x 16 B, four rows per LDG.128 warp instruction, and sustain ~1.8 cycles
// per row per SM (profiles/r01_ubench_gather_rows.txt).  The B300 notes (/opt/skills/guides/B300_MICROARCH.md, "L1tex
// wavefront queue") put one wavefront at ~1 cycle across instructions but ~2 cycles per extra line WITHIN one
// instruction -- which would make (1 + 3 * 2) / 4 = 1.75 cycles per row, i.e. the observed rate, a property of the
// access shape rather than of L2.  The resident-forward experiment (rows served by LDS.128) did
// not get faster, which argues against it.  This program measures the pieces in isolation:
//   ldg<LANES>   every group of LANES lanes loads one random row of an L2-resident table:
//                LANES = 32 -> LDG.32,  one line per instruction;  16 -> LDG.64, two lines;  8 -> LDG.128, four lines;
//                4 -> LDG.256 (sm_100+), eight lines
//   lds<LANES>   the same shapes from a 64 KB shared-memory table (LDS.32 / LDS.64 / LDS.128)
//   redg<SMS>    REDG.128 rows (8 lanes x 16 B) into an L2-resident table from only the first SMS CTAs' worth of SMs:
//                if cycles per row per SM stay ~5.3 with a quarter of the SMs active the limit is the SM's port,
//                if they drop it is the L2 side.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lines_per_instruction lines_per_instruction.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 4; }

constexpr int kIters = 2048;
constexpr int kUnroll = 4;

template <int LANES> struct Vec;
template <> struct Vec<32> { using T = float; };
template <> struct Vec<16> { using T = float2; };
template <> struct Vec<8> { using T = float4; };

__device__ __forceinline__ float sum(float v) { return v; }
__device__ __forceinline__ float sum(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float sum(float4 v) { return v.x + v.y + v.z + v.w; }

// LANES lanes cover one 128-byte row; a warp instruction therefore touches 32 / LANES different rows
template <int LANES>
__global__ void __launch_bounds__(256) ldg_rows(const float *__restrict__ table, uint32_t rows, float *out)
{
    using T = typename Vec<LANES>::T;
    const int lane = threadIdx.x & 31, part = lane % LANES;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x - part)) * 2654435761u + 7u;       // same stream per lane group
    float acc = 0.f;
    for (int it = 0; it < kIters / kUnroll; ++it) {
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t r = lcg(seed) % rows;
            v[u] = __ldg(reinterpret_cast<const T *>(table + (size_t)r * 32) + part);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc += sum(v[u]);
    }
    if (acc == 123.f) out[0] = acc;
}

// four lanes x 32 B (LDG.E.ENL2.256, sm_100+): eight different rows per warp instruction
__global__ void __launch_bounds__(256) ldg_rows_256(const float *__restrict__ table, uint32_t rows, float *out)
{
    const int lane = threadIdx.x & 31, part = lane % 4;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x - part)) * 2654435761u + 7u;
    float acc = 0.f;
    for (int it = 0; it < kIters / kUnroll; ++it) {
        float v[kUnroll][8];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t r = lcg(seed) % rows;
            const float *p = table + (size_t)r * 32 + part * 8;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                           "=f"(v[u][6]), "=f"(v[u][7])
                         : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc += v[u][c];
    }
    if (acc == 123.f) out[0] = acc;
}

template <int LANES>
__global__ void __launch_bounds__(256) lds_rows(float *out)
{
    using T = typename Vec<LANES>::T;
    constexpr uint32_t kRows = 512;                                                     // 64 KB
    extern __shared__ __align__(16) float s_table[];
    for (int i = threadIdx.x; i < kRows * 32; i += 256) s_table[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, part = lane % LANES;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x - part)) * 2654435761u + 7u;
    float acc = 0.f;
    for (int it = 0; it < kIters / kUnroll; ++it) {
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t r = lcg(seed) % kRows;
            v[u] = reinterpret_cast<const T *>(s_table + r * 32)[part];
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc += sum(v[u]);
    }
    if (acc == 123.f) out[0] = acc;
}

__global__ void __launch_bounds__(256) redg_rows(float *table, uint32_t rows)
{
    const int lane = threadIdx.x & 31, part = lane & 7;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x - part)) * 2654435761u + 7u;
    for (int it = 0; it < kIters; ++it) {
        const uint32_t r = lcg(seed) % rows;
        float *p = table + (size_t)r * 32 + part * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
    }
}

template <typename F>
static float time_ms(F launch)
{
    launch();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

static void report(const char *what, int lanes, float ms, int grid, int nsm_active)
{
    const double rows = (double)grid * 8 * (32 / lanes) * kIters;
    const double cyc = 1.965e9 * nsm_active * (ms * 1e-3) / rows;
    printf("%-28s %2d lanes/row (%d lines per instruction): %.3f ms, %.3f cycles per row per SM\n", what, lanes, 32 / lanes, ms, cyc);
}

int main()
{
    int nsm;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const size_t bytes = (size_t)10 << 20;                                             // L2 resident
    const uint32_t rows = (uint32_t)(bytes / 128);
    float *table, *out;
    cudaMalloc(&table, bytes); cudaMemset(table, 0, bytes); cudaMalloc(&out, 16);
    const int per_sm = 6, grid = nsm * per_sm;
    report("LDG, L2-resident table", 32, time_ms([&] { ldg_rows<32><<<grid, 256>>>(table, rows, out); }), grid, nsm);
    report("LDG, L2-resident table", 16, time_ms([&] { ldg_rows<16><<<grid, 256>>>(table, rows, out); }), grid, nsm);
    report("LDG, L2-resident table", 8, time_ms([&] { ldg_rows<8><<<grid, 256>>>(table, rows, out); }), grid, nsm);
    report("LDG.256, L2-resident table", 4, time_ms([&] { ldg_rows_256<<<grid, 256>>>(table, rows, out); }), grid, nsm);
    report("LDG.256, 4 CTAs/SM", 4, time_ms([&] { ldg_rows_256<<<nsm * 4, 256>>>(table, rows, out); }), nsm * 4, nsm);
    report("LDG.128, 4 CTAs/SM", 8, time_ms([&] { ldg_rows<8><<<nsm * 4, 256>>>(table, rows, out); }), nsm * 4, nsm);
    const int smem = 64 * 1024;
    cudaFuncSetAttribute(lds_rows<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(lds_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(lds_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid_s = nsm * 3;                                                        // 3 x 64 KB per SM
    report("LDS, 64 KB shared table", 32, time_ms([&] { lds_rows<32><<<grid_s, 256, smem>>>(out); }), grid_s, nsm);
    report("LDS, 64 KB shared table", 16, time_ms([&] { lds_rows<16><<<grid_s, 256, smem>>>(out); }), grid_s, nsm);
    report("LDS, 64 KB shared table", 8, time_ms([&] { lds_rows<8><<<grid_s, 256, smem>>>(out); }), grid_s, nsm);
    for (int frac = 1; frac <= 4; frac *= 2) {                                         // all, half, a quarter of the SMs
        const int sms = nsm / frac, g = sms * 2;                                       // <= one wave: 2 CTAs per active SM
        const float ms = time_ms([&] { redg_rows<<<g, 256>>>(table, rows); });
        const double rws = (double)g * 8 * 4 * kIters;
        printf("REDG.128 rows, %3d CTAs (~%3d SMs busy): %.3f ms, %.1f G rows/s chip-wide, %.2f cycles per row per busy SM\n", g, sms,
               ms, rws / (ms * 1e-3) / 1e9, 1.965e9 * sms * (ms * 1e-3) / rws);
    }
    return 0;
}
