// Micro-benchmark (VERDICT r1 "next" 2, the gate): native INTEGER shared-memory atomics (ATOMS.ADD) as a way to
// accumulate 128-byte grad_value rows inside the SM in fixed point -- sm_100a has no native fp32 shared atomic
// (ATOMS.CAST.SPIN loops, 14.5 cycles per row: profiles/r01_ubench_smem_atomics.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_int_atomics smem_int_atomics.cu && ./smem_int_atomics
// A warp = 4 lane groups of 8 lanes; a lane owns 4 channels of a 32-channel row; in every update each group adds
// its 4 x 8 values to a pseudo-random row of a ROWS x 32-word tile (4 ATOMS.ADD per lane).
//   MODE 0  ATOMS.ADD, channel order rotated per group so that the 32 lanes of one instruction hit 32 banks
//   MODE 1  same + the conversion a real kernel needs (FMUL by the scale, F2I.RNI) in front of every atomic
//   MODE 2  ATOMS.ADD, naive channel order (the 4 groups collide on the same 8 banks: 4-way conflict)
//   MODE 3  reference: plain LDS.128 + FADD + STS.128 read-modify-write (racy; what a native fp32 atomic would cost)
// Reported: cycles per 128-byte row per SM, for 8 / 16 / 32 resident warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 512;            // 64 KB tile
constexpr int ITER = 4096;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *gout, float scale)
{
    extern __shared__ __align__(16) int tile[];
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) tile[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & 7, grp = lane >> 3;
    uint32_t seed = (blockIdx.x * 256 + threadIdx.x) / 8 * 2654435761u + 12345u;      // one stream per lane group
    float v0 = 0.37f + gl, v1 = 1.1f + gl, v2 = -0.6f + gl, v3 = 0.02f * gl;
#pragma unroll 4
    for (int it = 0; it < ITER; ++it) {
        const int row = lcg(seed) % ROWS;
        int *p = tile + row * 32 + gl * 4;
        if (MODE == 3) {
            float4 *q = reinterpret_cast<float4 *>(p);
            float4 o = *q;
            o.x += v0; o.y += v1; o.z += v2; o.w += v3;
            *q = o;
        } else {
            int a0, a1, a2, a3;
            if (MODE == 1) {
                a0 = __float2int_rn(v0 * scale); a1 = __float2int_rn(v1 * scale);
                a2 = __float2int_rn(v2 * scale); a3 = __float2int_rn(v3 * scale);
                v0 += 1e-3f; v1 -= 1e-3f; v2 += 2e-3f; v3 -= 2e-3f;
            } else {
                a0 = it; a1 = it + 1; a2 = it + 2; a3 = it + 3;
            }
            if (MODE == 2) {
                atomicAdd(p + 0, a0); atomicAdd(p + 1, a1); atomicAdd(p + 2, a2); atomicAdd(p + 3, a3);
            } else {            // group g starts at channel g: lanes of one instruction cover all 32 banks
                atomicAdd(p + ((0 + grp) & 3), a0); atomicAdd(p + ((1 + grp) & 3), a1);
                atomicAdd(p + ((2 + grp) & 3), a2); atomicAdd(p + ((3 + grp) & 3), a3);
            }
        }
    }
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) s += (float)tile[i];
    if (s == 123.456f) gout[0] = s;
}

template <int MODE>
void run(const char *name, int ctas_per_sm, int sms, float *gout)
{
    const size_t smem = ROWS * 32 * 4;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * ctas_per_sm, 256, smem>>>(gout, 1048576.f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<sms * ctas_per_sm, 256, smem>>>(gout, 1048576.f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double rows_per_sm = (double)ctas_per_sm * 8 * 4 * ITER;       // 8 warps x 4 groups x ITER rows
    printf("%-44s %2d warps/SM  %s  %.3f ms  %.2f cycles per 128-B row per SM\n", name, ctas_per_sm * 8,
           cudaGetErrorString(cudaGetLastError()), ms, cycles / rows_per_sm);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *gout;
    cudaMalloc(&gout, 4);
    printf("SMs %d, tile %d rows (64 KB), %d row updates per lane group\n", sms, ROWS, ITER);
    for (int c : {1, 2, 3}) {
        run<0>("ATOMS.ADD s32, bank-rotated", c, sms, gout);
        run<1>("FMUL + F2I.RNI + ATOMS.ADD, bank-rotated", c, sms, gout);
        run<2>("ATOMS.ADD s32, naive (4-way bank conflict)", c, sms, gout);
        run<3>("plain LDS.128 + FADD + STS.128 (racy floor)", c, sms, gout);
    }
    return 0;
}
