// Micro-benchmark: cost of accumulating 128-byte rows (8 lanes x 16 B) into shared memory with
// the different atomic flavours sm_100a offers, versus REDG.128 to global/L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
// Each warp performs ITER updates; in every update each 8-lane group adds a float4 per lane to a
// pseudo-random row of a ROWS x 32-float tile.  Reports cycles per warp-instruction-equivalent
// (= 4 row updates) per SM, with all SMs busy (8 warps x 4 CTAs per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 1024;           // 128 KB tile
constexpr int ITER = 2048;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__device__ __forceinline__ void add_f32_cas(float *p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ void add_f32x2_cas64(float *p, float a, float b)
{
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *q, assumed;
    do {
        assumed = old;
        float lo = __uint_as_float((unsigned)assumed) + a;
        float hi = __uint_as_float((unsigned)(assumed >> 32)) + b;
        unsigned long long nv = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
        old = atomicCAS(q, assumed, nv);
    } while (old != assumed);
}

__device__ __forceinline__ void add_f32x4_cas128(float *p, float4 v)
{
    unsigned addr = (unsigned)__cvta_generic_to_shared(p);
    float4 old = *reinterpret_cast<float4 *>(p);
    while (true) {
        float4 nv = make_float4(old.x + v.x, old.y + v.y, old.z + v.z, old.w + v.w);
        unsigned long long rlo, rhi;
        asm volatile(
            "{\n\t.reg .b128 cmp, swp, res;\n\t"
            "mov.b128 cmp, {%3, %4};\n\t"
            "mov.b128 swp, {%5, %6};\n\t"
            "atom.shared.cas.b128 res, [%2], cmp, swp;\n\t"
            "mov.b128 {%0, %1}, res;\n\t}"
            : "=l"(rlo), "=l"(rhi)
            : "r"(addr),
              "l"(((unsigned long long)__float_as_uint(old.y) << 32) | __float_as_uint(old.x)),
              "l"(((unsigned long long)__float_as_uint(old.w) << 32) | __float_as_uint(old.z)),
              "l"(((unsigned long long)__float_as_uint(nv.y) << 32) | __float_as_uint(nv.x)),
              "l"(((unsigned long long)__float_as_uint(nv.w) << 32) | __float_as_uint(nv.z))
            : "memory");
        float4 got = make_float4(__uint_as_float((unsigned)rlo), __uint_as_float((unsigned)(rlo >> 32)),
                                 __uint_as_float((unsigned)rhi), __uint_as_float((unsigned)(rhi >> 32)));
        if (got.x == old.x && got.y == old.y && got.z == old.z && got.w == old.w) break;   // (bitwise-equal for finite data)
        old = got;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float *gout, long long *cycles)
{
    extern __shared__ __align__(16) float tile[];
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) tile[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane & 7, grp = lane >> 3;
    uint32_t seed = (blockIdx.x * 256 + (threadIdx.x & ~7)) * 2654435761u + 12345u;   // same per 8-lane group
    (void)grp;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        const int row = lcg(seed) % ROWS;
        float *p = tile + row * 32 + gl * 4;
        if (MODE == 0) { add_f32_cas(p, v.x); add_f32_cas(p + 1, v.y); add_f32_cas(p + 2, v.z); add_f32_cas(p + 3, v.w); }
        if (MODE == 1) { add_f32x2_cas64(p, v.x, v.y); add_f32x2_cas64(p + 2, v.z, v.w); }
        if (MODE == 2) { add_f32x4_cas128(p, v); }
        if (MODE == 3) {   // plain (racy) read-modify-write: the non-atomic floor
            float4 o = *reinterpret_cast<float4 *>(p);
            *reinterpret_cast<float4 *>(p) = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
        }
        if (MODE == 5) {   // TMA bulk reduce: stage the 128-B row in smem, one cp.reduce.async.bulk per row
            float *stage = tile + ((threadIdx.x >> 5) * 2 + (it & 1)) * 128 + (lane >> 3) * 32;   // per-warp double buffer
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            *reinterpret_cast<float4 *>(stage + gl * 4) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (gl == 0) {
                float *g = gout + ((size_t)blockIdx.x % 592) * ROWS * 32 + row * 32;
                unsigned saddr = (unsigned)__cvta_generic_to_shared(stage);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(g), "r"(saddr) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (MODE == 6) {   // split: groups 0-1 of each warp reduce through TMA, groups 2-3 through REDG.128
            float *stage = tile + ((threadIdx.x >> 5) * 2 + (it & 1)) * 128 + (lane >> 3) * 32;
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            float *g = gout + ((size_t)blockIdx.x % 592) * ROWS * 32 + row * 32;
            if (lane < 16) *reinterpret_cast<float4 *>(stage + gl * 4) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < 16) {
                if (gl == 0) {
                    unsigned saddr = (unsigned)__cvta_generic_to_shared(stage);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(g), "r"(saddr) : "memory");
                }
            } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(g + gl * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (MODE == 4) {   // REDG.128 to a global tile of the same shape (per-CTA distinct region)
            float *g = gout + ((size_t)blockIdx.x % 592) * ROWS * 32 + row * 32 + gl * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(g), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
    }
    if (MODE == 5 || MODE == 6) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0.f;
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) s += tile[i];
    if (s == 123.456f) gout[0] = s;
}

template <int MODE>
void run(const char *name, float *gout, long long *cyc, int nsm)
{
    const int ctas = nsm * 1;        // 128 KB tile -> 1 CTA (8 warps) per SM
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 128);
    k<MODE><<<ctas, 256, ROWS * 128>>>(gout, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256, ROWS * 128>>>(gout, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    // per SM: 8 warps x ITER warp-updates, each = 4 row updates
    double per_warp_update = avg / (8.0 * ITER);
    printf("%-28s %s  %.3f ms  %.1f cycles per warp-update (4 rows) per SM -> %.2f cycles per 128-B row\n", name,
           cudaGetErrorString(err), ms, per_warp_update, per_warp_update / 4);
}

int main()
{
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *gout; long long *cyc;
    cudaMalloc(&gout, (size_t)592 * ROWS * 128); cudaMemset(gout, 0, (size_t)592 * ROWS * 128);
    cudaMalloc(&cyc, sizeof(long long) * 4096);
    printf("SMs %d, 1 CTA x 8 warps per SM, %d updates per warp, tile %d rows\n", nsm, ITER, ROWS);
    run<3>("plain RMW (racy floor)", gout, cyc, nsm);
    run<0>("atomicAdd f32 x4 (CAS32)", gout, cyc, nsm);
    run<1>("CAS64 loop x2", gout, cyc, nsm);
    run<2>("CAS128 loop x1", gout, cyc, nsm);
    run<4>("REDG.128 global", gout, cyc, nsm);
    run<5>("TMA bulk reduce 128 B/row", gout, cyc, nsm);
    run<6>("half TMA + half REDG.128", gout, cyc, nsm);
    return 0;
}
