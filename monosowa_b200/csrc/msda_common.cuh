// msda_common.cuh -- shared device helpers for the sm_100a MSDA kernels.
//
// Semantics reproduced (reference = MonoDETR/lib/models/monodetr/ops/src/cuda/
// ms_deform_im2col_cuda.cuh): sample coordinate and in-range window :285-291, bilinear
// taps with per-corner zero padding :33-84, gradient formulas :87-159.  The code below is
// organised around 128-bit channel vectors and lane groups, not around the reference's
// one-thread-per-channel layout.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_b200.h"

namespace msda {

struct Dims {
    int N, S, M, D, L, Lq, P;
};

struct LevelInfo {
    int H, W, start, pad;
};

constexpr unsigned kFullMask = 0xffffffffu;

// spatial_shapes / level_start_index live in device memory (int64, as the reference passes
// them); every CTA stages them once so that no thread re-reads them per sample and the host
// never has to look at them (CUDA-graph safe).
__device__ __forceinline__ void stage_levels(LevelInfo *s_lv, const int64_t *__restrict__ shapes,
                                             const int64_t *__restrict__ lsi, int L)
{
    if (threadIdx.x < (unsigned)L) {
        LevelInfo li;
        li.H = (int)shapes[2 * threadIdx.x];
        li.W = (int)shapes[2 * threadIdx.x + 1];
        li.start = (int)lsi[threadIdx.x];
        li.pad = 0;
        s_lv[threadIdx.x] = li;
    }
    __syncthreads();
}

// One bilinear sample: integer corner, fractional parts, per-corner validity.
template <typename CT>
struct Tap {
    int y0, x0;
    CT ly, lx;
    bool inside;
};

// cuh:285-288: `h_im = loc_h * spatial_h - 0.5`.  What the reference EXECUTES there is one fused multiply-add:
// nvcc (default -fmad=true) compiles the line to `FFMA h_im, loc_h, (float)spatial_h, -0.5` for scalar_t = float and
// to a DFMA for double (cuobjdump -sass of oracle/_ref/libmsda_ref_sm100.so, ms_deformable_im2col_gpu_kernel<float>).
// The single rounding matters exactly where MonoDETR starts training: the module's initial sampling offsets are whole
// pixels from a pixel centre (ms_deform_attn.py:106-115), so every initial sample sits ON a pixel boundary and
// floor() of the two-rounding form lands on the other side for ~1e-4 of them -- d out / d loc is discontinuous there
// (round 1 rounded the product first and disagreed with the reference kernels on 862 of 5.2 M initial gradients;
// tests/test_msda_gpu.py::test_matches_reference_cuda_kernels now covers that distribution).
__device__ __forceinline__ Tap<float> make_tap(float loc_x, float loc_y, int H, int W)
{
    Tap<float> t;
    const float py = __fmaf_rn(loc_y, (float)H, -0.5f);
    const float px = __fmaf_rn(loc_x, (float)W, -0.5f);
    t.inside = (py > -1.f) && (px > -1.f) && (py < (float)H) && (px < (float)W);
    const float fy = floorf(py), fx = floorf(px);
    t.y0 = (int)fy;
    t.x0 = (int)fx;
    t.ly = py - fy;
    t.lx = px - fx;
    return t;
}

__device__ __forceinline__ Tap<double> make_tap(double loc_x, double loc_y, int H, int W)
{
    Tap<double> t;
    const double py = __fma_rn(loc_y, (double)H, -0.5);
    const double px = __fma_rn(loc_x, (double)W, -0.5);
    t.inside = (py > -1.0) && (px > -1.0) && (py < (double)H) && (px < (double)W);
    const double fy = floor(py), fx = floor(px);
    t.y0 = (int)fy;
    t.x0 = (int)fx;
    t.ly = py - fy;
    t.lx = px - fx;
    return t;
}

// REDG.E.ADD.F32x4: one 16-byte reduction per lane instead of four scalar atomics.
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d)
{
    // The "memory" clobber keeps every later value load behind this reduction, i.e. the backward consumes its
    // samples strictly one after the other (4 loads in flight per lane).  Measured on B200: that order is the FASTEST
    // one -- issuing the loads of 2 or 4 samples together, with or without L1 allocation, at 80 or 128 registers,
    // costs 15-90 % (profiles/r01b_sweep_load_grouping.jsonl); dropping only the clobber costs 2 %.
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// REDG.E.ADD.BF16x4: the same reduction into a bf16 gradient (short query sets with bf16 values, msda_backward.cu)
__device__ __forceinline__ void red_add_bf16x4(__nv_bfloat16 *p, float a, float b, float c, float d)
{
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p), "r"(*reinterpret_cast<const unsigned *>(&lo)),
                 "r"(*reinterpret_cast<const unsigned *>(&hi))
                 : "memory");
}

__device__ __forceinline__ void red_add_x4(float *p, float a, float b, float c, float d) { red_add_f32x4(p, a, b, c, d); }
__device__ __forceinline__ void red_add_x4(__nv_bfloat16 *p, float a, float b, float c, float d) { red_add_bf16x4(p, a, b, c, d); }

// ---- lane-group reduce-scatter ---------------------------------------------------------------
// G lanes (a power of two, aligned inside the warp) each hold NV partial sums (NV a power of
// two).  Butterfly: at every step a lane keeps one half of its values and ships the other half
// to its partner, so log2(G) steps cost NV/2 + NV/4 + ... shuffles instead of NV*log2(G).
// Afterwards (gl = lane index inside the group):
//   NV >= G : v[0 .. NV/G) are the totals of elements [gl*NV/G, (gl+1)*NV/G)
//   NV <  G : v[0] is the total of element gl / (G/NV)   (duplicated across G/NV lanes)
template <int OFF, int NV>
struct ReduceScatter {
    static __device__ __forceinline__ void run(float *v, int gl)
    {
        if constexpr (OFF >= 1) {
            if constexpr (NV >= 2) {
                constexpr int H = NV / 2;
                const bool upper = (gl & OFF) != 0;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const float send = upper ? v[i] : v[i + H];
                    const float keep = upper ? v[i + H] : v[i];
                    v[i] = keep + __shfl_xor_sync(kFullMask, send, OFF);
                }
                ReduceScatter<OFF / 2, H>::run(v, gl);
            } else {
                v[0] += __shfl_xor_sync(kFullMask, v[0], OFF);
                ReduceScatter<OFF / 2, 1>::run(v, gl);
            }
        }
    }
};

template <int G, int NV>
__device__ __forceinline__ void group_reduce_scatter(float (&v)[NV], int gl)
{
    static_assert((G & (G - 1)) == 0 && (NV & (NV - 1)) == 0, "power-of-two sizes only");
    ReduceScatter<G / 2, NV>::run(v, gl);
}

// ---- work decomposition ----------------------------------------------------------------------
// A warp owns QPW = 32/G consecutive queries of ONE head (lane group k -> query chunk*QPW+k),
// so that neighbouring queries -- which sample neighbouring pixels -- share L1 lines and, on
// coarse levels, coalesce into the same 128-byte line within one load instruction.
//   order 0 : consecutive warps walk the heads of one query chunk, then the next chunk
//   order 1 : a CTA's warps walk consecutive query chunks of one head (larger per-head tiles)
struct WorkItem {
    int n, q, m;
    bool valid;
};

template <int QPW>
__device__ __forceinline__ WorkItem decode_work(const Dims &d, int order, int k)
{
    const int warps_per_block = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int n_chunks = (d.Lq + QPW - 1) / QPW;
    WorkItem w;
    long chunk;
    if (order == 0) {
        const long wg = (long)blockIdx.x * warps_per_block + warp;
        w.m = (int)(wg % d.M);
        const long r = wg / d.M;
        chunk = r % n_chunks;
        w.n = (int)(r / n_chunks);
    } else {
        const int n_tiles = (n_chunks + warps_per_block - 1) / warps_per_block;
        const long b = blockIdx.x;
        w.m = (int)(b % d.M);
        const long r = b / d.M;
        chunk = (r % n_tiles) * warps_per_block + warp;
        w.n = (int)(r / n_tiles);
    }
    w.q = (int)(chunk * QPW + k);
    w.valid = (w.n < d.N) && (chunk < n_chunks) && (w.q < d.Lq);
    return w;
}

inline long grid_for(const Dims &d, int order, int qpw, int threads)
{
    const long wpb = threads / 32;
    const long n_chunks = (d.Lq + qpw - 1) / qpw;
    if (order == 0) return ((long)d.N * n_chunks * d.M + wpb - 1) / wpb;
    const long n_tiles = (n_chunks + wpb - 1) / wpb;
    return (long)d.N * n_tiles * d.M;
}

// ---- host-side launch plumbing (msda_capi.cu) ------------------------------------------------
// A/B knobs for measurements (msda_set_tuning); -1 = the shipped default everywhere.
struct Tuning {
    int fwd_variant = -1;   // 11 record kernel always, 12 tile kernel always, 99 generic
    int bwd_variant = -1;   // 11 record kernel always, 20 tile kernel always, 21 binned kernel always, 99 generic
    int fwd_pipe = -1;      // launch flavours, -DMSDA_AB builds only
    int bwd_pipe = -1;
    int bf16_direct = -1;   // bf16 backward: largest average number of additions per grad_value row that may go
                            // straight into the bf16 gradient (default 4; see backward_needs_scratch)
};
Tuning &tuning();
void count_launch(int n = 1);

// dtype tags for the launchers
enum class DType { F32, F64, BF16 };

constexpr int kUnsupported = -1000;     // no kernel of the requested family fits: the caller falls back
constexpr int kNeedsScratch = -1001;    // bf16 backward: this shape needs the fp32 accumulation buffer

// implemented in msda_forward.cu / msda_backward.cu; return cudaError_t (or one of the two codes above)
int launch_forward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                   const void *loc, const void *attn, void *out, const Dims &d, bool vec_ok,
                   cudaStream_t st);
// grad_value: float for fp32 values, double for fp64, bf16 for bf16 values (`scratch`: fp32 accumulation
// buffer of N*S*M*D floats when backward_needs_scratch(), else unused)
int launch_backward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                    const void *loc, const void *attn, const void *grad_out, void *grad_value,
                    void *grad_loc, void *grad_attn, void *scratch, const Dims &d, bool vec_ok, cudaStream_t st);
bool backward_needs_scratch(const Dims &d, DType dt, bool vec_ok);
// fused pre-processing entry points; ref_dim = 2 or 6; kUnsupported when no fused kernel fits
int launch_forward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                         int ref_dim, const void *offsets, const void *logits, void *out, const Dims &d, cudaStream_t st);
int launch_backward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                          int ref_dim, const void *offsets, const void *logits, const void *grad_out, void *grad_value,
                          void *grad_offsets, void *grad_logits, void *scratch, const Dims &d, cudaStream_t st);
// msda_backward_tiled.cu: persistent grid over 2-D image tiles, grad_value combined in shared memory (long query
// sets); grad_value (fp32) pre-zeroed
// msda_backward_binned.cu: chunks of consecutive queries, the COARSE levels' grad_value combined in shared memory
// (long query sets: the default); grad_value (fp32) pre-zeroed
bool binned_backward_applies(const Dims &d, DType dt, bool vec_ok);
int launch_backward_binned(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                           const void *attn, const void *grad_out, void *grad_value, void *grad_loc, void *grad_attn,
                           const Dims &d, const void *ref, int ref_dim, cudaStream_t st);
bool tiled_backward_applies(const Dims &d, DType dt, bool vec_ok);
int launch_backward_tiled(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                          const void *attn, const void *grad_out, void *grad_value, void *grad_loc, void *grad_attn,
                          const Dims &d, const void *ref, int ref_dim, cudaStream_t st);
const char *forward_kernel_name(DType dt, const Dims &d, bool vec_ok);
const char *backward_kernel_name(DType dt, const Dims &d, bool vec_ok);

}  // namespace msda
