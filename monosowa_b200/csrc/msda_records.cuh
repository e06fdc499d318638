// msda_records.cuh -- "shared geometry" building blocks of the second-generation kernels.
//
// Measured on B200 (profiles/r01_v1_ncu_full_summary.md): the first-generation vector kernels
// were instruction-issue bound -- every one of the G lanes that cover a (query, head) recomputed
// the same bilinear geometry (coordinate, floor, validity, weights, four addresses) for each of
// the L*P samples.  Here each lane of a lane group computes the geometry of ONE sample, drops a
// 32-byte record into shared memory, and all G lanes then consume the G records of their group
// with two broadcast LDS.128 per sample:
//     record = { int32 element offset of the 4 corners (clamped into the map), 4 weights }
// Consumption is branch-free so that all 4*G corner loads of a batch can be in flight together:
//   * an invalid corner keeps a clamped (in-map) address and a zero weight -- the clamped pixel is
//     always one of the sample's own valid corners;
//   * a sample outside the (-1,H)x(-1,W) window, or past L*P, points at element 0 of the image
//     (one hot L1 line -- these kernels gather with L1 allocation; measured: pointing such records at
//     per-query rows, or bypassing L1, is slower) with four zero weights.
// For finite `value` this is exactly the reference's skip logic (ms_deform_im2col_cuda.cuh:56-80,
// 288); a NaN/Inf stored at pixel 0 of a head would additionally reach queries that have outside
// samples (0 * NaN), which the reference's branches avoid -- documented in DESIGN.md.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kChannelsPerLane = 4;

// ---- 4-channel vectors: fp32 = 16-byte, bf16 = 8-byte accesses ------------------------------
template <typename VT>
struct Vec4;

template <>
struct Vec4<float> {
    static __device__ __forceinline__ void load(const float *p, float (&f)[4])
    {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[4])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(f[0], f[1], f[2], f[3]));      // written once, never re-read here
    }
    // gather loads of `value` rows.  H = 0: ld.global.nc (allocate in L1); 1: L1::no_allocate (hits are still
    // served by L1, misses do not evict); 2: ld.global.cg (L2 only)
    template <int H>
    static __device__ __forceinline__ void gather(const float *p, float (&f)[4])
    {
        if constexpr (H == 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "l"(p));
        } else if constexpr (H == 2) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p));
            f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        } else {
            load(p, f);
        }
    }
    static __device__ __forceinline__ void load_stream(const float *p, float (&f)[4])
    {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
};

template <>
struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[4])
    {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        f[0] = __uint_as_float(v.x << 16);
        f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16);
        f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[4])
    {
        const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
        const __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
        __stcs(reinterpret_cast<uint2 *>(p),
               make_uint2(*reinterpret_cast<const unsigned *>(&a), *reinterpret_cast<const unsigned *>(&b)));
    }
    template <int H>
    static __device__ __forceinline__ void gather(const __nv_bfloat16 *p, float (&f)[4])
    {
        uint2 v;
        if constexpr (H == 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
        } else if constexpr (H == 2) {
            v = __ldcg(reinterpret_cast<const uint2 *>(p));
        } else {
            v = __ldg(reinterpret_cast<const uint2 *>(p));
        }
        f[0] = __uint_as_float(v.x << 16);
        f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16);
        f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&f)[4])
    {
        const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(p));
        f[0] = __uint_as_float(v.x << 16);
        f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16);
        f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
};

// ---- per-warp record area --------------------------------------------------------------------
// Group k (of QPW = 32/G groups) owns G records, stored as G int4 offsets followed by G float4
// weights (so that the G lanes write 16-byte items at a 16-byte stride: conflict-free STS.128);
// groups are 4 words apart modulo the 32 banks so that the broadcast reads of the QPW groups
// never collide.
template <int G>
struct RecordLayout {
    static constexpr int QPW = 32 / G;
    static constexpr int GROUP_WORDS = G * 8 + 4;
    static constexpr int WARP_WORDS = QPW * GROUP_WORDS;
    static constexpr int WEIGHTS = G * 4;              // word offset of the weights inside a group
};

// What the owning lane keeps privately about its sample (needed again by the backward).
struct SampleGeom {
    float w00, w01, w10, w11;   // bilinear weights, zero for invalid corners
    float hy, ly, hx, lx;
    float a;                    // attention weight
    float Wf, Hf;
    unsigned vmask;             // bit0..3: corner 00, 01, 10, 11 lies inside the map
    int cell;                   // (y0+1)*(W+1) + (x0+1): the sample's base-corner cell on the (H+1)x(W+1) lattice
    bool live;                  // sample exists (index < L*P, query valid) and is inside the window
};

// The raw inputs of one sample, fetched ahead of use (the next batch's are in flight while the
// current batch is consumed).
struct SampleIn {
    float x, y, a;
};

__device__ __forceinline__ SampleIn fetch_sample(bool has, const float *__restrict__ loc,
                                                 const float *__restrict__ attn, long sample_index)
{
    SampleIn in{0.f, 0.f, 0.f};
    if (has) {
        // read-once streams: keep them from displacing the value lines that the gather re-uses in L1/L2
        const float2 xy = __ldcs(reinterpret_cast<const float2 *>(loc) + sample_index);
        in.x = xy.x; in.y = xy.y;
        in.a = __ldcs(attn + sample_index);
    }
    return in;
}

// ---- fused pre-processing -------------------------------------------------
// In fused mode the kernels consume the RAW outputs of the module's two Linears
// (reference ops/modules/ms_deform_attn.py:145-151):
//   loc  = reference_point[l] + sampling_offset / (W_l, H_l)          (2-dim reference points)
//   attn = softmax over the L*P logits of a (query, head)
// so the (N,Lq,M,L,P,2) locations and (N,Lq,M,L,P) weights never exist in HBM.  A lane owns the
// samples {gl, gl+G, gl+2G, ...} of its (query, head): at most kMaxBatches of them.
constexpr int kMaxBatches = 4;

template <int G>
__device__ __forceinline__ float group_allreduce_max(float v)
{
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, off));
    return v;
}

template <int G>
__device__ __forceinline__ float group_allreduce_sum(float v)
{
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(kFullMask, v, off);
    return v;
}

// softmax over the L*P logits of this lane group; a[b] is the weight of sample b*G + gl
template <int G>
__device__ __forceinline__ void group_softmax(const float *__restrict__ logits, long base, int LP, int gl, bool valid,
                                              float (&a)[kMaxBatches])
{
    float x[kMaxBatches];
    float m = -INFINITY;
#pragma unroll
    for (int b = 0; b < kMaxBatches; ++b) {
        const bool has = valid && (b * G + gl < LP);
        x[b] = has ? __ldg(logits + base + b * G + gl) : -INFINITY;
        m = fmaxf(m, x[b]);
    }
    m = group_allreduce_max<G>(m);
    float sum = 0.f;
#pragma unroll
    for (int b = 0; b < kMaxBatches; ++b) {
        a[b] = (x[b] == -INFINITY) ? 0.f : expf(x[b] - m);
        sum += a[b];
    }
    sum = group_allreduce_sum<G>(sum);
#pragma unroll
    for (int b = 0; b < kMaxBatches; ++b) a[b] = (sum > 0.f) ? a[b] / sum : 0.f;
}

// fused fetch: offset + reference point -> normalised location (same operation order as torch:
// ref + off / size, IEEE division); the attention weight comes from group_softmax
__device__ __forceinline__ SampleIn fetch_sample_fused(bool has, const float *__restrict__ offsets,
                                                       const float *__restrict__ ref, long sample_index,
                                                       long ref_index, const LevelInfo *s_lv, int l, float a)
{
    SampleIn in{0.f, 0.f, 0.f};
    if (has) {
        const float2 o = __ldcs(reinterpret_cast<const float2 *>(offsets) + sample_index);
        const float2 r = __ldg(reinterpret_cast<const float2 *>(ref) + ref_index);
        const LevelInfo li = s_lv[l];
        in.x = r.x + __fdiv_rn(o.x, (float)li.W);
        in.y = r.y + __fdiv_rn(o.y, (float)li.H);
        in.a = a;
    }
    return in;
}

// Build the record of one sample and write it to `rec_off` / `rec_w` (16-byte aligned).  The
// weights stored are w_ij * a (what both forward and grad_value need).  The four offsets are
// (row0 + y * W + x) * stride | tag: for rows in global memory row0 = the level's start index and
// stride = M * D; a kernel that keeps a level in shared memory passes the level's first shared row, stride = D
// and a tag bit that tells the consumer where to read (msda_forward_resident.cu).
__device__ __forceinline__ SampleGeom build_record_at(uint32_t *rec_off, uint32_t *rec_w, bool has, const SampleIn in,
                                                      const LevelInfo *s_lv, int l, int row0, int stride, int tag)
{
    SampleGeom gm;
    gm.live = false;
    gm.w00 = gm.w01 = gm.w10 = gm.w11 = 0.f;
    gm.hy = gm.ly = gm.hx = gm.lx = 0.f;
    gm.a = 0.f; gm.Wf = 0.f; gm.Hf = 0.f; gm.vmask = 0u; gm.cell = 0;
    int4 off = make_int4(0, 0, 0, 0);
    float4 wa = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has) {
        const float a = in.a;
        const LevelInfo li = s_lv[l];
        const Tap<float> t = make_tap(in.x, in.y, li.H, li.W);
        if (t.inside) {
            const bool y0ok = t.y0 >= 0, y1ok = t.y0 + 1 <= li.H - 1;
            const bool x0ok = t.x0 >= 0, x1ok = t.x0 + 1 <= li.W - 1;
            const int y0c = max(t.y0, 0), y1c = min(t.y0 + 1, li.H - 1);
            const int x0c = max(t.x0, 0), x1c = min(t.x0 + 1, li.W - 1);
            gm.hy = 1.f - t.ly; gm.ly = t.ly; gm.hx = 1.f - t.lx; gm.lx = t.lx;
            const float wy0 = y0ok ? gm.hy : 0.f, wy1 = y1ok ? gm.ly : 0.f;
            const float wx0 = x0ok ? gm.hx : 0.f, wx1 = x1ok ? gm.lx : 0.f;
            gm.w00 = wy0 * wx0; gm.w01 = wy0 * wx1; gm.w10 = wy1 * wx0; gm.w11 = wy1 * wx1;
            gm.a = a; gm.Wf = (float)li.W; gm.Hf = (float)li.H;
            gm.vmask = (y0ok && x0ok ? 1u : 0u) | (y0ok && x1ok ? 2u : 0u) | (y1ok && x0ok ? 4u : 0u) |
                       (y1ok && x1ok ? 8u : 0u);
            gm.live = true;
            gm.cell = (t.y0 + 1) * (li.W + 1) + t.x0 + 1;
            const int r0 = (row0 + y0c * li.W) * stride, r1 = (row0 + y1c * li.W) * stride;
            off = make_int4((r0 + x0c * stride) | tag, (r0 + x1c * stride) | tag, (r1 + x0c * stride) | tag,
                            (r1 + x1c * stride) | tag);
            wa = make_float4(gm.w00 * a, gm.w01 * a, gm.w10 * a, gm.w11 * a);
        }
    }
    *reinterpret_cast<int4 *>(rec_off) = off;
    *reinterpret_cast<float4 *>(rec_w) = wa;
    return gm;
}

__device__ __forceinline__ SampleGeom build_record(uint32_t *rec_off, uint32_t *rec_w, bool has, const SampleIn in,
                                                   const LevelInfo *s_lv, int l, int xs)
{
    return build_record_at(rec_off, rec_w, has, in, s_lv, l, has ? s_lv[l].start : 0, xs, 0);
}

}  // namespace msda
