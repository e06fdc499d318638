// msda_records.cuh -- "shared geometry" building blocks of the record / tile kernels.
//
// Measured on B200 (profiles/r01_v1_ncu_full_summary.md): the first-generation vector kernels
// were instruction-issue bound -- every one of the G lanes that cover a (query, head) recomputed
// the same bilinear geometry (coordinate, floor, validity, weights, four addresses) for each of
// the L*P samples.  Here each lane of a lane group computes the geometry of ONE sample, drops a
// 32-byte record into shared memory, and all G lanes then consume the G records of their group
// with two broadcast LDS.128 per sample:
//     record = { int32 element offset of the 4 corners (clamped into the map), 4 weights }
//   * an invalid corner keeps a clamped (in-map) address and a zero weight -- the clamped pixel is
//     always one of the sample's own valid corners;
//   * the forward packs only the records of samples inside the (-1,H)x(-1,W) window (the reference's
//     branch, ms_deform_im2col_cuda.cuh:288): outside samples cost nothing and read nothing;
//   * the backward walks all G records of a batch (it needs a fixed sample -> lane mapping for its
//     reduce-scatter); an outside sample points at element 0 of the image with four zero weights, its
//     partial dots are DISCARDED by the owning lane with a select (not multiplied), so a NaN stored
//     there cannot leak -- the reference's skip logic (cuh:56-80, 288, 365-367) exactly.
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

// ---- CPL-channel vectors: CPL = 4 (above) or 8 (fp32: one 256-bit LDG.E.ENL2.256, bf16: one 128-bit load)
template <typename VT, int CPL>
struct VecN;

template <typename VT>
struct VecN<VT, 4> {
    template <int H>
    static __device__ __forceinline__ void gather(const VT *p, float (&f)[4]) { Vec4<VT>::template gather<H>(p, f); }
    static __device__ __forceinline__ void store(VT *p, const float (&f)[4]) { Vec4<VT>::store(p, f); }
    static __device__ __forceinline__ void load(const VT *p, float (&f)[4]) { Vec4<VT>::load(p, f); }
    static __device__ __forceinline__ void load_stream(const VT *p, float (&f)[4]) { Vec4<VT>::load_stream(p, f); }
};

template <>
struct VecN<float, 8> {
    template <int H>
    static __device__ __forceinline__ void gather(const float *p, float (&f)[8])
    {
        if constexpr (H == 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
        } else if constexpr (H == 3) {     // as 1, and the row is the last to leave L2
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
        } else {
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "l"(p));
        }
    }
    static __device__ __forceinline__ void store(float *p, const float (&f)[8])
    {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(f[0], f[1], f[2], f[3]));
        __stcs(reinterpret_cast<float4 *>(p) + 1, make_float4(f[4], f[5], f[6], f[7]));
    }
    static __device__ __forceinline__ void load(const float *p, float (&f)[8])
    {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void load_stream(const float *p, float (&f)[8])
    {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(p)), b = __ldcs(reinterpret_cast<const float4 *>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};

template <>
struct VecN<__nv_bfloat16, 8> {
    template <int H>
    static __device__ __forceinline__ void gather(const __nv_bfloat16 *p, float (&f)[8])
    {
        uint4 v;
        if constexpr (H == 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        } else {
            v = __ldg(reinterpret_cast<const uint4 *>(p));
        }
        const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(u[i] << 16);
            f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&f)[8])
    {
        unsigned u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            u[i] = *reinterpret_cast<const unsigned *>(&h);
        }
        __stcs(reinterpret_cast<uint4 *>(p), make_uint4(u[0], u[1], u[2], u[3]));
    }
    static __device__ __forceinline__ void widen(const uint4 v, float (&f)[8])
    {
        const unsigned u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(u[i] << 16);
            f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&f)[8]) { widen(__ldg(reinterpret_cast<const uint4 *>(p)), f); }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p, float (&f)[8]) { widen(__ldcs(reinterpret_cast<const uint4 *>(p)), f); }
};

// ---- per-warp record area --------------------------------------------------------------------
// Group k (of QPW = 32/G groups) owns G records, stored as G int4 offsets followed by G float4
// weights (so that the G lanes write 16-byte items at a 16-byte stride).  Groups are 4 words apart
// modulo the 32 banks so that the broadcast READS of the QPW groups (same record index, one address
// per group) never collide -- there are 2*G reads for every 2 writes.  For G < 8 the record WRITES of
// neighbouring groups then overlap in 3 of 4 bank quads (ncu, 4-lane forward: 6.2 M of 24.3 M shared
// wavefronts); the alternative stride 4*G (conflict-free writes, 4-way read conflicts: 10.7 M of 28.8 M)
// was measured no faster (profiles/r02_fwd_layout_interleaved.jsonl) and a stride that serves both does
// not exist without a per-record swizzle (reads need stride/4 odd, writes need stride = 16 mod 32).
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
    int cy, cx;                 // (y0+1, x0+1): the sample's base-corner cell on the (H+1)x(W+1) lattice
    bool live;                  // sample exists (index < L*P, query valid) and is inside the window
};

// The raw inputs of one sample, fetched ahead of use (the next batch's are in flight while the
// current batch is consumed).
struct SampleIn {
    float x, y, a;
    float ex, ey;               // fused 6-dim reference points only: (l+r, t+b) of the reference box
};

// Loads of the read-once input streams (locations / offsets, weights).  A lane group reads the G samples of a batch,
// i.e. G*8 bytes of its (query, head)'s location line per batch, the batches microseconds apart.
// SP = 0: ld.global.cs (evict-first: keeps the streams from displacing the value lines the gather re-uses) -- for
//   G >= 8, where a batch asks for 64 bytes, what DRAM delivers.
// SP = 1: ld.global.nc (evict-normal) -- for G < 8: a batch asks for one 32-byte sector, and the evict-first half of
//   the 64-byte DRAM burst that is not asked for yet is gone again when its batch comes (ncu, 8-channel fp32 forward at
//   configs[1]: 572 MB of DRAM reads with SP = 0, 418 MB = the compulsory inputs with SP = 1, same run time;
//   profiles/r02_fwd_dram_by_flavour.txt).
// SP = 2: evict-first with an L2::128B prefetch hint (A/B: does not keep the line either, 574 MB).
// bf16 values keep SP = 0 whatever G is: there the evict-normal streams cost time instead (D = 32 forward at configs[1]:
// 0.450 ms / 334 MB read against 0.431 ms / 529 MB -- the run time is what counts; profiles/r02_fwd_stream_policy_bf16.jsonl).
template <int G, typename VT>
__host__ __device__ constexpr int stream_policy() { return (G < 8 && sizeof(VT) == 4) ? 1 : 0; }

template <int SP>
__device__ __forceinline__ float2 ld_stream2(const float2 *p)
{
    if constexpr (SP == 1) return __ldg(p);
    else if constexpr (SP == 2) {
        float2 v;
        asm volatile("ld.global.cs.L2::128B.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
        return v;
    } else return __ldcs(p);
}
template <int SP>
__device__ __forceinline__ float ld_stream1(const float *p)
{
    if constexpr (SP == 1) return __ldg(p);
    else if constexpr (SP == 2) {
        float v;
        asm volatile("ld.global.cs.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
        return v;
    } else return __ldcs(p);
}

template <int SP = 0>
__device__ __forceinline__ SampleIn fetch_sample(bool has, const float *__restrict__ loc,
                                                 const float *__restrict__ attn, long sample_index)
{
    SampleIn in{0.f, 0.f, 0.f, 0.f, 0.f};
    if (has) {
        const float2 xy = ld_stream2<SP>(reinterpret_cast<const float2 *>(loc) + sample_index);
        in.x = xy.x; in.y = xy.y;
        in.a = ld_stream1<SP>(attn + sample_index);
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

// fused fetch: offset + reference point -> normalised location, same operation order as the reference module
// (ops/modules/ms_deform_attn.py:149-155), every step rounded (no FMA contraction):
//   ref_dim 2:  loc = ref + off / (W, H)
//   ref_dim 6:  loc = ref[:2] + ((off / P) * (ref[2]+ref[3], ref[4]+ref[5])) * 0.5
// the attention weight comes from group_softmax
template <int SP = 0>
__device__ __forceinline__ SampleIn fetch_sample_fused(bool has, const float *__restrict__ offsets,
                                                       const float *__restrict__ ref, int ref_dim, long sample_index,
                                                       long ref_index, const LevelInfo *s_lv, int l, int P, float a)
{
    SampleIn in{0.f, 0.f, 0.f, 0.f, 0.f};
    if (has) {
        const float2 o = ld_stream2<SP>(reinterpret_cast<const float2 *>(offsets) + sample_index);
        if (ref_dim == 2) {
            const float2 r = __ldg(reinterpret_cast<const float2 *>(ref) + ref_index);
            const LevelInfo li = s_lv[l];
            in.x = __fadd_rn(r.x, __fdiv_rn(o.x, (float)li.W));
            in.y = __fadd_rn(r.y, __fdiv_rn(o.y, (float)li.H));
        } else {
            const float2 *rp = reinterpret_cast<const float2 *>(ref) + ref_index * 3;
            const float2 r01 = __ldg(rp), r23 = __ldg(rp + 1), r45 = __ldg(rp + 2);
            in.ex = __fadd_rn(r23.x, r23.y);
            in.ey = __fadd_rn(r45.x, r45.y);
            in.x = __fadd_rn(r01.x, __fmul_rn(__fmul_rn(__fdiv_rn(o.x, (float)P), in.ex), 0.5f));
            in.y = __fadd_rn(r01.y, __fmul_rn(__fmul_rn(__fdiv_rn(o.y, (float)P), in.ey), 0.5f));
        }
        in.a = a;
    }
    return in;
}

// d loc / d offset applied to a location gradient, in autograd's operation order
__device__ __forceinline__ float2 fused_offset_grad(int ref_dim, float gx, float gy, float Wf, float Hf, float ex,
                                                    float ey, int P)
{
    if (ref_dim == 2) return make_float2(__fdiv_rn(gx, Wf), __fdiv_rn(gy, Hf));
    return make_float2(__fdiv_rn(__fmul_rn(__fmul_rn(gx, 0.5f), ex), (float)P),
                       __fdiv_rn(__fmul_rn(__fmul_rn(gy, 0.5f), ey), (float)P));
}

// Geometry of one sample: the four corner offsets (lsi_l + y * W + x) * xs, clamped into the map, and the
// four weights w_ij * a (what both the forward and grad_value need); `off` / `wa` are all zero for a sample
// that does not exist or lies outside the window.
__device__ __forceinline__ SampleGeom sample_geometry(bool has, const SampleIn in, const LevelInfo *s_lv, int l, int xs,
                                                      int4 &off, float4 &wa)
{
    SampleGeom gm;
    gm.live = false;
    gm.w00 = gm.w01 = gm.w10 = gm.w11 = 0.f;
    gm.hy = gm.ly = gm.hx = gm.lx = 0.f;
    gm.a = 0.f; gm.Wf = 0.f; gm.Hf = 0.f; gm.vmask = 0u; gm.cy = 0; gm.cx = 0;
    off = make_int4(0, 0, 0, 0);
    wa = make_float4(0.f, 0.f, 0.f, 0.f);
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
            gm.cy = t.y0 + 1;
            gm.cx = t.x0 + 1;
            const int r0 = (li.start + y0c * li.W) * xs, r1 = (li.start + y1c * li.W) * xs;
            off = make_int4(r0 + x0c * xs, r0 + x1c * xs, r1 + x0c * xs, r1 + x1c * xs);
            wa = make_float4(gm.w00 * a, gm.w01 * a, gm.w10 * a, gm.w11 * a);
        }
    }
    return gm;
}

// Build the record of one sample and write it to `rec_off` / `rec_w` (16-byte aligned).
__device__ __forceinline__ SampleGeom build_record(uint32_t *rec_off, uint32_t *rec_w, bool has, const SampleIn in,
                                                   const LevelInfo *s_lv, int l, int xs)
{
    int4 off;
    float4 wa;
    const SampleGeom gm = sample_geometry(has, in, s_lv, l, xs, off, wa);
    *reinterpret_cast<int4 *>(rec_off) = off;
    *reinterpret_cast<float4 *>(rec_w) = wa;
    return gm;
}

}  // namespace msda
