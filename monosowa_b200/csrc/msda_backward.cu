// msda_backward.cu -- MSDA backward kernels for sm_100a.
//
// Replaces the reference's seven col2im kernels and their dispatcher (reference
// MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920, 956-1327;
// gradient formulas :87-159).  Machine mapping (see also msda_forward.cu):
//   * a lane owns a 128-bit channel vector; D/VEC lanes form the lane group of one (query, head);
//   * grad_value: one REDG.E.ADD.F32x4 (red.global.add.v4.f32) per lane per corner instead of
//     four scalar atomicAdds -- the reference issues 4 scalar atomics per channel per sample;
//   * grad_sampling_loc / grad_attn_weight: per-lane partial dot products are combined with a
//     butterfly reduce-scatter over the lane group (warp shuffles only) and written once, fully
//     overwriting the outputs -- no shared memory, no block barriers, no zero-fill needed
//     (the reference: smem staging, 2 __syncthreads and a serial D-way sum per sample point);
//   * level geometry is staged in shared memory once per CTA.
#include "msda_common.cuh"
#include "msda_records.cuh"

namespace msda {

template <typename VT, int D, int P>
__global__ void __launch_bounds__(512)
bwd_vec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, const VT *__restrict__ grad_out,
               float *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_attn, const Dims d, const int order)
{
    constexpr int VEC = Vec<VT>::N;
    constexpr int G = D / VEC;
    constexpr int QPW = 32 / G;
    static_assert(D % VEC == 0 && G >= 1 && G <= 32 && (32 % G) == 0, "unsupported D");

    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    stage_levels(s_lv, shapes, lsi, d.L);

    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    WorkItem w = decode_work<QPW>(d, order, lane / G);
    // warps are all-or-nothing past the end of the problem; inside a live warp, lane groups
    // beyond Lq keep running (they take part in the shuffles) with every sample masked off.
    if (__ballot_sync(kFullMask, w.valid) == 0) return;
    if (!w.valid) { w.n = 0; w.q = 0; w.m = 0; }

    const long qm = ((long)w.n * d.Lq + w.q) * d.M + w.m;
    const long img = ((long)w.n * d.S * d.M + w.m) * D + gl * VEC;
    const VT *vimg = value + img;
    float *gvimg = grad_value + img;
    const float *lp = loc + qm * (long)(d.L * P * 2);
    const float *ap = attn + qm * (long)(d.L * P);
    float *glp = grad_loc + qm * (long)(d.L * P * 2);
    float *gap = grad_attn + qm * (long)(d.L * P);
    const int xs = d.M * D;

    float g[VEC];
    Vec<VT>::load(grad_out + qm * D + gl * VEC, g);

    for (int l = 0; l < d.L; ++l) {
        const LevelInfo li = s_lv[l];
        const long lvl_off = (long)li.start * xs;
        const VT *vl = vimg + lvl_off;
        float *gvl = gvimg + lvl_off;
        const int ys = li.W * xs;

        float lxy[2 * P], aw[P];
#pragma unroll
        for (int i = 0; i < P / 2; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(lp) + i);
            lxy[4 * i] = t.x; lxy[4 * i + 1] = t.y; lxy[4 * i + 2] = t.z; lxy[4 * i + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < P / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(ap) + i);
            aw[4 * i] = t.x; aw[4 * i + 1] = t.y; aw[4 * i + 2] = t.z; aw[4 * i + 3] = t.w;
        }
        lp += 2 * P;
        ap += P;

        float pxy[2 * P], pa[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            pxy[2 * p] = 0.f; pxy[2 * p + 1] = 0.f; pa[p] = 0.f;
            const Tap<float> t = make_tap(lxy[2 * p], lxy[2 * p + 1], li.H, li.W);
            if (!t.inside || !w.valid) continue;
            const bool y0ok = t.y0 >= 0, y1ok = t.y0 + 1 <= li.H - 1;
            const bool x0ok = t.x0 >= 0, x1ok = t.x0 + 1 <= li.W - 1;
            const int o00 = t.y0 * ys + t.x0 * xs;
            float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) v00[c] = v01[c] = v10[c] = v11[c] = 0.f;
            if (y0ok && x0ok) Vec<VT>::load(vl + o00, v00);
            if (y0ok && x1ok) Vec<VT>::load(vl + o00 + xs, v01);
            if (y1ok && x0ok) Vec<VT>::load(vl + o00 + ys, v10);
            if (y1ok && x1ok) Vec<VT>::load(vl + o00 + ys + xs, v11);
            const float hy = 1.f - t.ly, hx = 1.f - t.lx;
            const float w00 = hy * hx, w01 = hy * t.lx, w10 = t.ly * hx, w11 = t.ly * t.lx;
            const float a = aw[p];
            float sx = 0.f, sy = 0.f, sa = 0.f;
            float tg[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                tg[c] = g[c] * a;                                          // cuh:111
                const float dgx = hy * (v01[c] - v00[c]) + t.ly * (v11[c] - v10[c]);   // cuh:119-149
                const float dgy = hx * (v10[c] - v00[c]) + t.lx * (v11[c] - v01[c]);
                sx += dgx * tg[c];
                sy += dgy * tg[c];
                sa += g[c] * (w00 * v00[c] + w01 * v01[c] + w10 * v10[c] + w11 * v11[c]);  // cuh:155-156
            }
            pxy[2 * p] = (float)li.W * sx;                                 // cuh:157
            pxy[2 * p + 1] = (float)li.H * sy;                             // cuh:158
            pa[p] = sa;
#pragma unroll
            for (int c = 0; c < VEC; c += 4) {
                if (y0ok && x0ok)
                    red_add_f32x4(gvl + o00 + c, w00 * tg[c], w00 * tg[c + 1], w00 * tg[c + 2], w00 * tg[c + 3]);
                if (y0ok && x1ok)
                    red_add_f32x4(gvl + o00 + xs + c, w01 * tg[c], w01 * tg[c + 1], w01 * tg[c + 2], w01 * tg[c + 3]);
                if (y1ok && x0ok)
                    red_add_f32x4(gvl + o00 + ys + c, w10 * tg[c], w10 * tg[c + 1], w10 * tg[c + 2], w10 * tg[c + 3]);
                if (y1ok && x1ok)
                    red_add_f32x4(gvl + o00 + ys + xs + c, w11 * tg[c], w11 * tg[c + 1], w11 * tg[c + 2], w11 * tg[c + 3]);
            }
        }
        group_reduce_scatter<G, 2 * P>(pxy, gl);
        group_reduce_scatter<G, P>(pa, gl);
        if (w.valid) {
            group_store<G, 2 * P>(glp, pxy, gl);
            group_store<G, P>(gap, pa, gl);
        }
        glp += 2 * P;
        gap += P;
    }
}

// ------------------------------------------------------------------------------------------------
// record kernel (second generation, see msda_records.cuh).  Per batch of G samples:
//   1. lane s of a lane group builds the geometry record of sample s (kept privately as well);
//   2. every lane walks the G records: 4 corner loads of its 4 channels, 4 partial dot products
//      t_ij = sum_c g_c * v_ij,c  and one REDG.E.ADD.F32x4 per contributing corner ((w_ij*a) * g);
//   3. a butterfly reduce-scatter hands lane s the four totals t_ij of ITS sample; that lane turns
//      them into grad_attn = sum w_ij t_ij and grad_loc = (W*a*(hy(t01-t00)+ly(t11-t10)),
//      H*a*(hx(t10-t00)+lx(t11-t01)))  (ms_deform_im2col_cuda.cuh:119-158) and writes them,
//      overwriting every element exactly once (zeros for outside samples, cuh:365-367).
// ------------------------------------------------------------------------------------------------
// FUSED: `loc` = raw sampling offsets, `attn` = raw logits, `ref` = (N,Lq,L,2)
// reference points; `grad_loc` receives d/d offsets = grad_loc / (W,H) and `grad_attn` receives
// d/d logits = a * (grad_a - sum_j a_j grad_a_j)  (softmax backward over the L*P samples).
template <typename VT, int D, int MINB, bool FUSED = false>
__global__ void __launch_bounds__(256, MINB)
bwd_rec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, const VT *__restrict__ grad_out,
               float *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_attn, const Dims d, const int order,
               const float *__restrict__ ref = nullptr)
{
    constexpr int G = D / kChannelsPerLane;
    using RL = RecordLayout<G>;
    constexpr int QPW = RL::QPW;
    static_assert(G >= 2 && G <= 32 && (32 % G) == 0, "unsupported D");

    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    __shared__ __align__(16) uint32_t s_rec[8 * RL::WARP_WORDS];
    stage_levels(s_lv, shapes, lsi, d.L);

    const int lane = threadIdx.x & 31;
    const int gl = lane % G, k = lane / G;
    WorkItem w = decode_work<QPW>(d, order, k);
    if (__ballot_sync(kFullMask, w.valid) == 0) return;
    if (!w.valid) { w.n = 0; w.q = 0; w.m = 0; }

    const int LP = d.L * d.P;
    const long qm = ((long)w.n * d.Lq + w.q) * d.M + w.m;
    const long img = ((long)w.n * d.S * d.M + w.m) * D + gl * kChannelsPerLane;
    const VT *vimg = value + img;
    float *gvimg = grad_value + img;
    const int xs = d.M * D;
    uint32_t *grp = s_rec + (threadIdx.x >> 5) * RL::WARP_WORDS + k * RL::GROUP_WORDS;

    float g[4];
    Vec4<VT>::load(grad_out + qm * D + gl * kChannelsPerLane, g);

    float aw[kMaxBatches];                          // FUSED: softmax weights of this lane's samples ...
    float pa[kMaxBatches], pg[kMaxBatches];         // ... and, per processed batch, (a, d out / d a)
    if constexpr (FUSED) {
        group_softmax<G>(attn, qm * LP, LP, gl, w.valid, aw);
#pragma unroll
        for (int b = 0; b < kMaxBatches; ++b) pa[b] = pg[b] = 0.f;
    }
    auto fetch = [&](int sidx) -> SampleIn {
        const bool has = w.valid && sidx < LP;
        if constexpr (FUSED) {
            const int l = has ? sidx / d.P : 0;
            const SampleIn r = fetch_sample_fused(has, loc, ref, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv, l, aw[0]);
            aw[0] = aw[1]; aw[1] = aw[2]; aw[2] = aw[3];
            return r;
        } else {
            return fetch_sample(has, loc, attn, qm * LP + sidx);
        }
    };
    SampleIn in = fetch(gl);
    for (int b0 = 0; b0 < LP; b0 += G) {
        const int sidx = b0 + gl;
        const bool has = w.valid && sidx < LP;
        const SampleGeom gm = build_record(grp + gl * 4, grp + RL::WEIGHTS + gl * 4, has && d.S > 0, in, s_lv,
                                           sidx / d.P, xs);
        const float a_cur = in.a;
        __syncwarp();
        in = fetch(sidx + G);                                                              // next batch, in flight

        float t[4 * G];
#pragma unroll
        for (int s = 0; s < G; ++s) {
            t[4 * s] = t[4 * s + 1] = t[4 * s + 2] = t[4 * s + 3] = 0.f;
            const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
            const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
            if (d.S > 0) {
                float v00[4], v01[4], v10[4], v11[4];
                Vec4<VT>::load(vimg + off.x, v00);
                Vec4<VT>::load(vimg + off.y, v01);
                Vec4<VT>::load(vimg + off.z, v10);
                Vec4<VT>::load(vimg + off.w, v11);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    t[4 * s] += g[c] * v00[c];
                    t[4 * s + 1] += g[c] * v01[c];
                    t[4 * s + 2] += g[c] * v10[c];
                    t[4 * s + 3] += g[c] * v11[c];
                }
            }
            // a zero weight contributes nothing to grad_value (corner outside the map, sample outside
            // the window, a == 0, or an exactly integral coordinate): skip the reduction -- the
            // SM->L2 reduction port is the scarce resource of this kernel
            if (wa.x != 0.f) red_add_f32x4(gvimg + off.x, wa.x * g[0], wa.x * g[1], wa.x * g[2], wa.x * g[3]);
            if (wa.y != 0.f) red_add_f32x4(gvimg + off.y, wa.y * g[0], wa.y * g[1], wa.y * g[2], wa.y * g[3]);
            if (wa.z != 0.f) red_add_f32x4(gvimg + off.z, wa.z * g[0], wa.z * g[1], wa.z * g[2], wa.z * g[3]);
            if (wa.w != 0.f) red_add_f32x4(gvimg + off.w, wa.w * g[0], wa.w * g[1], wa.w * g[2], wa.w * g[3]);
        }
        __syncwarp();

        group_reduce_scatter<G, 4 * G>(t, gl);          // lane s now holds t00,t01,t10,t11 of sample s
        if (has) {
            float gx = 0.f, gy = 0.f, ga = 0.f;
            if (gm.live) {
                // corners outside the map were read from a clamped address; the reference counts 0
                const float t00 = (gm.vmask & 1u) ? t[0] : 0.f, t01 = (gm.vmask & 2u) ? t[1] : 0.f;
                const float t10 = (gm.vmask & 4u) ? t[2] : 0.f, t11 = (gm.vmask & 8u) ? t[3] : 0.f;
                ga = gm.w00 * t00 + gm.w01 * t01 + gm.w10 * t10 + gm.w11 * t11;
                gx = gm.Wf * gm.a * (gm.hy * (t01 - t00) + gm.ly * (t11 - t10));
                gy = gm.Hf * gm.a * (gm.hx * (t10 - t00) + gm.lx * (t11 - t01));
            }
            const long si = qm * LP + sidx;
            if constexpr (FUSED) {
                // d loc / d offset = 1 / (W, H); outside samples have gx = gy = 0 (Wf = Hf = 0 there)
                const float ox = gm.live ? __fdiv_rn(gx, gm.Wf) : 0.f, oy = gm.live ? __fdiv_rn(gy, gm.Hf) : 0.f;
                *reinterpret_cast<float2 *>(grad_loc + 2 * si) = make_float2(ox, oy);
                pa[0] = pa[1]; pa[1] = pa[2]; pa[2] = pa[3]; pa[3] = a_cur;       // queue of (a, grad_a), newest last
                pg[0] = pg[1]; pg[1] = pg[2]; pg[2] = pg[3]; pg[3] = ga;
            } else {
                *reinterpret_cast<float2 *>(grad_loc + 2 * si) = make_float2(gx, gy);
                grad_attn[si] = ga;
            }
        } else if constexpr (FUSED) {
            pa[0] = pa[1]; pa[1] = pa[2]; pa[2] = pa[3]; pa[3] = 0.f;
            pg[0] = pg[1]; pg[1] = pg[2]; pg[2] = pg[3]; pg[3] = 0.f;
        }
    }
    if constexpr (FUSED) {
        // softmax backward: grad_logit_i = a_i * (g_i - sum_j a_j g_j); the queue holds the last
        // nb = ceil(LP / G) batches at positions kMaxBatches - nb ... kMaxBatches - 1
        float dot = 0.f;
#pragma unroll
        for (int b = 0; b < kMaxBatches; ++b) dot += pa[b] * pg[b];
        dot = group_allreduce_sum<G>(dot);
        const int nb = (LP + G - 1) / G;
#pragma unroll
        for (int b = 0; b < kMaxBatches; ++b) {
            const int sidx = (b - (kMaxBatches - nb)) * G + gl;
            if (b >= kMaxBatches - nb && w.valid && sidx < LP) grad_attn[qm * LP + sidx] = pa[b] * (pg[b] - dot);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic kernel: one warp per (n, q, m); lanes stride over channels; any D / P / alignment.
// VT value & grad_out type, CT arithmetic / loc / attn / all-gradients type.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_ct(const float *p, float) { return *p; }
__device__ __forceinline__ double ld_ct(const double *p, double) { return *p; }
__device__ __forceinline__ float ld_ct(const __nv_bfloat16 *p, float) { return __bfloat162float(*p); }

template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
bwd_generic_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                   const int64_t *__restrict__ lsi, const CT *__restrict__ loc,
                   const CT *__restrict__ attn, const VT *__restrict__ grad_out,
                   CT *__restrict__ grad_value, CT *__restrict__ grad_loc,
                   CT *__restrict__ grad_attn, const Dims d)
{
    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    stage_levels(s_lv, shapes, lsi, d.L);

    const int lane = threadIdx.x & 31;
    const long qm = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qm >= (long)d.N * d.Lq * d.M) return;          // whole warp leaves together
    const int m = (int)(qm % d.M);
    const long n = qm / ((long)d.Lq * d.M);

    const long xs = (long)d.M * d.D;
    const long img = (n * d.S * d.M + m) * (long)d.D;
    const VT *gq = grad_out + qm * d.D;
    const CT *lp = loc + qm * (long)d.L * d.P * 2;
    const CT *ap = attn + qm * (long)d.L * d.P;
    CT *glp = grad_loc + qm * (long)d.L * d.P * 2;
    CT *gap = grad_attn + qm * (long)d.L * d.P;

    for (int l = 0; l < d.L; ++l) {
        const LevelInfo li = s_lv[l];
        const long base = img + (long)li.start * xs;
        const long ys = (long)li.W * xs;
        for (int p = 0; p < d.P; ++p, lp += 2, ++ap, glp += 2, ++gap) {
            const Tap<CT> t = make_tap(lp[0], lp[1], li.H, li.W);
            CT sx = 0, sy = 0, sa = 0;
            if (t.inside) {
                const bool y0ok = t.y0 >= 0, y1ok = t.y0 + 1 <= li.H - 1;
                const bool x0ok = t.x0 >= 0, x1ok = t.x0 + 1 <= li.W - 1;
                const long o00 = base + t.y0 * ys + t.x0 * xs;
                const CT hy = CT(1) - t.ly, hx = CT(1) - t.lx;
                const CT w00 = hy * hx, w01 = hy * t.lx, w10 = t.ly * hx, w11 = t.ly * t.lx;
                const CT a = ap[0];
                for (int c = lane; c < d.D; c += 32) {
                    const CT gc = ld_ct(gq + c, CT());
                    const CT tg = gc * a;
                    CT v00 = 0, v01 = 0, v10 = 0, v11 = 0;
                    if (y0ok && x0ok) { v00 = ld_ct(value + o00 + c, CT()); atomicAdd(grad_value + o00 + c, w00 * tg); }
                    if (y0ok && x1ok) { v01 = ld_ct(value + o00 + xs + c, CT()); atomicAdd(grad_value + o00 + xs + c, w01 * tg); }
                    if (y1ok && x0ok) { v10 = ld_ct(value + o00 + ys + c, CT()); atomicAdd(grad_value + o00 + ys + c, w10 * tg); }
                    if (y1ok && x1ok) { v11 = ld_ct(value + o00 + ys + xs + c, CT()); atomicAdd(grad_value + o00 + ys + xs + c, w11 * tg); }
                    sx += (hy * (v01 - v00) + t.ly * (v11 - v10)) * tg;
                    sy += (hx * (v10 - v00) + t.lx * (v11 - v01)) * tg;
                    sa += gc * (w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11);
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                sx += __shfl_xor_sync(kFullMask, sx, off);
                sy += __shfl_xor_sync(kFullMask, sy, off);
                sa += __shfl_xor_sync(kFullMask, sa, off);
            }
            if (lane == 0) {
                glp[0] = (CT)li.W * sx;
                glp[1] = (CT)li.H * sy;
                gap[0] = sa;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
namespace {

template <typename VT>
constexpr bool vec_supported(int D, int P)
{
    if (P != 4) return false;
    if (sizeof(VT) == 4) return D == 16 || D == 32 || D == 64;
    return D == 32 || D == 64;
}

// Kernel choice (bwd_variant): -1 = measured default, 0/1 = vector kernel with work order 0/1,
// 10/11 = record kernel with order 0/1, 99 = generic.
// Default, measured on B200 at configs[1] (profiles/r01_v2_sweep.jsonl): the backward is bound by
// the SM->L2 reduction port, not by instruction issue, so for fp32 the older vector kernel (64
// registers, 4 CTAs/SM) still edges out the record kernel (80 registers, 3 CTAs/SM): 1.56 vs
// 1.61 ms.  For bf16 the record kernel wins clearly (1.56 vs 2.77 ms) because its 8-lane groups
// emit full 128-byte reduction lines.
template <typename VT>
bool use_vec(const Dims &d, bool vec_ok)
{
    const int v = tuning().bwd_variant;
    const bool pick = (v == 0 || v == 1) || (v < 0 && sizeof(VT) == 4);
    return vec_ok && pick && vec_supported<VT>(d.D, d.P) && (long)d.S * d.M * d.D < (1L << 31);
}

template <typename VT>
bool use_rec(const Dims &d, bool vec_ok)
{
    const int v = tuning().bwd_variant;
    const bool pick = (v == 10 || v == 11) || (v < 0 && !use_vec<VT>(d, vec_ok));
    return vec_ok && pick && (d.D == 16 || d.D == 32 || d.D == 64) && (long)d.S * d.M * d.D < (1L << 31);
}

template <typename VT, int D>
int run_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
            const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
            cudaStream_t st)
{
    constexpr int QPW = 32 / (D / kChannelsPerLane);
    const int threads = 256;                       // s_rec is sized for 8 warps
    const int order = tuning().bwd_variant == 10 ? 0 : 1;
    const long grid = grid_for(d, order, QPW, threads);
    // bwd_pipe = requested minimum CTAs/SM (register cap 64K / (256 * MINB)); trades ILP for TLP
#define MSDA_BWD_REC(MINB)                                                                                  \
    bwd_rec_kernel<VT, D, MINB><<<(unsigned)grid, threads, 0, st>>>(                                        \
        (const VT *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const VT *)grad_out,      \
        (float *)gv, (float *)gl, (float *)ga, d, order)
    switch (tuning().bwd_pipe) {
    case 1: MSDA_BWD_REC(1); break;
    case 2: MSDA_BWD_REC(2); break;
    case 4: MSDA_BWD_REC(4); break;
    default: if (D <= 32) MSDA_BWD_REC(3); else MSDA_BWD_REC(1); break;
    }
#undef MSDA_BWD_REC
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT>
int dispatch_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                 const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
                 cudaStream_t st)
{
    switch (d.D) {
    case 16: return run_rec<VT, 16>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, st);
    case 32: return run_rec<VT, 32>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, st);
    case 64: return run_rec<VT, 64>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, st);
    }
    return (int)cudaErrorInvalidValue;
}

template <typename VT, int D, int P>
int run_vec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
            const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
            cudaStream_t st)
{
    constexpr int QPW = 32 / (D / Vec<VT>::N);
    const int threads = tuning().block_threads > 0 ? tuning().block_threads : 256;
    const int order = tuning().bwd_variant == 0 ? 0 : 1;
    const long grid = grid_for(d, order, QPW, threads);
    bwd_vec_kernel<VT, D, P><<<(unsigned)grid, threads, 0, st>>>(
        (const VT *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const VT *)grad_out,
        (float *)gv, (float *)gl, (float *)ga, d, order);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT, typename CT>
int run_generic(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
                cudaStream_t st)
{
    const long warps = (long)d.N * d.Lq * d.M;
    const int threads = 256;
    const long grid = (warps + (threads / 32) - 1) / (threads / 32);
    bwd_generic_kernel<VT, CT><<<(unsigned)grid, threads, 0, st>>>(
        (const VT *)value, shapes, lsi, (const CT *)loc, (const CT *)attn, (const VT *)grad_out,
        (CT *)gv, (CT *)gl, (CT *)ga, d);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace

namespace {
template <typename VT, int D>
int run_rec_fused(const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref, const void *offsets,
                  const void *logits, const void *grad_out, void *gv, void *goff, void *glogit, const Dims &d,
                  cudaStream_t st)
{
    constexpr int G = D / kChannelsPerLane;
    const long grid = grid_for(d, 1, 32 / G, 256);
#define MSDA_BWD_FUSED(MINB)                                                                                         \
    bwd_rec_kernel<VT, D, MINB, true><<<(unsigned)grid, 256, 0, st>>>(                                               \
        (const VT *)value, shapes, lsi, (const float *)offsets, (const float *)logits, (const VT *)grad_out, (float *)gv, \
        (float *)goff, (float *)glogit, d, 1, (const float *)ref)
    if (D > 32) MSDA_BWD_FUSED(1);
    else if (tuning().bwd_pipe == 2) MSDA_BWD_FUSED(2);
    else MSDA_BWD_FUSED(3);
#undef MSDA_BWD_FUSED
    count_launch();
    return (int)cudaGetLastError();
}
}  // namespace

int launch_backward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                          const void *offsets, const void *logits, const void *grad_out, void *gv, void *goff,
                          void *glogit, const Dims &d, cudaStream_t st)
{
    if ((long)d.S * d.M * d.D >= (1L << 31) || dt == DType::F64) return kUnsupported;
    if (!(d.D == 16 || d.D == 32 || d.D == 64) || d.L * d.P > kMaxBatches * (d.D / kChannelsPerLane)) return kUnsupported;
    const size_t gv_bytes = 4 * (size_t)d.N * d.S * d.M * d.D;
    if (gv_bytes) {
        cudaError_t e = cudaMemsetAsync(gv, 0, gv_bytes, st);
        if (e != cudaSuccess) return (int)e;
    }
    if ((long)d.N * d.Lq * d.M == 0) return 0;
#define FUSED_ARGS value, shapes, lsi, ref, offsets, logits, grad_out, gv, goff, glogit, d, st
    if (dt == DType::F32) {
        switch (d.D) {
        case 16: return run_rec_fused<float, 16>(FUSED_ARGS);
        case 32: return run_rec_fused<float, 32>(FUSED_ARGS);
        case 64: return run_rec_fused<float, 64>(FUSED_ARGS);
        }
    } else {
        switch (d.D) {
        case 16: return run_rec_fused<__nv_bfloat16, 16>(FUSED_ARGS);
        case 32: return run_rec_fused<__nv_bfloat16, 32>(FUSED_ARGS);
        case 64: return run_rec_fused<__nv_bfloat16, 64>(FUSED_ARGS);
        }
    }
#undef FUSED_ARGS
    return kUnsupported;
}

const char *backward_kernel_name(DType dt, int D, int L, int P, bool vec_ok)
{
    (void)L;
    Dims d{1, 1, 1, D, 1, 1, P};
    switch (dt) {
    case DType::F64: return "bwd_generic_f64";
    case DType::F32:
        return use_rec<float>(d, vec_ok) ? "bwd_rec_f32" : (use_vec<float>(d, vec_ok) ? "bwd_vec_f32" : "bwd_generic_f32");
    case DType::BF16:
        return use_rec<__nv_bfloat16>(d, vec_ok) ? "bwd_rec_bf16"
                                                 : (use_vec<__nv_bfloat16>(d, vec_ok) ? "bwd_vec_bf16" : "bwd_generic_bf16");
    }
    return "?";
}

int launch_backward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                    const void *loc, const void *attn, const void *grad_out, void *gv, void *gl,
                    void *ga, const Dims &d, bool vec_ok, cudaStream_t st)
{
    const size_t gv_elt = (dt == DType::F64) ? 8 : 4;
    const size_t gv_bytes = gv_elt * (size_t)d.N * d.S * d.M * d.D;
    if (gv_bytes) {
        cudaError_t e = cudaMemsetAsync(gv, 0, gv_bytes, st);      // ms_deform_attn_cuda.cu:121
        if (e != cudaSuccess) return (int)e;
    }
    if ((long)d.N * d.Lq * d.M == 0) return 0;
#define ARGS value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, st
    if (dt == DType::F64) return run_generic<double, double>(ARGS);
    if (dt == DType::F32 && use_rec<float>(d, vec_ok)) return dispatch_rec<float>(ARGS);
    if (dt == DType::BF16 && use_rec<__nv_bfloat16>(d, vec_ok)) return dispatch_rec<__nv_bfloat16>(ARGS);
    if (dt == DType::F32) {
        if (use_vec<float>(d, vec_ok)) {
            switch (d.D) {
            case 16: return run_vec<float, 16, 4>(ARGS);
            case 32: return run_vec<float, 32, 4>(ARGS);
            case 64: return run_vec<float, 64, 4>(ARGS);
            }
        }
        return run_generic<float, float>(ARGS);
    }
    if (use_vec<__nv_bfloat16>(d, vec_ok)) {
        switch (d.D) {
        case 32: return run_vec<__nv_bfloat16, 32, 4>(ARGS);
        case 64: return run_vec<__nv_bfloat16, 64, 4>(ARGS);
        }
    }
    return run_generic<__nv_bfloat16, float>(ARGS);
#undef ARGS
}

}  // namespace msda
