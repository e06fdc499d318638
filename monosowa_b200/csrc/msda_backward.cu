// msda_backward.cu -- MSDA backward kernels for sm_100a.
//
// Replaces the reference's seven col2im kernels and their dispatcher (reference
// MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920, 956-1327;
// gradient formulas :87-159).  Machine mapping (see also msda_forward.cu, msda_records.cuh):
//   * a lane owns 4 channels; D/4 lanes form the lane group of one (query, head); the bilinear
//     geometry of a sample is computed once by one lane and shared through a shared-memory record;
//   * grad_value: one REDG.E.ADD.F32x4 (red.global.add.v4.f32) per lane per corner -- 8 lanes emit one
//     full 128-byte reduction line -- instead of the reference's 4 scalar atomics per channel per sample;
//   * grad_sampling_loc / grad_attn_weight: per-lane partial dot products are combined with a
//     butterfly reduce-scatter over the lane group (warp shuffles only) and every element is written
//     exactly once -- no block barriers, no zero-fill (the reference: smem staging, 2 __syncthreads
//     and a serial D-way sum per sample point).
// The bound is the rate at which L2 absorbs reductions, ~49 G 128-byte lines per second chip-wide.
#include "msda_common.cuh"
#include "msda_records.cuh"

namespace msda {

// ------------------------------------------------------------------------------------------------
// record kernel (see msda_records.cuh).  Per batch of G samples:
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
// GVT: element type of grad_value.  float: REDG.E.ADD.F32x4 lines.  bf16 (short query sets with bf16 values):
// REDG.E.ADD.BF16x4 straight into the bf16 gradient -- a row receives a handful of contributions there, so the
// per-addition rounding stays within the bf16 tolerance (tests/test_msda_gpu.py) and neither an fp32 staging
// buffer nor a narrowing pass is needed.
// CPL: channels per lane.  4: D/4 lanes per (query, head), LDG.128 gathers.  8: D/8 lanes, one 256-bit (fp32) /
// 128-bit (bf16) gather per corner of 8 CONTIGUOUS channels -- the two record reads of a warp step then serve twice
// as many samples.  The reductions stay 16 bytes per lane (there is no 32-byte red): chunk j of a lane covers
// channels [j*4G + 4*gl, +4), so one REDG instruction still writes G*16 contiguous bytes of a row (whole 32-byte
// sectors: L2 reductions are sector-bound) and the lane keeps grad_out twice: g[] for the dot
// products (contiguous) and gr[] for the reductions (chunked).
template <typename VT, int D, int MINB, bool FUSED = false, typename GVT = float, int CPL = kChannelsPerLane>
__global__ void __launch_bounds__(256, MINB)
bwd_rec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, const VT *__restrict__ grad_out,
               GVT *__restrict__ grad_value, float *__restrict__ grad_loc,
               float *__restrict__ grad_attn, const Dims d, const int order,
               const float *__restrict__ ref = nullptr, const int ref_dim = 2)
{
    constexpr int LOADH = CPL == 4 ? 0 : 1;
    constexpr int G = D / CPL;
    constexpr int NCH = CPL / 4;                    // 16-byte reduction chunks per lane
    using RL = RecordLayout<G>;
    using V = VecN<VT, CPL>;
    constexpr int QPW = RL::QPW;
    static_assert(G >= 2 && G <= 32 && (32 % G) == 0 && CPL % 4 == 0, "unsupported D");

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
    const long img = ((long)w.n * d.S * d.M + w.m) * D;
    const VT *vimg = value + img + gl * CPL;
    GVT *gvimg = grad_value + img + gl * 4;
    const int xs = d.M * D;
    uint32_t *grp = s_rec + (threadIdx.x >> 5) * RL::WARP_WORDS + k * RL::GROUP_WORDS;

    float g[CPL], gr[NCH][4];
    V::load_stream(grad_out + qm * D + gl * CPL, g);
    if constexpr (NCH == 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) gr[0][c] = g[c];
    } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j) Vec4<VT>::load(grad_out + qm * D + j * 4 * G + gl * 4, gr[j]);
    }
    auto reduce_row = [&](int off, float wgt) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            red_add_x4(gvimg + off + j * 4 * G, wgt * gr[j][0], wgt * gr[j][1], wgt * gr[j][2], wgt * gr[j][3]);
    };

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
            const SampleIn r = fetch_sample_fused<stream_policy<G, VT>()>(has, loc, ref, ref_dim, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv, l,
                                                  d.P, aw[0]);
            aw[0] = aw[1]; aw[1] = aw[2]; aw[2] = aw[3];
            return r;
        } else {
            return fetch_sample<stream_policy<G, VT>()>(has, loc, attn, qm * LP + sidx);
        }
    };
    SampleIn in = fetch(gl);
    for (int b0 = 0; b0 < LP; b0 += G) {
        const int sidx = b0 + gl;
        const bool has = w.valid && sidx < LP;
        int4 roff;
        float4 rwa;
        const SampleGeom gm = sample_geometry(has && d.S > 0, in, s_lv, sidx / d.P, xs, roff, rwa);
        if (!gm.live) roff.x = -1;                               // consumers skip the gathers of this sample
        *reinterpret_cast<int4 *>(grp + gl * 4) = roff;
        *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + gl * 4) = rwa;
        const float a_cur = in.a, ex_cur = in.ex, ey_cur = in.ey;
        __syncwarp();
        in = fetch(sidx + G);                                                              // next batch, in flight

        // The G samples of the batch are consumed in two halves so that only 2*G partial dot products
        // are live at a time (t[4*G] costs 32 registers at G = 8 and caps the kernel at 3 CTAs/SM).
        // After the reduce-scatter of a half, lane j holds corner pair (j & 1) of the half's sample j / 2;
        // the owner of sample s then pulls its four totals from lanes 2*(s % (G/2)) and +1 of half s / (G/2).
        constexpr int GH = G / 2;
        float th[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float t[4 * GH];
#pragma unroll
            for (int u = 0; u < GH; ++u) {
                const int s = h * GH + u;
                t[4 * u] = t[4 * u + 1] = t[4 * u + 2] = t[4 * u + 3] = 0.f;
                const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
                const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                // a sample outside the window (or past L*P) reads nothing, like the reference's branch (cuh:288, 365-367):
                // its lanes are predicated off, so it costs no L1 wavefronts (15 % of the samples at MonoDETR's shapes)
                if (off.x >= 0) {
                    float v00[CPL], v01[CPL], v10[CPL], v11[CPL];
                    V::template gather<LOADH>(vimg + off.x, v00);
                    V::template gather<LOADH>(vimg + off.y, v01);
                    V::template gather<LOADH>(vimg + off.z, v10);
                    V::template gather<LOADH>(vimg + off.w, v11);
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        t[4 * u] += g[c] * v00[c];
                        t[4 * u + 1] += g[c] * v01[c];
                        t[4 * u + 2] += g[c] * v10[c];
                        t[4 * u + 3] += g[c] * v11[c];
                    }
                }
                // a zero weight contributes nothing to grad_value (corner outside the map, sample outside
                // the window, a == 0, or an exactly integral coordinate): skip the reduction -- the
                // reduction throughput of L2 is the scarce resource of this kernel
                if (wa.x != 0.f) reduce_row(off.x, wa.x);
                if (wa.y != 0.f) reduce_row(off.y, wa.y);
                if (wa.z != 0.f) reduce_row(off.z, wa.z);
                if (wa.w != 0.f) reduce_row(off.w, wa.w);
            }
            group_reduce_scatter<G, 4 * GH>(t, gl);
            th[h][0] = t[0];
            th[h][1] = t[1];
        }
        __syncwarp();

        float t[4];
        {
            const int src = (lane & ~(G - 1)) | (2 * (gl % GH));
            const bool second = gl >= GH;
            const float a0 = __shfl_sync(kFullMask, th[0][0], src), a1 = __shfl_sync(kFullMask, th[0][1], src);
            const float a2 = __shfl_sync(kFullMask, th[0][0], src + 1), a3 = __shfl_sync(kFullMask, th[0][1], src + 1);
            const float b0 = __shfl_sync(kFullMask, th[1][0], src), b1 = __shfl_sync(kFullMask, th[1][1], src);
            const float b2 = __shfl_sync(kFullMask, th[1][0], src + 1), b3 = __shfl_sync(kFullMask, th[1][1], src + 1);
            t[0] = second ? b0 : a0; t[1] = second ? b1 : a1; t[2] = second ? b2 : a2; t[3] = second ? b3 : a3;
        }
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
                // d loc / d offset; outside samples have gx = gy = 0
                const float2 go = gm.live ? fused_offset_grad(ref_dim, gx, gy, gm.Wf, gm.Hf, ex_cur, ey_cur, d.P)
                                          : make_float2(0.f, 0.f);
                __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), go);
                pa[0] = pa[1]; pa[1] = pa[2]; pa[2] = pa[3]; pa[3] = a_cur;       // queue of (a, grad_a), newest last
                pg[0] = pg[1]; pg[1] = pg[2]; pg[2] = pg[3]; pg[3] = ga;
            } else {
                __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), make_float2(gx, gy));
                __stcs(grad_attn + si, ga);
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
// fp32 accumulation buffer -> bf16 gradient (bf16 values whose scatter ran in fp32: the tile kernel, the
// generic kernel).  One pass, 32 bytes in / 16 bytes out per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
narrow_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, const long n)
{
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(src + i));
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(src + i) + 1);
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<const unsigned *>(&p0); o.y = *reinterpret_cast<const unsigned *>(&p1);
        o.z = *reinterpret_cast<const unsigned *>(&p2); o.w = *reinterpret_cast<const unsigned *>(&p3);
        *reinterpret_cast<uint4 *>(dst + i) = o;
    } else {
        for (long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
namespace {

// bwd_variant: -1 default (binned kernel for long query sets, else the record kernel); 11 record kernel always;
// 20 tile kernel always; 21 binned kernel always; 99 generic.
bool use_rec(const Dims &d, bool vec_ok)
{
    return vec_ok && tuning().bwd_variant != 99 && (d.D == 16 || d.D == 32 || d.D == 64) &&
           (long)d.S * d.M * d.D < (1L << 31);
}

template <typename VT, int D, bool FUSED, typename GVT, int CPL, int MINB>
int launch_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
               const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
               cudaStream_t st)
{
    constexpr int QPW = 32 / (D / CPL);
    const long grid = grid_for(d, 1, QPW, 256);
    if (grid > 0x7fffffffL) return kUnsupported;
    bwd_rec_kernel<VT, D, MINB, FUSED, GVT, CPL><<<(unsigned)grid, 256, 0, st>>>(
        (const VT *)value, shapes, lsi, (const float *)loc, (const float *)attn, (const VT *)grad_out, (GVT *)gv,
        (float *)gl, (float *)ga, d, 1, (const float *)ref, ref_dim);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT, int D, bool FUSED, typename GVT>
int run_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
            const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
            cudaStream_t st)
{
#define MSDA_BWD_REC(CPL, MINB) \
    launch_rec<VT, D, FUSED, GVT, CPL, MINB>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st)
    // 8 channels per lane (profiles/r02_bwd_rec_cpl8_interleaved.jsonl, configs[1] shapes, record kernels timed
    // alternately): D = 64 1.76 vs 2.23 ms -- 8 lanes per (query, head) like D = 32, one REDG instruction still
    // covers whole 128-byte rows; D = 32 2.02 vs 1.54 ms -- a REDG instruction of 4-lane groups touches 8 half rows,
    // and the reduction path charges per (instruction, row), not per byte: D = 32 keeps 4 channels per lane.
    if constexpr (D % 8 == 0 && D / 8 >= 4) {
        const bool ok = reinterpret_cast<uintptr_t>(value) % (8 * sizeof(VT)) == 0 &&
                        (!FUSED || d.L * d.P <= kMaxBatches * (D / 8));
#ifdef MSDA_AB
        if (ok && tuning().bwd_pipe == 82) return MSDA_BWD_REC(8, 2);
        if (ok && tuning().bwd_pipe == 83) return MSDA_BWD_REC(8, 3);
#endif
        if constexpr (D == 64) {                   // (D = 32 with 8 channels exists in the measurement build only)
            if (ok && tuning().bwd_pipe != 4) return MSDA_BWD_REC(8, 3);
        }
        (void)ok;
    }
    constexpr int MINB = D <= 32 ? 3 : 1;          // 80 registers, 3 CTAs/SM (profiles/r01b_sweep_binned_flavours.jsonl)
    return MSDA_BWD_REC(4, MINB);
#undef MSDA_BWD_REC
}

template <typename VT, bool FUSED, typename GVT>
int dispatch_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                 const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
                 cudaStream_t st)
{
    switch (d.D) {
    case 16: return run_rec<VT, 16, FUSED, GVT>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 32: return run_rec<VT, 32, FUSED, GVT>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 64: return run_rec<VT, 64, FUSED, GVT>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    }
    return kUnsupported;
}

template <typename VT, typename CT>
int run_generic(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
                cudaStream_t st)
{
    const long warps = (long)d.N * d.Lq * d.M;
    const int threads = 256;
    const long grid = (warps + (threads / 32) - 1) / (threads / 32);
    if (grid > 0x7fffffffL) return (int)cudaErrorInvalidConfiguration;
    bwd_generic_kernel<VT, CT><<<(unsigned)grid, threads, 0, st>>>(
        (const VT *)value, shapes, lsi, (const CT *)loc, (const CT *)attn, (const VT *)grad_out,
        (CT *)gv, (CT *)gl, (CT *)ga, d);
    count_launch();
    return (int)cudaGetLastError();
}

int narrow_to_bf16(const void *scratch, void *gv, size_t n, cudaStream_t st)
{
    if (n == 0) return 0;
    const long grid = (long)((n + 8 * 256 - 1) / (8 * 256));
    narrow_bf16_kernel<<<(unsigned)grid, 256, 0, st>>>((const float *)scratch, (__nv_bfloat16 *)gv, (long)n);
    count_launch();
    return (int)cudaGetLastError();
}

// the one backward launcher: `ref` != nullptr selects the fused pre-processing flavour
int backward_any(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref, int ref_dim,
                 const void *loc, const void *attn, const void *grad_out, void *gv, void *gl, void *ga, void *scratch,
                 const Dims &d, bool vec_ok, cudaStream_t st)
{
    const bool fused = ref != nullptr;
    const size_t n_gv = (size_t)d.N * d.S * d.M * d.D;
    const bool direct_bf16 = dt == DType::BF16 && !backward_needs_scratch(d, dt, vec_ok);
    if (dt == DType::BF16 && !direct_bf16 && n_gv && !scratch) return kNeedsScratch;
    void *acc = (dt == DType::BF16 && !direct_bf16) ? scratch : gv;         // where the scatter accumulates
    const size_t acc_elt = dt == DType::F64 ? 8 : (direct_bf16 ? 2 : 4);
    if (n_gv) {
        cudaError_t e = cudaMemsetAsync(acc, 0, acc_elt * n_gv, st);         // ms_deform_attn_cuda.cu:121
        if (e != cudaSuccess) return (int)e;
    }
    int rc = kUnsupported;
    if ((long)d.N * d.Lq * d.M == 0) {
        rc = 0;
    } else if (!direct_bf16 && tiled_backward_applies(d, dt, vec_ok)) {
        rc = launch_backward_tiled(dt, value, shapes, lsi, loc, attn, grad_out, acc, gl, ga, d, ref, ref_dim, st);
    } else if (!direct_bf16 && binned_backward_applies(d, dt, vec_ok)) {
        rc = launch_backward_binned(dt, value, shapes, lsi, loc, attn, grad_out, acc, gl, ga, d, ref, ref_dim, st);
    } else if (dt != DType::F64 && use_rec(d, vec_ok)) {
#define ARGS value, shapes, lsi, loc, attn, grad_out, acc, gl, ga, d, ref, ref_dim, st
        if (dt == DType::F32) rc = fused ? dispatch_rec<float, true, float>(ARGS) : dispatch_rec<float, false, float>(ARGS);
        else if (direct_bf16) rc = fused ? dispatch_rec<__nv_bfloat16, true, __nv_bfloat16>(ARGS)
                                         : dispatch_rec<__nv_bfloat16, false, __nv_bfloat16>(ARGS);
        else rc = fused ? dispatch_rec<__nv_bfloat16, true, float>(ARGS) : dispatch_rec<__nv_bfloat16, false, float>(ARGS);
#undef ARGS
    }
    if (rc == kUnsupported) {                                               // odd D, fp64, misaligned, grid overflow
        if (fused || direct_bf16) return kUnsupported;
#define ARGS value, shapes, lsi, loc, attn, grad_out, acc, gl, ga, d, st
        if (dt == DType::F64) rc = run_generic<double, double>(ARGS);
        else if (dt == DType::F32) rc = run_generic<float, float>(ARGS);
        else rc = run_generic<__nv_bfloat16, float>(ARGS);
#undef ARGS
    }
    if (rc == 0 && dt == DType::BF16 && !direct_bf16) rc = narrow_to_bf16(scratch, gv, n_gv, st);
    return rc;
}

}  // namespace

// bf16 values: does the scatter need the caller's fp32 accumulation buffer (then narrowed into the bf16
// gradient by one extra pass), or can it go straight into the bf16 gradient with REDG.E.ADD.BF16x4?
// Direct only for the record kernel when a row receives a handful of additions: every addition rounds to
// bf16, k of them cost about 1.1e-3 * sqrt((k + 1) / 2) relative error on top of the final rounding.  The host
// knows the AVERAGE number of contributions per row, Lq*L*P*4 / S; it must not exceed 4 (MonoDETR's decoder:
// 3.5 with 550 training queries, 0.3 with 50; the coarsest level then sees ~70 per row, measured rel-L2 of the
// whole gradient 3e-3 -- tests/test_msda_gpu.py::test_config2_decoder_bf16).  Anything denser accumulates in fp32.
bool backward_needs_scratch(const Dims &d, DType dt, bool vec_ok)
{
    if (dt != DType::BF16) return false;
    if (!use_rec(d, vec_ok)) return true;
    if (tuning().bf16_direct < 0 && (tiled_backward_applies(d, dt, vec_ok) || binned_backward_applies(d, dt, vec_ok))) return true;
    if (grid_for(d, 1, 32 / max(1, d.D / kChannelsPerLane), 256) > 0x7fffffffL) return true;
    const long max_adds = tuning().bf16_direct >= 0 ? tuning().bf16_direct : 4;
    return (long)d.Lq * d.L * d.P * 4 > max_adds * d.S;
}

int launch_backward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                          int ref_dim, const void *offsets, const void *logits, const void *grad_out, void *gv,
                          void *goff, void *glogit, void *scratch, const Dims &d, cudaStream_t st)
{
    if ((long)d.S * d.M * d.D >= (1L << 31) || dt == DType::F64) return kUnsupported;
    if (!(d.D == 16 || d.D == 32 || d.D == 64) || d.L * d.P > kMaxBatches * (d.D / kChannelsPerLane)) return kUnsupported;
    if (ref_dim != 2 && ref_dim != 6) return kUnsupported;
    return backward_any(dt, value, shapes, lsi, ref, ref_dim, offsets, logits, grad_out, gv, goff, glogit, scratch, d, true, st);
}

const char *backward_kernel_name(DType dt, const Dims &d, bool vec_ok)
{
    if (dt == DType::F64) return "bwd_generic_f64";
    const bool bf = dt == DType::BF16;
    if (tiled_backward_applies(d, dt, vec_ok)) return bf ? "bwd_tile_bf16" : "bwd_tile_f32";
    if (binned_backward_applies(d, dt, vec_ok)) return bf ? "bwd_bin_bf16" : "bwd_bin_f32";
    if (use_rec(d, vec_ok)) return bf ? "bwd_rec_bf16" : "bwd_rec_f32";
    return bf ? "bwd_generic_bf16" : "bwd_generic_f32";
}

int launch_backward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                    const void *loc, const void *attn, const void *grad_out, void *gv, void *gl,
                    void *ga, void *scratch, const Dims &d, bool vec_ok, cudaStream_t st)
{
    return backward_any(dt, value, shapes, lsi, nullptr, 2, loc, attn, grad_out, gv, gl, ga, scratch, d, vec_ok, st);
}

}  // namespace msda
