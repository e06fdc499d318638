// msda_forward.cu -- MSDA forward kernels for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel (reference
// MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299) and its
// launcher (:923-954).  Same arithmetic per sample (see msda_common.cuh), different machine
// mapping:
//   * one lane owns 4 channels (one LDG.E.128 per corner in fp32, LDG.E.64 in bf16) and D/4 lanes
//     (a "lane group") cover one (query, head); a warp owns 32/G consecutive queries of one head:
//     neighbouring queries gather neighbouring pixels, so their corner lines coincide in L1;
//   * the bilinear geometry of a sample is computed ONCE, by one lane of the group, and shared
//     through a 32-byte shared-memory record (msda_records.cuh) -- the first-generation kernel
//     recomputed it in all 8 lanes and was instruction-issue bound (profiles/r01_v1_*);
//   * samples outside the sampling window are compacted away before the gather loop.
// The op is a gather: no tensor cores; the bound is L1 data-pipe wavefronts.
#include "msda_common.cuh"
#include "msda_records.cuh"

namespace msda {

// ------------------------------------------------------------------------------------------------
// record kernel (see msda_records.cuh): geometry computed once per sample by
// one lane, shared through shared memory; any L and P; D in {16, 32, 64}; fp32 or bf16 values.
// ------------------------------------------------------------------------------------------------
// FUSED: `loc` holds the raw sampling offsets, `attn` the raw attention logits and
// `ref` the (N,Lq,L,2) reference points; locations and softmax weights are formed in registers.
template <typename VT, int D, int MINB, bool COMPACT, bool FUSED = false, int LOADH = 0>
__global__ void __launch_bounds__(256, MINB)
fwd_rec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, VT *__restrict__ out, const Dims d, const int order,
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
    const VT *vimg = value + ((long)w.n * d.S * d.M + w.m) * D + gl * kChannelsPerLane;
    const int xs = d.M * D;
    uint32_t *grp = s_rec + (threadIdx.x >> 5) * RL::WARP_WORDS + k * RL::GROUP_WORDS;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (d.S > 0) {
        float aw[kMaxBatches];                      // FUSED: softmax weights of this lane's samples
        if constexpr (FUSED) group_softmax<G>(attn, qm * LP, LP, gl, w.valid, aw);
        auto fetch = [&](int sidx) -> SampleIn {
            const bool has = w.valid && sidx < LP;
            if constexpr (FUSED) {
                const int l = has ? sidx / d.P : 0;
                const SampleIn r = fetch_sample_fused(has, loc, ref, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv, l, aw[0]);
                aw[0] = aw[1]; aw[1] = aw[2]; aw[2] = aw[3];          // aw[0] = weight of the next batch
                return r;
            } else {
                return fetch_sample(has, loc, attn, qm * LP + sidx);
            }
        };
        SampleIn in = fetch(gl);
        for (int b0 = 0; b0 < LP; b0 += G) {
            const int sidx = b0 + gl;
            if constexpr (COMPACT) {
                // live records only, packed to the front of the group's area: samples outside the
                // window cost neither loads nor FMAs (15 % of them at MonoDETR's shapes)
                __align__(16) uint32_t tmp[8];
                const SampleGeom gm = build_record(tmp, tmp + 4, w.valid && sidx < LP, in, s_lv, sidx / d.P, xs);
                const unsigned gmask = (__ballot_sync(kFullMask, gm.live) >> (k * G)) & ((G == 32) ? ~0u : ((1u << G) - 1u));
                const int slot = __popc(gmask & ((1u << gl) - 1u));
                const int cnt = __popc(gmask);
                if (gm.live) {
                    *reinterpret_cast<int4 *>(grp + slot * 4) = *reinterpret_cast<const int4 *>(tmp);
                    *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + slot * 4) = *reinterpret_cast<const float4 *>(tmp + 4);
                }
                __syncwarp();
                in = fetch(sidx + G);
                for (int s = 0; s < cnt; ++s) {
                    const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
                    const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                    float v00[4], v01[4], v10[4], v11[4];
                    Vec4<VT>::template gather<LOADH>(vimg + off.x, v00);
                    Vec4<VT>::template gather<LOADH>(vimg + off.y, v01);
                    Vec4<VT>::template gather<LOADH>(vimg + off.z, v10);
                    Vec4<VT>::template gather<LOADH>(vimg + off.w, v11);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[c] += wa.x * v00[c] + wa.y * v01[c] + wa.z * v10[c] + wa.w * v11[c];
                }
                __syncwarp();
                continue;
            }
            build_record(grp + gl * 4, grp + RL::WEIGHTS + gl * 4, w.valid && sidx < LP, in, s_lv, sidx / d.P, xs);
            __syncwarp();
            in = fetch(sidx + G);                                     // next batch, in flight
#pragma unroll
            for (int s = 0; s < G; ++s) {
                const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
                const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                float v00[4], v01[4], v10[4], v11[4];
                Vec4<VT>::template gather<LOADH>(vimg + off.x, v00);
                Vec4<VT>::template gather<LOADH>(vimg + off.y, v01);
                Vec4<VT>::template gather<LOADH>(vimg + off.z, v10);
                Vec4<VT>::template gather<LOADH>(vimg + off.w, v11);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    acc[c] += wa.x * v00[c] + wa.y * v01[c] + wa.z * v10[c] + wa.w * v11[c];
            }
            __syncwarp();
        }
    }
    if (w.valid) Vec4<VT>::store(out + qm * D + gl * kChannelsPerLane, acc);
}

// ------------------------------------------------------------------------------------------------
// generic kernel: any D / P / alignment; VT value type, CT coordinate + arithmetic type.
// One thread per output element, 64-bit indexing throughout.
// ------------------------------------------------------------------------------------------------
template <typename VT>
__device__ __forceinline__ float to_ct(VT v, float) { return (float)v; }
__device__ __forceinline__ double to_ct(double v, double) { return v; }
__device__ __forceinline__ float to_ct(__nv_bfloat16 v, float) { return __bfloat162float(v); }

template <typename VT, typename CT>
__device__ __forceinline__ VT from_ct(CT v);
template <> __device__ __forceinline__ float from_ct<float, float>(float v) { return v; }
template <> __device__ __forceinline__ double from_ct<double, double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_ct<__nv_bfloat16, float>(float v)
{
    return __float2bfloat16_rn(v);
}

template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
fwd_generic_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                   const int64_t *__restrict__ lsi, const CT *__restrict__ loc,
                   const CT *__restrict__ attn, VT *__restrict__ out, const Dims d)
{
    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    stage_levels(s_lv, shapes, lsi, d.L);

    const long total = (long)d.N * d.Lq * d.M * d.D;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % d.D);
    const long qm = idx / d.D;
    const int m = (int)(qm % d.M);
    const long n = qm / ((long)d.Lq * d.M);

    const long xs = (long)d.M * d.D;
    const VT *vimg = value + (n * d.S * d.M + m) * (long)d.D + c;
    const CT *lp = loc + qm * (long)d.L * d.P * 2;
    const CT *ap = attn + qm * (long)d.L * d.P;
    CT acc = 0;
    for (int l = 0; l < d.L; ++l) {
        const LevelInfo li = s_lv[l];
        const VT *vl = vimg + (long)li.start * xs;
        const long ys = (long)li.W * xs;
        for (int p = 0; p < d.P; ++p, lp += 2, ++ap) {
            const Tap<CT> t = make_tap(lp[0], lp[1], li.H, li.W);
            if (!t.inside) continue;
            const bool y0ok = t.y0 >= 0, y1ok = t.y0 + 1 <= li.H - 1;
            const bool x0ok = t.x0 >= 0, x1ok = t.x0 + 1 <= li.W - 1;
            const VT *c00 = vl + (t.y0 * ys + t.x0 * xs);
            const CT v00 = (y0ok && x0ok) ? to_ct(c00[0], CT()) : CT(0);
            const CT v01 = (y0ok && x1ok) ? to_ct(c00[xs], CT()) : CT(0);
            const CT v10 = (y1ok && x0ok) ? to_ct(c00[ys], CT()) : CT(0);
            const CT v11 = (y1ok && x1ok) ? to_ct(c00[ys + xs], CT()) : CT(0);
            const CT hy = CT(1) - t.ly, hx = CT(1) - t.lx;
            acc += (hy * hx * v00 + hy * t.lx * v01 + t.ly * hx * v10 + t.ly * t.lx * v11) * ap[0];
        }
    }
    out[idx] = from_ct<VT, CT>(acc);
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
namespace {

template <typename VT, int D>
int run_rec(const VT *value, const int64_t *shapes, const int64_t *lsi, const float *loc,
            const float *attn, VT *out, const Dims &d, int order, cudaStream_t st)
{
    constexpr int QPW = 32 / (D / kChannelsPerLane);
    const int threads = 256;                       // s_rec is sized for 8 warps
    const long grid = grid_for(d, order, QPW, threads);
    // fwd_pipe = requested minimum CTAs/SM (register cap 64K / (256 * MINB)); trades ILP for TLP
#define MSDA_FWD_REC(MINB, COMPACT) \
    fwd_rec_kernel<VT, D, MINB, COMPACT><<<(unsigned)grid, threads, 0, st>>>(value, shapes, lsi, loc, attn, out, d, order)
    // Default, measured on B200 at configs[1] (profiles/r01_v2_compact_sweep.jsonl, r01_loadhint_sweep.jsonl):
    // fp32 -- compacting loop at <= 48 registers (5 CTAs/SM) with L1::no_allocate gathers 0.529 ms (0.592 with
    // allocating loads: L1 fills compete with the gather for the data pipe, hits are still served);
    // bf16 -- unrolled loop at <= 40 registers with allocating loads 0.493 ms (hints make no difference there).
    int flavour = tuning().fwd_pipe;
    if (flavour < 0) flavour = sizeof(VT) == 4 ? 25 : 6;
    switch (flavour) {
    case 3: MSDA_FWD_REC(3, false); break;
    case 5: MSDA_FWD_REC(5, false); break;
    case 6: MSDA_FWD_REC(6, false); break;
    case 15: MSDA_FWD_REC(5, true); break;
#define MSDA_FWD_REC_H(MINB, COMPACT, H) \
    fwd_rec_kernel<VT, D, MINB, COMPACT, false, H><<<(unsigned)grid, threads, 0, st>>>(value, shapes, lsi, loc, attn, out, d, order)
    case 25: MSDA_FWD_REC_H(5, true, 1); break;
    case 35: MSDA_FWD_REC_H(5, true, 2); break;
    case 26: MSDA_FWD_REC_H(6, false, 1); break;
    case 36: MSDA_FWD_REC_H(6, false, 2); break;
    case 24: MSDA_FWD_REC_H(4, true, 1); break;
    case 27: MSDA_FWD_REC_H(6, true, 1); break;
#undef MSDA_FWD_REC_H
    case 14: MSDA_FWD_REC(4, true); break;
    case 16: MSDA_FWD_REC(6, true); break;
    default: MSDA_FWD_REC(4, false); break;
    }
#undef MSDA_FWD_REC
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT>
int dispatch_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                 const void *attn, void *out, const Dims &d, int order, cudaStream_t st)
{
    const VT *v = (const VT *)value;
    const float *lo = (const float *)loc, *at = (const float *)attn;
    VT *o = (VT *)out;
    switch (d.D) {
    case 16: return run_rec<VT, 16>(v, shapes, lsi, lo, at, o, d, order, st);
    case 32: return run_rec<VT, 32>(v, shapes, lsi, lo, at, o, d, order, st);
    case 64: return run_rec<VT, 64>(v, shapes, lsi, lo, at, o, d, order, st);
    }
    return (int)cudaErrorInvalidValue;
}

// fwd_variant: -1 default (record kernel, work order 1); 10/11 record kernel with order 0/1; 99 generic.
inline bool rec_supported(const Dims &d) { return d.D == 16 || d.D == 32 || d.D == 64; }
inline bool want_rec(int variant) { return variant != 99; }

template <typename VT, typename CT>
int run_generic(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                const void *attn, void *out, const Dims &d, cudaStream_t st)
{
    const long total = (long)d.N * d.Lq * d.M * d.D;
    const int threads = 256;
    const long grid = (total + threads - 1) / threads;
    fwd_generic_kernel<VT, CT><<<(unsigned)grid, threads, 0, st>>>(
        (const VT *)value, shapes, lsi, (const CT *)loc, (const CT *)attn, (VT *)out, d);
    count_launch();
    return (int)cudaGetLastError();
}

bool use_rec(const Dims &d, bool vec_ok)
{
    return vec_ok && want_rec(tuning().fwd_variant) && rec_supported(d) && (long)d.S * d.M * d.D < (1L << 31);
}

}  // namespace

namespace {
template <typename VT, int D>
int run_rec_fused(const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref, const void *offsets,
                  const void *logits, void *out, const Dims &d, cudaStream_t st)
{
    constexpr int G = D / kChannelsPerLane;
    if (d.L * d.P > kMaxBatches * G) return kUnsupported;
    const long grid = grid_for(d, 1, 32 / G, 256);
    if (sizeof(VT) == 4)
        fwd_rec_kernel<VT, D, 5, true, true, 1><<<(unsigned)grid, 256, 0, st>>>(
            (const VT *)value, shapes, lsi, (const float *)offsets, (const float *)logits, (VT *)out, d, 1, (const float *)ref);
    else
        fwd_rec_kernel<VT, D, 6, false, true><<<(unsigned)grid, 256, 0, st>>>(
            (const VT *)value, shapes, lsi, (const float *)offsets, (const float *)logits, (VT *)out, d, 1, (const float *)ref);
    count_launch();
    return (int)cudaGetLastError();
}
}  // namespace

int launch_forward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                         const void *offsets, const void *logits, void *out, const Dims &d, cudaStream_t st)
{
    if ((long)d.S * d.M * d.D >= (1L << 31) || dt == DType::F64) return kUnsupported;
    if (resident_forward_applies(d, dt, true)) return launch_forward_resident(dt, value, shapes, lsi, offsets, logits, out, d, ref, st);
#define FUSED_ARGS value, shapes, lsi, ref, offsets, logits, out, d, st
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

const char *forward_kernel_name(DType dt, int D, int L, int P, bool vec_ok)
{
    (void)L;
    Dims d{1, 1, 1, D, 1, 1, P};
    switch (dt) {
    case DType::F64: return "fwd_generic_f64";
    case DType::F32: return use_rec(d, vec_ok) ? "fwd_rec_f32" : "fwd_generic_f32";
    case DType::BF16: return use_rec(d, vec_ok) ? "fwd_rec_bf16" : "fwd_generic_bf16";
    }
    return "?";
}

int launch_forward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                   const void *loc, const void *attn, void *out, const Dims &d, bool vec_ok,
                   cudaStream_t st)
{
    if (dt == DType::F64) return run_generic<double, double>(value, shapes, lsi, loc, attn, out, d, st);
    if (resident_forward_applies(d, dt, vec_ok)) return launch_forward_resident(dt, value, shapes, lsi, loc, attn, out, d, nullptr, st);
    const int rec_order = tuning().fwd_variant == 10 ? 0 : 1;
    if (dt == DType::F32 && use_rec(d, vec_ok)) return dispatch_rec<float>(value, shapes, lsi, loc, attn, out, d, rec_order, st);
    if (dt == DType::BF16 && use_rec(d, vec_ok))
        return dispatch_rec<__nv_bfloat16>(value, shapes, lsi, loc, attn, out, d, rec_order, st);
    if (dt == DType::F32) return run_generic<float, float>(value, shapes, lsi, loc, attn, out, d, st);
    return run_generic<__nv_bfloat16, float>(value, shapes, lsi, loc, attn, out, d, st);
}

}  // namespace msda
