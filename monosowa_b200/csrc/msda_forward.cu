// msda_forward.cu -- MSDA forward kernels for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel (reference
// MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299) and its
// launcher (:923-954).  Same arithmetic per sample (see msda_common.cuh), different machine
// mapping:
//   * one lane owns 4 channels (one LDG.E.128 per corner in fp32, LDG.E.64 in bf16) and D/4 lanes
//     (a "lane group") cover one (query, head); a warp owns 32/G queries of one head that are
//     neighbours in the image, so their corner lines coincide in L1;
//   * the bilinear geometry of a sample is computed ONCE, by one lane of the group, and shared
//     through a 32-byte shared-memory record (msda_records.cuh) -- the first-generation kernel
//     recomputed it in all 8 lanes and was instruction-issue bound (profiles/r01_v1_*);
//   * samples outside the sampling window are skipped by predicate (they cost neither loads nor FMAs,
//     and -- like the reference's branch, cuh:288 -- never touch `value`);
//   * (measurement build only, -DMSDA_AB) fwd_tile_kernel: a persistent grid over 2-D image tiles (msda_tiles.cuh) --
//     built in round 2, parity-green, slower than the record kernel (see use_tile below, profiles/r02_tile_kernels.md).
// The op is a gather: no tensor cores; the bound is 128-byte rows through the L1 data pipe.
#include "msda_common.cuh"
#include "msda_records.cuh"
#include "msda_tiles.cuh"

namespace msda {

// ------------------------------------------------------------------------------------------------
// One (query, head) per lane group: the body shared by the record kernel (one pass per CTA) and the
// tile kernel (persistent CTAs).  `grp` = the lane group's record area in shared memory.
// FUSED: `loc` holds the raw sampling offsets, `attn` the raw attention logits and
// `ref` the (N,Lq,L,ref_dim) reference points; locations and softmax weights are formed in registers.
// ------------------------------------------------------------------------------------------------
// !COMPACT (shipped): all G records of a batch are walked by an unrolled loop and a record outside the window is
// skipped by predicate; COMPACT (measurement build): live records are packed to the front and walked with a runtime
// trip count.  Either way a sample outside the window reads nothing.
template <typename VT, int D, bool FUSED, int LOADH_, bool COMPACT = true, int CPL = kChannelsPerLane>
__device__ __forceinline__ void fwd_group(const VT *__restrict__ value, const float *__restrict__ loc,
                                          const float *__restrict__ attn, VT *__restrict__ out,
                                          const float *__restrict__ ref, const int ref_dim, const Dims &d,
                                          const LevelInfo *s_lv, uint32_t *grp, const bool valid, const int n,
                                          const int q, const int m, const int gl, const int k)
{
    constexpr int G = D / CPL;
    using RL = RecordLayout<G>;
    using V = VecN<VT, CPL>;
    // LOADH_ = gather policy (Vec4 / VecN::gather) + 16 * stream-policy override (0: by lane-group size, see
    // ld_stream2 in msda_records.cuh; k > 0 forces policy k - 1)
    constexpr int LOADH = LOADH_ % 16, SP = ((LOADH_ / 16) % 4) ? (LOADH_ / 16) % 4 - 1 : stream_policy<G, VT>();
    constexpr bool CHAIN = LOADH_ >= 64;            // A/B: acc = fma(w, v, acc) four times instead of acc += (sum of 4)
    const int LP = d.L * d.P;
    const long qm = ((long)n * d.Lq + q) * d.M + m;
    const VT *vimg = value + ((long)n * d.S * d.M + m) * D + gl * CPL;
    const int xs = d.M * D;

    float acc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
    if (d.S > 0) {
        float aw[kMaxBatches];                      // FUSED: softmax weights of this lane's samples
        if constexpr (FUSED) group_softmax<G>(attn, qm * LP, LP, gl, valid, aw);
        auto fetch = [&](int sidx) -> SampleIn {
            const bool has = valid && sidx < LP;
            if constexpr (FUSED) {
                const int l = has ? sidx / d.P : 0;
                const SampleIn r = fetch_sample_fused<SP>(has, loc, ref, ref_dim, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv,
                                                          l, d.P, aw[0]);
                aw[0] = aw[1]; aw[1] = aw[2]; aw[2] = aw[3];          // aw[0] = weight of the next batch
                return r;
            } else {
                return fetch_sample<SP>(has, loc, attn, qm * LP + sidx);
            }
        };
        SampleIn in = fetch(gl);
        for (int b0 = 0; b0 < LP; b0 += G) {
            const int sidx = b0 + gl;
            if constexpr (!COMPACT) {
                int4 roff;
                float4 rwa;
                const SampleGeom gm = sample_geometry(valid && sidx < LP, in, s_lv, sidx / d.P, xs, roff, rwa);
                if (!gm.live) roff.x = -1;
                *reinterpret_cast<int4 *>(grp + gl * 4) = roff;
                *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + gl * 4) = rwa;
                __syncwarp();
                in = fetch(sidx + G);                                 // next batch, in flight
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
                    if (off.x >= 0) {
                        const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                        float v00[CPL], v01[CPL], v10[CPL], v11[CPL];
                        V::template gather<LOADH>(vimg + off.x, v00);
                        V::template gather<LOADH>(vimg + off.y, v01);
                        V::template gather<LOADH>(vimg + off.z, v10);
                        V::template gather<LOADH>(vimg + off.w, v11);
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            if constexpr (CHAIN) {
                                acc[c] = __fmaf_rn(wa.x, v00[c], acc[c]);
                                acc[c] = __fmaf_rn(wa.y, v01[c], acc[c]);
                                acc[c] = __fmaf_rn(wa.z, v10[c], acc[c]);
                                acc[c] = __fmaf_rn(wa.w, v11[c], acc[c]);
                            } else {
                                acc[c] += wa.x * v00[c] + wa.y * v01[c] + wa.z * v10[c] + wa.w * v11[c];
                            }
                        }
                    }
                }
                __syncwarp();
                continue;
            }
            // live records only, packed to the front of the group's area
            __align__(16) uint32_t tmp[8];
            const SampleGeom gm = build_record(tmp, tmp + 4, valid && sidx < LP, in, s_lv, sidx / d.P, xs);
            const unsigned gmask = (__ballot_sync(kFullMask, gm.live) >> (k * G)) & ((G == 32) ? ~0u : ((1u << G) - 1u));
            const int slot = __popc(gmask & ((1u << gl) - 1u));
            const int cnt = __popc(gmask);
            if (gm.live) {
                *reinterpret_cast<int4 *>(grp + slot * 4) = *reinterpret_cast<const int4 *>(tmp);
                *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + slot * 4) = *reinterpret_cast<const float4 *>(tmp + 4);
            }
            __syncwarp();
            in = fetch(sidx + G);                                     // next batch, in flight
            for (int s = 0; s < cnt; ++s) {
                const int4 off = *reinterpret_cast<const int4 *>(grp + s * 4);
                const float4 wa = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                float v00[CPL], v01[CPL], v10[CPL], v11[CPL];
                V::template gather<LOADH>(vimg + off.x, v00);
                V::template gather<LOADH>(vimg + off.y, v01);
                V::template gather<LOADH>(vimg + off.z, v10);
                V::template gather<LOADH>(vimg + off.w, v11);
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    acc[c] += wa.x * v00[c] + wa.y * v01[c] + wa.z * v10[c] + wa.w * v11[c];
            }
            __syncwarp();
        }
    }
    if (valid) V::store(out + qm * D + gl * CPL, acc);
}

// ------------------------------------------------------------------------------------------------
// record kernel: one pass, a CTA's 8 warps cover 8 * 32/G consecutive queries of one head.
// Any L and P; D in {16, 32, 64}; fp32 or bf16 values.  Used for short query sets (the decoder).
// ------------------------------------------------------------------------------------------------
template <typename VT, int D, int MINB, bool FUSED = false, int LOADH = 0, bool COMPACT = true, int CPL = kChannelsPerLane>
__global__ void __launch_bounds__(256, MINB)
fwd_rec_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ lsi, const float *__restrict__ loc,
               const float *__restrict__ attn, VT *__restrict__ out, const Dims d, const int order,
               const float *__restrict__ ref = nullptr, const int ref_dim = 2)
{
    constexpr int G = D / CPL;
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
    uint32_t *grp = s_rec + (threadIdx.x >> 5) * RL::WARP_WORDS + k * RL::GROUP_WORDS;
    fwd_group<VT, D, FUSED, LOADH, COMPACT, CPL>(value, loc, attn, out, ref, ref_dim, d, s_lv, grp, w.valid, w.n, w.q, w.m, gl, k);
}

#ifdef MSDA_AB
// ------------------------------------------------------------------------------------------------
// tile kernel (measurement build only, see use_tile below): persistent CTAs walk (image, head, 2-D image tile) work items (msda_tiles.cuh); all warps
// of a CTA work on the same tile, pass after pass, so the rows they gather are each other's L1 hits.
// ------------------------------------------------------------------------------------------------
template <typename VT, int D, int THREADS, int MINB, bool FUSED = false, int LOADH = 0>
__global__ void __launch_bounds__(THREADS, MINB)
fwd_tile_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ lsi, const float *__restrict__ loc,
                const float *__restrict__ attn, VT *__restrict__ out, const Dims d,
                const float *__restrict__ ref, const int ref_dim)
{
    constexpr int G = D / kChannelsPerLane;
    using RL = RecordLayout<G>;
    constexpr int QPW = RL::QPW, WARPS = THREADS / 32;
    static_assert(G >= 2 && G <= 32 && (32 % G) == 0, "unsupported D");

    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    __shared__ TilePlan s_plan;
    __shared__ TileItem s_item;
    __shared__ __align__(16) uint32_t s_rec[WARPS * RL::WARP_WORDS];
    stage_levels(s_lv, shapes, lsi, d.L);
    if (threadIdx.x == 0) make_tile_plan(s_plan, s_lv, d, 256);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, k = lane / G;
    uint32_t *grp = s_rec + warp * RL::WARP_WORDS + k * RL::GROUP_WORDS;
    const long n_items = (long)d.N * d.M * s_plan.n_tiles;

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();                                        // everyone is done with the previous item
        if (threadIdx.x == 0) make_tile_item(s_item, s_plan, s_lv, d, item);
        __syncthreads();
        const int nq = s_item.nq, n = s_item.n, m = s_item.m;
        for (int i0 = warp * QPW; i0 < nq; i0 += WARPS * QPW) {
            const bool valid = i0 + k < nq;
            const int q = valid ? tile_query(s_item, s_plan, s_lv, d.L, i0 + k) : 0;
            fwd_group<VT, D, FUSED, LOADH>(value, loc, attn, out, ref, ref_dim, d, s_lv, grp, valid, n, q, m, gl, k);
        }
    }
}

#endif  // MSDA_AB

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

// fwd_variant: -1 / 11 record kernel; 12 tile kernel (opt-in, see use_tile); 99 generic.  fwd_pipe selects A/B launch flavours in -DMSDA_AB builds only.
inline bool rec_supported(const Dims &d) { return d.D == 16 || d.D == 32 || d.D == 64; }

bool use_rec(const Dims &d, bool vec_ok)
{
    return vec_ok && tuning().fwd_variant != 99 && rec_supported(d) && (long)d.S * d.M * d.D < (1L << 31);
}

// Measured on B200 at configs[1] (profiles/r02_sweep_tile_vs_rec.jsonl, r02_ncu_tile_v1_summary.md): the tile kernel
// raises the L1 hit rate from 0 to 70 % and cuts L2->SM traffic 3.5x, yet runs 0.64 ms against the record kernel's
// 0.53 ms -- both execute the same ~0.42 ms worth of L1 data-pipe wavefronts (a row costs a wavefront whether it
// hits L1 or arrives from L2) and the persistent loop hides latency worse.  The forward is bound by the number of
// rows through the pipe, not by where they come from; the record kernel stays the default for every Lq and the
// tile kernel exists only in the measurement build (-DMSDA_AB, fwd_variant = 12), where tests keep it covered.
bool use_tile(const Dims &d)
{
    (void)d;
#ifdef MSDA_AB
    return tuning().fwd_variant == 12;
#else
    return false;
#endif
}

template <typename VT, int D, bool FUSED, int CPL, int MINB, int LOADH, bool COMPACT>
int launch_rec(const VT *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn, VT *out,
               const Dims &d, const float *ref, int ref_dim, cudaStream_t st)
{
    constexpr int QPW = 32 / (D / CPL);
#ifdef MSDA_AB
    const int order = tuning().fwd_pipe == 41 ? 0 : 1;      // A/B: a CTA's warps walk the heads of one query chunk
#else
    const int order = 1;
#endif
    const long grid = grid_for(d, order, QPW, 256);
    if (grid > 0x7fffffffL) return kUnsupported;
    fwd_rec_kernel<VT, D, MINB, FUSED, LOADH, COMPACT, CPL><<<(unsigned)grid, 256, 0, st>>>(value, shapes, lsi, loc, attn, out, d, order, ref, ref_dim);
    count_launch();
    return (int)cudaGetLastError();
}

// Eight channels per lane (fp32: one 256-bit load per corner, bf16: one 128-bit load) when D allows at least four
// lanes per (query, head): a broadcast LDS.128 costs the L1 data pipe ~2.5 wavefronts whether it serves 4 or 8 lane
// groups (ncu), so a warp step that serves 32 / (D/8) samples instead of 32 / (D/4) halves the record share of the pipe
// (23.5 M -> 12.9 M wavefronts per launch at configs[1]) and the number of gather requests.
// The 256-bit load needs a 32-byte aligned base; rows are D * sizeof(VT) apart, a multiple of 32 B when D % 8 == 0.
template <typename VT, int D, bool FUSED>
bool wide_lanes(const VT *value, const Dims &d)
{
    if constexpr (D % 8 != 0 || D / 8 < 4) return false;
    else {
        if (tuning().fwd_pipe == 4) return false;                                  // knob: the 4-channel flavour
        if (reinterpret_cast<uintptr_t>(value) % (8 * sizeof(VT)) != 0) return false;
        if (FUSED && d.L * d.P > kMaxBatches * (D / 8)) return false;              // softmax batches per lane group
        return true;
    }
}

template <typename VT, int D, bool FUSED>
int run_rec(const VT *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn, VT *out,
            const Dims &d, const float *ref, int ref_dim, cudaStream_t st)
{
    // measured on B200 at configs[1], the flavours timed alternately (tools/ab_interleaved.py):
    //  * walk (profiles/r02_fwd_walks_interleaved.jsonl): the unrolled walk over all G records with outside samples
    //    skipped by predicate against the compacting loop (runtime trip count) -- bf16 0.511 vs 0.531 ms, fp32 a wash
    //    (0.551 vs 0.557 model-like, 0.602 vs 0.594 uniform, 0.461 vs 0.460 initial pattern) -- one walk for both;
    //  * channels per lane (profiles/r02_fwd_cpl8_interleaved.jsonl): 8 channels per lane at <= 64 registers / 4 CTAs
    //    per SM against 4 channels -- fp32 D=32 0.536 vs 0.562 ms (uniform 0.587 vs 0.607, initial pattern 0.445 vs
    //    0.469; 3 CTAs the same, 5 CTAs spill: 0.65), fp32 D=64 0.495 vs 0.530, bf16 D=32 0.430 vs 0.511, bf16 D=64
    //    0.361 vs 0.470;
    //  * gathers with L1::no_allocate (a strip of consecutive queries has little reuse, fills only compete with the
    //    gather for the data pipe): 8-channel fp32 0.63 ms with allocating loads, 8-channel bf16 0.439; the 4-channel
    //    bf16 flavour (D = 16, unaligned value) keeps allocating loads as measured in round 1.
#define MSDA_FWD_REC(CPL, MINB, LOADH, COMPACT) \
    launch_rec<VT, D, FUSED, CPL, MINB, LOADH, COMPACT>(value, shapes, lsi, loc, attn, out, d, ref, ref_dim, st)
#ifdef MSDA_AB
    if (tuning().fwd_pipe == 15)                 // A/B: the compacting loop
    {
        if constexpr (sizeof(VT) == 4) return MSDA_FWD_REC(4, 5, 1, true);
        else return MSDA_FWD_REC(4, 5, 0, true);
    }
    if constexpr (D % 8 == 0 && D / 8 >= 4) {
        if (wide_lanes<VT, D, FUSED>(value, d)) {
            if (tuning().fwd_pipe == 24) return MSDA_FWD_REC(8, 4, 0, false);   // A/B: allocating loads
            if constexpr (sizeof(VT) == 4) {
                if (tuning().fwd_pipe == 30) return MSDA_FWD_REC(8, 4, 3, false);        // A/B: L2::evict_last on the gathers
            }
            if (tuning().fwd_pipe == 31) return MSDA_FWD_REC(8, 4, 16 + 1, false);       // A/B: evict-first streams
            if (tuning().fwd_pipe == 33) return MSDA_FWD_REC(8, 4, 48 + 1, false);       // A/B: evict-first streams, L2::128B hint
            if (tuning().fwd_pipe == 40) return MSDA_FWD_REC(8, 4, 64 + 1, false);       // A/B: chained FMAs
            if (tuning().fwd_pipe == 27) return MSDA_FWD_REC(8, 5, 1, false);   // A/B: 5 CTAs per SM (48 registers, spills)
        }
    }
#endif
    if constexpr (D % 8 == 0 && D / 8 >= 4) {
        if (wide_lanes<VT, D, FUSED>(value, d)) {
            return MSDA_FWD_REC(8, 4, 1, false);
        }
    }
    if constexpr (sizeof(VT) == 4) return MSDA_FWD_REC(4, 5, 1, false);
    else return MSDA_FWD_REC(4, 6, 0, false);
#undef MSDA_FWD_REC
}

template <typename VT, int D, bool FUSED>
int run_tile(const VT *value, const int64_t *shapes, const int64_t *lsi, const float *loc, const float *attn, VT *out,
             const Dims &d, const float *ref, int ref_dim, cudaStream_t st)
{
#ifndef MSDA_AB
    (void)value; (void)shapes; (void)lsi; (void)loc; (void)attn; (void)out; (void)d; (void)ref; (void)ref_dim; (void)st;
    return kUnsupported;
#else
#define MSDA_FWD_TILE(THREADS, MINB, H)                                                                     \
    fwd_tile_kernel<VT, D, THREADS, MINB, FUSED, H><<<persistent_grid(MINB), THREADS, 0, st>>>(             \
        value, shapes, lsi, loc, attn, out, d, ref, ref_dim)
    switch (tuning().fwd_pipe) {
    case 1: MSDA_FWD_TILE(256, 5, 1); break;
    case 2: MSDA_FWD_TILE(512, 2, 0); break;
    case 3: MSDA_FWD_TILE(512, 2, 1); break;
    case 4: MSDA_FWD_TILE(1024, 1, 0); break;
    case 5: MSDA_FWD_TILE(1024, 1, 1); break;
    case 6: MSDA_FWD_TILE(256, 4, 0); break;
    case 7: MSDA_FWD_TILE(256, 6, 0); break;
    case 8: MSDA_FWD_TILE(512, 3, 0); break;
    default: MSDA_FWD_TILE(256, 5, 0); break;
    }
#undef MSDA_FWD_TILE
    count_launch();
    return (int)cudaGetLastError();
#endif
}

template <typename VT, bool FUSED>
int dispatch_rec(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                 void *out, const Dims &d, const void *ref, int ref_dim, cudaStream_t st)
{
    const VT *v = (const VT *)value;
    const float *lo = (const float *)loc, *at = (const float *)attn, *rf = (const float *)ref;
    VT *o = (VT *)out;
    const bool tile = use_tile(d);
#define MSDA_FWD_D(DD)                                                                            \
    case DD:                                                                                      \
        return tile ? run_tile<VT, DD, FUSED>(v, shapes, lsi, lo, at, o, d, rf, ref_dim, st)      \
                    : run_rec<VT, DD, FUSED>(v, shapes, lsi, lo, at, o, d, rf, ref_dim, st)
    switch (d.D) {
        MSDA_FWD_D(16);
        MSDA_FWD_D(32);
        MSDA_FWD_D(64);
    }
#undef MSDA_FWD_D
    return kUnsupported;
}

template <typename VT, typename CT>
int run_generic(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                const void *attn, void *out, const Dims &d, cudaStream_t st)
{
    const long total = (long)d.N * d.Lq * d.M * d.D;
    const int threads = 256;
    const long grid = (total + threads - 1) / threads;
    if (grid > 0x7fffffffL) return (int)cudaErrorInvalidConfiguration;
    fwd_generic_kernel<VT, CT><<<(unsigned)grid, threads, 0, st>>>(
        (const VT *)value, shapes, lsi, (const CT *)loc, (const CT *)attn, (VT *)out, d);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace

int launch_forward_fused(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *ref,
                         int ref_dim, const void *offsets, const void *logits, void *out, const Dims &d, cudaStream_t st)
{
    if ((long)d.S * d.M * d.D >= (1L << 31) || dt == DType::F64 || !rec_supported(d)) return kUnsupported;
    if (d.L * d.P > kMaxBatches * (d.D / kChannelsPerLane) || (ref_dim != 2 && ref_dim != 6)) return kUnsupported;
    if (dt == DType::F32) return dispatch_rec<float, true>(value, shapes, lsi, offsets, logits, out, d, ref, ref_dim, st);
    return dispatch_rec<__nv_bfloat16, true>(value, shapes, lsi, offsets, logits, out, d, ref, ref_dim, st);
}

const char *forward_kernel_name(DType dt, const Dims &d, bool vec_ok)
{
    if (dt == DType::F64) return "fwd_generic_f64";
    const bool bf = dt == DType::BF16;
    if (!use_rec(d, vec_ok)) return bf ? "fwd_generic_bf16" : "fwd_generic_f32";
    if (use_tile(d)) return bf ? "fwd_tile_bf16" : "fwd_tile_f32";
    return bf ? "fwd_rec_bf16" : "fwd_rec_f32";
}

int launch_forward(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi,
                   const void *loc, const void *attn, void *out, const Dims &d, bool vec_ok,
                   cudaStream_t st)
{
    if (dt == DType::F64) return run_generic<double, double>(value, shapes, lsi, loc, attn, out, d, st);
    if (use_rec(d, vec_ok)) {
        const int rc = dt == DType::F32 ? dispatch_rec<float, false>(value, shapes, lsi, loc, attn, out, d, nullptr, 2, st)
                                        : dispatch_rec<__nv_bfloat16, false>(value, shapes, lsi, loc, attn, out, d, nullptr, 2, st);
        if (rc != kUnsupported) return rc;                 // grid overflow: the generic kernel takes it
    }
    if (dt == DType::F32) return run_generic<float, float>(value, shapes, lsi, loc, attn, out, d, st);
    return run_generic<__nv_bfloat16, float>(value, shapes, lsi, loc, attn, out, d, st);
}

}  // namespace msda
