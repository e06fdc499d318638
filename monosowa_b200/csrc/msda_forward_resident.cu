// msda_forward_resident.cu -- MSDA forward for long query sets with the COARSE levels resident in shared memory.
//
// Why: the forward is bound by the LSU, and a gathered row costs more when it comes through
// LDG than through LDS.  One LDG.128 of the record kernel touches four different 128-byte lines (four queries x
// eight lanes); the L1TEX tag stage replays such an instruction once per extra line at ~2 cycles per line, which is
// the 1.8 cycles per row that a pure L2-resident row gather sustains (profiles/r01_ubench_gather_rows.txt) and that
// the record kernel runs at.  The same four rows read with one LDS.128 cost four conflict-free wavefronts: 1 cycle
// per row.  On a feature pyramid half of all gathered rows belong to the two coarsest levels, which are tiny
// (KITTI: 600 pixels per head = 77 KB in fp32), so a CTA that works through a long run of queries of one head
// can afford to copy them into shared memory once and serve half of its gathers from there.
//
// Which levels are resident is decided on the device from spatial_shapes (the host never reads them): the coarsest
// levels while their rows of one head fit kResBytes.  A record's four offsets carry a tag bit telling the consumer
// whether they index shared memory (row * D) or the image in global memory (row * M * D); all corners of a sample
// share the tag.  Otherwise the kernel is the compacting record kernel of msda_forward.cu (same arithmetic, same
// order of accumulation per query: bitwise identical outputs -- tested).
//
// RESULT (B200, configs[1], profiles/r01b_sweep_resident_forward.jsonl): 0.80 ms vs 0.53 ms for the record kernel --
// the hypothesis above is wrong where it matters.  An LDS row occupies the same L1 data-pipe wavefront as an LDG row
// (the pipe is the forward's binding unit at 79 %), the 97 KB per CTA cap the SM at 32 warps instead of 40, and 1280
// long CTAs leave a 14 % tail.  Kept as an opt-in A/B kernel (fwd_variant = 30), not used by default.
#include "msda_common.cuh"
#include "msda_records.cuh"

namespace msda {

namespace {

constexpr int kResThreads = 512;
constexpr int kResBytes = 80 * 1024;        // shared-memory budget for the resident levels of one head
constexpr int kResQueries = 1024;           // queries per CTA: the 77 KB copy is amortised over 1024 * L * P * 4 row reads
constexpr int kResMinQueries = 2048;
constexpr int kSmemTag = (int)0x80000000u;
constexpr bool kResidentForwardDefault = false;   // opt-in (fwd_variant = 30): measured SLOWER than the record kernel, see below

struct ResPlan {
    int lr;                                  // first resident level (levels lr..L-1 live in shared memory); L if none
    int rowbase[MSDA_MAX_LEVELS];            // first shared-memory row of each resident level
};

template <typename VT>
__device__ __forceinline__ void copy4(VT *dst, const VT *src);
template <>
__device__ __forceinline__ void copy4<float>(float *dst, const float *src)
{
    *reinterpret_cast<float4 *>(dst) = __ldg(reinterpret_cast<const float4 *>(src));
}
template <>
__device__ __forceinline__ void copy4<__nv_bfloat16>(__nv_bfloat16 *dst, const __nv_bfloat16 *src)
{
    *reinterpret_cast<uint2 *>(dst) = __ldg(reinterpret_cast<const uint2 *>(src));
}

template <typename VT>
__device__ __forceinline__ void load4_shared(const VT *p, float (&f)[4]);
template <>
__device__ __forceinline__ void load4_shared<float>(const float *p, float (&f)[4])
{
    const float4 v = *reinterpret_cast<const float4 *>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4_shared<__nv_bfloat16>(const __nv_bfloat16 *p, float (&f)[4])
{
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    f[0] = __uint_as_float(v.x << 16);
    f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16);
    f[3] = __uint_as_float(v.y & 0xffff0000u);
}

template <typename VT, int D, bool FUSED>
__global__ void __launch_bounds__(kResThreads, 2)
fwd_res_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ loc, const float *__restrict__ attn, VT *__restrict__ out, const Dims d,
               const float *__restrict__ ref)
{
    constexpr int G = D / kChannelsPerLane;
    using RL = RecordLayout<G>;
    constexpr int QPW = RL::QPW, WARPS = kResThreads / 32, QPI = WARPS * QPW;
    static_assert(kResQueries % QPI == 0, "chunk must be a whole number of passes");

    extern __shared__ __align__(16) unsigned char smem[];
    VT *s_val = reinterpret_cast<VT *>(smem);
    uint32_t *s_rec = reinterpret_cast<uint32_t *>(smem + kResBytes);
    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    __shared__ ResPlan s_plan;

    const int tid = threadIdx.x;
    stage_levels(s_lv, shapes, lsi, d.L);
    if (tid == 0) {
        ResPlan p;
        p.lr = d.L;
        long rows = 0;
        for (int l = d.L - 1; l >= 0; --l) {
            const int H = s_lv[l].H, W = s_lv[l].W;
            if (H < 0 || W < 0 || H > 8192 || W > 8192) break;
            if ((rows + (long)H * W) * D * (long)sizeof(VT) > kResBytes) break;
            rows += (long)H * W;
            p.lr = l;
        }
        int base = 0;
        for (int l = p.lr; l < d.L; ++l) {
            p.rowbase[l] = base;
            base += s_lv[l].H * s_lv[l].W;
        }
        s_plan = p;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int gl = lane % G, k = lane / G;
    const int n_chunks = (d.Lq + kResQueries - 1) / kResQueries;
    const int m = (int)(blockIdx.x % d.M);
    const long cr = blockIdx.x / d.M;
    const int q0 = (int)(cr % n_chunks) * kResQueries;
    const int n = (int)(cr / n_chunks);
    const int xs = d.M * D;
    const int lr = s_plan.lr;
    const VT *vhead = value + ((long)n * d.S * d.M + m) * D;          // pixel p of this head: vhead + p * xs

    // copy the resident levels of this (image, head) into shared memory, one 128-byte (fp32) row per lane group
    for (int l = lr; l < d.L; ++l) {
        const LevelInfo li = s_lv[l];
        const int px = li.H * li.W;
        VT *dst = s_val + (long)s_plan.rowbase[l] * D;
        const VT *src = vhead + (long)li.start * xs;
        for (int i = tid; i < px * G; i += kResThreads) {
            const int r = i / G, part = i % G;
            copy4<VT>(dst + r * D + part * kChannelsPerLane, src + (long)r * xs + part * kChannelsPerLane);
        }
    }
    __syncthreads();

    const int LP = d.L * d.P;
    const VT *vimg = vhead + gl * kChannelsPerLane;
    const VT *simg = s_val + gl * kChannelsPerLane;
    uint32_t *grp = s_rec + warp * RL::WARP_WORDS + k * RL::GROUP_WORDS;

    for (int it = 0; it < kResQueries / QPI; ++it) {
        const int qw = q0 + it * QPI + warp * QPW;
        if (qw >= d.Lq) break;                           // whole warp past the end
        const bool qvalid = qw + k < d.Lq;
        const long qm = ((long)n * d.Lq + (qvalid ? qw + k : q0)) * d.M + m;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (d.S > 0) {
            float aw[kMaxBatches];
            if constexpr (FUSED) group_softmax<G>(attn, qm * LP, LP, gl, qvalid, aw);
            auto fetch = [&](int sidx) -> SampleIn {
                const bool has = qvalid && sidx < LP;
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
                const bool has = qvalid && sidx < LP;
                const int l = has ? sidx / d.P : 0;
                const bool res = l >= lr;
                __align__(16) uint32_t tmp[8];
                const SampleGeom gm = build_record_at(tmp, tmp + 4, has, in, s_lv, l, res ? s_plan.rowbase[l] : s_lv[l].start,
                                                      res ? D : xs, res ? kSmemTag : 0);
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
                    if (off.x < 0) {                     // tag: the four corners are rows in shared memory
                        load4_shared<VT>(simg + (off.x & 0x7fffffff), v00);
                        load4_shared<VT>(simg + (off.y & 0x7fffffff), v01);
                        load4_shared<VT>(simg + (off.z & 0x7fffffff), v10);
                        load4_shared<VT>(simg + (off.w & 0x7fffffff), v11);
                    } else {
                        Vec4<VT>::template gather<1>(vimg + off.x, v00);
                        Vec4<VT>::template gather<1>(vimg + off.y, v01);
                        Vec4<VT>::template gather<1>(vimg + off.z, v10);
                        Vec4<VT>::template gather<1>(vimg + off.w, v11);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[c] += wa.x * v00[c] + wa.y * v01[c] + wa.z * v10[c] + wa.w * v11[c];
                }
                __syncwarp();
            }
        }
        if (qvalid) Vec4<VT>::store(out + qm * D + gl * kChannelsPerLane, acc);
    }
}

template <typename VT, int D, bool FUSED>
int run_res(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn, void *out,
            const Dims &d, const void *ref, cudaStream_t st)
{
    constexpr int G = D / kChannelsPerLane;
    constexpr int SMEM = kResBytes + (kResThreads / 32) * RecordLayout<G>::WARP_WORDS * 4;
    auto kern = fwd_res_kernel<VT, D, FUSED>;
    static bool prepared[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !prepared[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) prepared[dev] = true;
    }
    const long n_chunks = (d.Lq + kResQueries - 1) / kResQueries;
    const long grid = (long)d.N * d.M * n_chunks;
    if (grid > 0x7fffffffL) return kUnsupported;
    kern<<<(unsigned)grid, kResThreads, SMEM, st>>>((const VT *)value, shapes, lsi, (const float *)loc, (const float *)attn,
                                                     (VT *)out, d, (const float *)ref);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT, bool FUSED>
int dispatch_res(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                 void *out, const Dims &d, const void *ref, cudaStream_t st)
{
    switch (d.D) {
    case 16: return run_res<VT, 16, FUSED>(value, shapes, lsi, loc, attn, out, d, ref, st);
    case 32: return run_res<VT, 32, FUSED>(value, shapes, lsi, loc, attn, out, d, ref, st);
    case 64: return run_res<VT, 64, FUSED>(value, shapes, lsi, loc, attn, out, d, ref, st);
    }
    return kUnsupported;
}

}  // namespace

// fwd_variant: 30 forces this kernel for any Lq; 10 / 11 / 99 never use it.
bool resident_forward_applies(const Dims &d, DType dt, bool vec_ok)
{
    const int v = tuning().fwd_variant;
    if (!vec_ok || dt == DType::F64 || !(d.D == 16 || d.D == 32 || d.D == 64)) return false;
    if ((long)d.S * d.M * d.D >= (1L << 31) || d.L * d.P < 1 || d.S <= 0) return false;
    if (v == 30) return true;
    return v == -1 && tuning().fwd_pipe == -1 && d.Lq >= kResMinQueries && kResidentForwardDefault;
}

int launch_forward_resident(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                            const void *attn, void *out, const Dims &d, const void *ref, cudaStream_t st)
{
    if (ref) {
        if (d.L * d.P > kMaxBatches * (d.D / kChannelsPerLane)) return kUnsupported;
        if (dt == DType::F32) return dispatch_res<float, true>(value, shapes, lsi, loc, attn, out, d, ref, st);
        return dispatch_res<__nv_bfloat16, true>(value, shapes, lsi, loc, attn, out, d, ref, st);
    }
    if (dt == DType::F32) return dispatch_res<float, false>(value, shapes, lsi, loc, attn, out, d, nullptr, st);
    return dispatch_res<__nv_bfloat16, false>(value, shapes, lsi, loc, attn, out, d, nullptr, st);
}

}  // namespace msda
