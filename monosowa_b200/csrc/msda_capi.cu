// msda_capi.cu -- the extern "C" surface declared in include/msda_b200.h.
// Argument validation mirrors what the reference's host wrapper asserts
// (MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_attn_cuda.cu:28-52, 93-119) as far as a
// raw-pointer interface can (contiguity / device placement are the Python layer's job).
#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "msda_common.cuh"

#define MSDA_STR2(x) #x
#define MSDA_STR(x) MSDA_STR2(x)

namespace msda {

static Tuning g_tuning;
Tuning &tuning() { return g_tuning; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char t_err[256] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_result(int e, const char *what)
{
    if (e == 0) {
        t_err[0] = 0;
        return 0;
    }
    snprintf(t_err, sizeof(t_err), "%s: %s (cudaError %d)", what, cudaGetErrorString((cudaError_t)e), e);
    return e;
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

struct Checked {
    Dims d;
    bool vec_ok;
    bool empty_out;
};

static int check_common(const char *fn, DType dt, const void *value, const int64_t *shapes,
                        const int64_t *lsi, const void *loc, const void *attn, int N, int S, int M,
                        int D, int L, int Lq, int P, Checked *out)
{
    if (N < 0 || S < 0 || M < 0 || D < 0 || L < 0 || Lq < 0 || P < 0)
        return fail(MSDA_ERR_BAD_SHAPE, "%s: negative dimension (N=%d S=%d M=%d D=%d L=%d Lq=%d P=%d)", fn, N, S, M,
                    D, L, Lq, P);
    if (L > MSDA_MAX_LEVELS)
        return fail(MSDA_ERR_BAD_SHAPE, "%s: L=%d exceeds MSDA_MAX_LEVELS=%d", fn, L, MSDA_MAX_LEVELS);
    const long long big = 1LL << 62;
    const long long nv = (long long)N * S * M * D, nl = (long long)N * Lq * M * L * P * 2;
    if (nv < 0 || nv > big || nl < 0 || nl > big)
        return fail(MSDA_ERR_BAD_SHAPE, "%s: tensor size overflows", fn);
    const bool has_samples = (long long)N * Lq * M * L * P > 0;
    if (L > 0 && (!shapes || !lsi))
        return fail(MSDA_ERR_NULL_POINTER, "%s: spatial_shapes / level_start_index is NULL", fn);
    if (nv > 0 && !value) return fail(MSDA_ERR_NULL_POINTER, "%s: value is NULL", fn);
    if (has_samples && (!loc || !attn)) return fail(MSDA_ERR_NULL_POINTER, "%s: sampling_loc / attn_weight is NULL", fn);
    const size_t ve = dt == DType::F64 ? 8 : (dt == DType::F32 ? 4 : 2);
    const size_t ce = dt == DType::F64 ? 8 : 4;
    if (!aligned(value, ve) || !aligned(loc, ce) || !aligned(attn, ce) || !aligned(shapes, 8) || !aligned(lsi, 8))
        return fail(MSDA_ERR_MISALIGNED, "%s: a pointer is not aligned to its element size", fn);
    out->d = Dims{N, S, M, D, L, Lq, P};
    out->vec_ok = aligned(value, 16) && aligned(loc, 16) && aligned(attn, 16);
    out->empty_out = (long long)N * Lq * M * D == 0;
    return 0;
}

static int forward_impl(const char *fn, DType dt, const void *value, const int64_t *shapes,
                        const int64_t *lsi, const void *loc, const void *attn, void *out, int N, int S,
                        int M, int D, int L, int Lq, int P, void *stream)
{
    Checked c;
    if (int rc = check_common(fn, dt, value, shapes, lsi, loc, attn, N, S, M, D, L, Lq, P, &c)) return rc;
    if (c.empty_out) return cuda_result(0, fn);
    if (!out) return fail(MSDA_ERR_NULL_POINTER, "%s: out is NULL", fn);
    c.vec_ok = c.vec_ok && aligned(out, 16);
    return cuda_result(launch_forward(dt, value, shapes, lsi, loc, attn, out, c.d, c.vec_ok, (cudaStream_t)stream), fn);
}

static int backward_impl(const char *fn, DType dt, const void *value, const int64_t *shapes,
                         const int64_t *lsi, const void *loc, const void *attn, const void *grad_out,
                         void *gv, void *gl, void *ga, void *scratch, int N, int S, int M, int D, int L, int Lq, int P,
                         void *stream)
{
    Checked c;
    if (int rc = check_common(fn, dt, value, shapes, lsi, loc, attn, N, S, M, D, L, Lq, P, &c)) return rc;
    const bool has_value = (long long)N * S * M * D > 0;
    const bool has_samples = (long long)N * Lq * M * L * P > 0;
    if (has_value && !gv) return fail(MSDA_ERR_NULL_POINTER, "%s: grad_value is NULL", fn);
    if (has_samples && (!gl || !ga)) return fail(MSDA_ERR_NULL_POINTER, "%s: grad_loc / grad_attn is NULL", fn);
    if (!c.empty_out && !grad_out) return fail(MSDA_ERR_NULL_POINTER, "%s: grad_out is NULL", fn);
    c.vec_ok = c.vec_ok && aligned(grad_out, 16) && aligned(gv, 16) && aligned(gl, 16) && aligned(ga, 16) &&
               aligned(scratch, 16);
    const int rc = launch_backward(dt, value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, scratch, c.d, c.vec_ok,
                                   (cudaStream_t)stream);
    if (rc == kNeedsScratch)
        return fail(MSDA_ERR_NULL_POINTER, "%s: scratch_f32 is NULL but this shape accumulates in fp32 "
                    "(msda_backward_bf16_scratch_bytes)", fn);
    return cuda_result(rc, fn);
}

static int fused_impl(const char *fn, DType dt, bool backward, const void *value, const int64_t *shapes,
                      const int64_t *lsi, const void *ref, int ref_dim, const void *offsets, const void *logits,
                      const void *grad_out, void *o0, void *o1, void *o2, void *scratch, int N, int S, int M, int D,
                      int L, int Lq, int P, void *stream)
{
    Checked c;
    if (int rc = check_common(fn, dt, value, shapes, lsi, offsets, logits, N, S, M, D, L, Lq, P, &c)) return rc;
    if (ref_dim != 2 && ref_dim != 6)
        return fail(MSDA_ERR_UNSUPPORTED, "%s: reference points must have 2 or 6 components, got %d", fn, ref_dim);
    const bool has_q = (long long)N * Lq * M > 0;
    if (has_q && L > 0 && !ref) return fail(MSDA_ERR_NULL_POINTER, "%s: reference_points is NULL", fn);
    if (!backward && !c.empty_out && !o0) return fail(MSDA_ERR_NULL_POINTER, "%s: out is NULL", fn);
    if (backward && ((!c.empty_out && !grad_out) || ((long long)N * S * M * D > 0 && !o0) ||
                     ((long long)N * Lq * M * L * P > 0 && (!o1 || !o2))))
        return fail(MSDA_ERR_NULL_POINTER, "%s: a gradient pointer is NULL", fn);
    const bool al = c.vec_ok && aligned(ref, 8) && aligned(o0, 16) && aligned(o1, 8) && aligned(o2, 4) &&
                    aligned(grad_out, 8) && aligned(scratch, 16);
    if (!al) return fail(MSDA_ERR_UNSUPPORTED, "%s: pointers are not 16-byte aligned", fn);
    int rc;
    if (!backward) {
        if (c.empty_out) return cuda_result(0, fn);
        rc = launch_forward_fused(dt, value, shapes, lsi, ref, ref_dim, offsets, logits, o0, c.d, (cudaStream_t)stream);
    } else {
        rc = launch_backward_fused(dt, value, shapes, lsi, ref, ref_dim, offsets, logits, grad_out, o0, o1, o2, scratch, c.d,
                                   (cudaStream_t)stream);
    }
    if (rc == kUnsupported)
        return fail(MSDA_ERR_UNSUPPORTED, "%s: no fused kernel for D=%d L*P=%d (use the unfused entry points)", fn, D, L * P);
    if (rc == kNeedsScratch)
        return fail(MSDA_ERR_NULL_POINTER, "%s: scratch_f32 is NULL but this shape accumulates in fp32 "
                    "(msda_backward_bf16_scratch_bytes)", fn);
    return cuda_result(rc, fn);
}

// ---- host-buffer step ------------------------------------------------------------------------------
// forward + backward of a batch whose tensors live in HOST memory (include/msda_b200.h, "host-buffer
// step").  Images are independent (the value index is prefixed by n, reference cuh:269), so the batch is
// pipelined in chunks of whole images over two internal copy streams and the caller's stream:
//   s_in : H2D of chunk i+1     |  stream : forward + backward of chunk i  |  s_out : D2H of chunk i-1
// through a ring of device stages carved from the caller's workspace: kHostStages (3) at least, as many more as the
// workspace holds (up to kHostStagesMax) -- a deeper ring lets the H2D side run ahead when the two directions do not
// proceed at the same pace (8 GPUs behind one host fabric: profiles/r02_n8_e2e_stages.txt).  Nothing here blocks
// the host; the caller synchronises `stream` before reading the results.
constexpr int kHostStages = 3;
constexpr int kHostStagesMax = 16;

struct HostPipe {
    std::mutex mu;                  // serialises the calls of one device (they share the copy streams)
    bool ready = false, has_prev = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t start = nullptr, done = nullptr, in[kHostStagesMax] = {}, cmp[kHostStagesMax] = {}, out[kHostStagesMax] = {};
};
static HostPipe g_pipe[64];

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

struct HostStage {
    size_t value, loc, attn, grad_out, out, gv, gl, ga, scratch, total;
};

static HostStage host_stage_layout(DType dt, const Dims &d, int images)
{
    const size_t ve = dt == DType::F64 ? 8 : (dt == DType::F32 ? 4 : 2), ce = dt == DType::F64 ? 8 : 4;
    const size_t n = (size_t)images;
    const size_t vb = n * d.S * d.M * d.D * ve, lb = n * d.Lq * d.M * d.L * d.P * 2 * ce, ab = lb / 2;
    const size_t ob = n * d.Lq * d.M * d.D * ve;
    Dims dc = d;
    dc.N = images;
    const size_t sb = backward_needs_scratch(dc, dt, true) ? n * d.S * d.M * d.D * 4 : 0;
    HostStage st;
    size_t o = 0;
    st.value = o; o += align256(vb);
    st.loc = o; o += align256(lb);
    st.attn = o; o += align256(ab);
    st.grad_out = o; o += align256(ob);
    st.out = o; o += align256(ob);
    st.gv = o; o += align256(vb);
    st.gl = o; o += align256(lb);
    st.ga = o; o += align256(ab);
    st.scratch = o; o += align256(sb);
    st.total = o;
    return st;
}

static void host_pipe_destroy(HostPipe &p)
{
    if (p.s_in) cudaStreamDestroy(p.s_in);
    if (p.s_out) cudaStreamDestroy(p.s_out);
    if (p.start) cudaEventDestroy(p.start);
    if (p.done) cudaEventDestroy(p.done);
    for (int i = 0; i < kHostStagesMax; ++i) {
        if (p.in[i]) cudaEventDestroy(p.in[i]);
        if (p.cmp[i]) cudaEventDestroy(p.cmp[i]);
        if (p.out[i]) cudaEventDestroy(p.out[i]);
        p.in[i] = p.cmp[i] = p.out[i] = nullptr;
    }
    p.s_in = p.s_out = nullptr;
    p.start = p.done = nullptr;
    p.ready = p.has_prev = false;
}

// called with p.mu held
static int host_pipe_prepare(HostPipe &p)
{
    if (p.ready) return 0;
    cudaError_t e = cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming);
    for (int i = 0; i < kHostStagesMax && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p.in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.cmp[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        host_pipe_destroy(p);               // nothing half-built survives a failed attempt
        return (int)e;
    }
    p.ready = true;
    return 0;
}

#define MSDA_CU(call)                                           \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) return cuda_result((int)e_, fn); \
    } while (0)

static int host_step_impl(const char *fn, DType dt, const void *h_value, const int64_t *shapes, const int64_t *lsi,
                          const void *h_loc, const void *h_attn, const void *h_grad_out, void *h_out, void *h_gv,
                          void *h_gl, void *h_ga, void *workspace, size_t workspace_bytes, int N, int S, int M, int D,
                          int L, int Lq, int P, int images_per_chunk, void *stream)
{
    Checked c;
    if (int rc = check_common(fn, dt, h_value, shapes, lsi, h_loc, h_attn, N, S, M, D, L, Lq, P, &c)) return rc;
    if (images_per_chunk < 1) return fail(MSDA_ERR_BAD_SHAPE, "%s: images_per_chunk must be >= 1", fn);
    if (N == 0) return cuda_result(0, fn);
    if (!h_grad_out || !h_out || !h_gv || !h_gl || !h_ga || !workspace)
        return fail(MSDA_ERR_NULL_POINTER, "%s: a host result pointer, grad_out or the workspace is NULL", fn);
    const int cb = images_per_chunk < N ? images_per_chunk : N;
    const HostStage lay = host_stage_layout(dt, c.d, cb);
    if (workspace_bytes < (size_t)kHostStages * lay.total || !aligned(workspace, 256))
        return fail(MSDA_ERR_BAD_SHAPE, "%s: workspace too small or not 256-byte aligned (%zu bytes needed)", fn,
                    (size_t)kHostStages * lay.total);
    const size_t fit = lay.total ? workspace_bytes / lay.total : (size_t)kHostStagesMax;
    const int stages = (int)(fit < (size_t)kHostStagesMax ? fit : (size_t)kHostStagesMax);
    int dev = 0;
    MSDA_CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return cuda_result((int)cudaErrorInvalidDevice, fn);
    HostPipe &p = g_pipe[dev];
    std::lock_guard<std::mutex> lock(p.mu);
    if (int rc = host_pipe_prepare(p)) return cuda_result(rc, fn);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ve = dt == DType::F64 ? 8 : (dt == DType::F32 ? 4 : 2), ce = dt == DType::F64 ? 8 : 4;
    const size_t vb = (size_t)S * M * D * ve, lb = (size_t)Lq * M * L * P * 2 * ce, ab = lb / 2;
    const size_t ob = (size_t)Lq * M * D * ve;

    MSDA_CU(cudaEventRecord(p.start, st));
    MSDA_CU(cudaStreamWaitEvent(p.s_in, p.start, 0));
    MSDA_CU(cudaStreamWaitEvent(p.s_out, p.start, 0));
    // the previous call on this device (possibly from another caller stream) may still be computing in or copying
    // out of the stages this call is about to overwrite: its last D2H is behind all of that
    if (p.has_prev) MSDA_CU(cudaStreamWaitEvent(p.s_in, p.done, 0));
    int chunk = 0;
    for (int n0 = 0; n0 < N; n0 += cb, ++chunk) {
        const int nb = (N - n0 < cb) ? N - n0 : cb;
        const int s = chunk % stages;
        char *base = (char *)workspace + (size_t)s * lay.total;
        if (chunk >= stages) MSDA_CU(cudaStreamWaitEvent(p.s_in, p.out[s], 0));               // stage drained
        MSDA_CU(cudaMemcpyAsync(base + lay.value, (const char *)h_value + n0 * vb, nb * vb, cudaMemcpyHostToDevice, p.s_in));
        MSDA_CU(cudaMemcpyAsync(base + lay.loc, (const char *)h_loc + n0 * lb, nb * lb, cudaMemcpyHostToDevice, p.s_in));
        MSDA_CU(cudaMemcpyAsync(base + lay.attn, (const char *)h_attn + n0 * ab, nb * ab, cudaMemcpyHostToDevice, p.s_in));
        MSDA_CU(cudaMemcpyAsync(base + lay.grad_out, (const char *)h_grad_out + n0 * ob, nb * ob, cudaMemcpyHostToDevice, p.s_in));
        MSDA_CU(cudaEventRecord(p.in[s], p.s_in));
        MSDA_CU(cudaStreamWaitEvent(st, p.in[s], 0));
        Dims dc = c.d;
        dc.N = nb;
        int rc = c.empty_out ? 0 : launch_forward(dt, base + lay.value, shapes, lsi, base + lay.loc, base + lay.attn,
                                                  base + lay.out, dc, true, st);
        if (rc) return cuda_result(rc, fn);
        rc = launch_backward(dt, base + lay.value, shapes, lsi, base + lay.loc, base + lay.attn, base + lay.grad_out,
                             base + lay.gv, base + lay.gl, base + lay.ga, base + lay.scratch, dc, true, st);
        if (rc) return cuda_result(rc, fn);
        MSDA_CU(cudaEventRecord(p.cmp[s], st));
        MSDA_CU(cudaStreamWaitEvent(p.s_out, p.cmp[s], 0));
        MSDA_CU(cudaMemcpyAsync((char *)h_out + n0 * ob, base + lay.out, nb * ob, cudaMemcpyDeviceToHost, p.s_out));
        MSDA_CU(cudaMemcpyAsync((char *)h_gv + n0 * vb, base + lay.gv, nb * vb, cudaMemcpyDeviceToHost, p.s_out));
        MSDA_CU(cudaMemcpyAsync((char *)h_gl + n0 * lb, base + lay.gl, nb * lb, cudaMemcpyDeviceToHost, p.s_out));
        MSDA_CU(cudaMemcpyAsync((char *)h_ga + n0 * ab, base + lay.ga, nb * ab, cudaMemcpyDeviceToHost, p.s_out));
        MSDA_CU(cudaEventRecord(p.out[s], p.s_out));
    }
    MSDA_CU(cudaEventRecord(p.done, p.s_out));
    p.has_prev = true;
    MSDA_CU(cudaStreamWaitEvent(st, p.done, 0));
    return cuda_result(0, fn);
}

}  // namespace msda

using namespace msda;

#define FWD_ARGS                                                                                     \
    const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,              \
        const void *sampling_loc, const void *attn_weight, void *out, int N, int S, int M, int D,    \
        int L, int Lq, int P, void *stream
#define FWD_PASS value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out, N, S, M, D, L, Lq, P, stream
#define BWD_ARGS                                                                                     \
    const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,              \
        const void *sampling_loc, const void *attn_weight, const void *grad_out, void *grad_value,   \
        void *grad_loc, void *grad_attn, int N, int S, int M, int D, int L, int Lq, int P, void *stream
#define BWD_PTRS value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_out, grad_value, grad_loc, grad_attn

extern "C" {

int msda_forward_f32(FWD_ARGS) { return forward_impl("msda_forward_f32", DType::F32, FWD_PASS); }
int msda_forward_f64(FWD_ARGS) { return forward_impl("msda_forward_f64", DType::F64, FWD_PASS); }
int msda_forward_bf16(FWD_ARGS) { return forward_impl("msda_forward_bf16", DType::BF16, FWD_PASS); }
int msda_backward_f32(BWD_ARGS)
{
    return backward_impl("msda_backward_f32", DType::F32, BWD_PTRS, nullptr, N, S, M, D, L, Lq, P, stream);
}
int msda_backward_f64(BWD_ARGS)
{
    return backward_impl("msda_backward_f64", DType::F64, BWD_PTRS, nullptr, N, S, M, D, L, Lq, P, stream);
}
int msda_backward_bf16(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, const void *grad_out, void *grad_value,
                       void *grad_loc, void *grad_attn, void *scratch_f32, int N, int S, int M, int D, int L, int Lq,
                       int P, void *stream)
{
    return backward_impl("msda_backward_bf16", DType::BF16, BWD_PTRS, scratch_f32, N, S, M, D, L, Lq, P, stream);
}

size_t msda_backward_bf16_scratch_bytes(int N, int S, int M, int D, int L, int Lq, int P, int pointers_aligned16)
{
    if (N < 0 || S < 0 || M < 0 || D < 0 || L < 0 || Lq < 0 || P < 0) return 0;
    const Dims d{N, S, M, D, L, Lq, P};
    return backward_needs_scratch(d, DType::BF16, pointers_aligned16 != 0) ? (size_t)N * S * M * D * 4 : 0;
}

#define FUSED_FWD_ARGS                                                                               \
    const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,              \
        const void *reference_points, int ref_dim, const void *sampling_offsets, const void *attn_logits, void *out, \
        int N, int S, int M, int D, int L, int Lq, int P, void *stream
#define FUSED_BWD_PTRS                                                                               \
    const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,              \
        const void *reference_points, int ref_dim, const void *sampling_offsets, const void *attn_logits, \
        const void *grad_out, void *grad_value, void *grad_offsets, void *grad_logits

int msda_forward_fused_f32(FUSED_FWD_ARGS)
{
    return fused_impl("msda_forward_fused_f32", DType::F32, false, value, spatial_shapes, level_start_index, reference_points,
                      ref_dim, sampling_offsets, attn_logits, nullptr, out, nullptr, nullptr, nullptr, N, S, M, D, L, Lq, P, stream);
}
int msda_forward_fused_bf16(FUSED_FWD_ARGS)
{
    return fused_impl("msda_forward_fused_bf16", DType::BF16, false, value, spatial_shapes, level_start_index, reference_points,
                      ref_dim, sampling_offsets, attn_logits, nullptr, out, nullptr, nullptr, nullptr, N, S, M, D, L, Lq, P, stream);
}
int msda_backward_fused_f32(FUSED_BWD_PTRS, int N, int S, int M, int D, int L, int Lq, int P, void *stream)
{
    return fused_impl("msda_backward_fused_f32", DType::F32, true, value, spatial_shapes, level_start_index, reference_points,
                      ref_dim, sampling_offsets, attn_logits, grad_out, grad_value, grad_offsets, grad_logits, nullptr, N, S, M,
                      D, L, Lq, P, stream);
}
int msda_backward_fused_bf16(FUSED_BWD_PTRS, void *scratch_f32, int N, int S, int M, int D, int L, int Lq, int P,
                             void *stream)
{
    return fused_impl("msda_backward_fused_bf16", DType::BF16, true, value, spatial_shapes, level_start_index, reference_points,
                      ref_dim, sampling_offsets, attn_logits, grad_out, grad_value, grad_offsets, grad_logits, scratch_f32, N, S,
                      M, D, L, Lq, P, stream);
}

#define HOST_ARGS                                                                                              \
    const void *h_value, const int64_t *spatial_shapes, const int64_t *level_start_index, const void *h_sampling_loc, \
        const void *h_attn_weight, const void *h_grad_out, void *h_out, void *h_grad_value, void *h_grad_loc,  \
        void *h_grad_attn, void *workspace, size_t workspace_bytes, int N, int S, int M, int D, int L, int Lq, \
        int P, int images_per_chunk, void *stream
#define HOST_PASS                                                                                              \
    h_value, spatial_shapes, level_start_index, h_sampling_loc, h_attn_weight, h_grad_out, h_out, h_grad_value, \
        h_grad_loc, h_grad_attn, workspace, workspace_bytes, N, S, M, D, L, Lq, P, images_per_chunk, stream

int msda_host_step_f32(HOST_ARGS) { return host_step_impl("msda_host_step_f32", DType::F32, HOST_PASS); }
int msda_host_step_bf16(HOST_ARGS) { return host_step_impl("msda_host_step_bf16", DType::BF16, HOST_PASS); }

size_t msda_host_step_workspace_bytes(int is_bf16, int S, int M, int D, int L, int Lq, int P, int images_per_chunk)
{
    if (S < 0 || M < 0 || D < 0 || L < 0 || Lq < 0 || P < 0 || images_per_chunk < 1) return 0;
    const Dims d{images_per_chunk, S, M, D, L, Lq, P};
    return (size_t)kHostStages * host_stage_layout(is_bf16 ? DType::BF16 : DType::F32, d, images_per_chunk).total;
}

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char *msda_build_info(void)
{
#ifdef MSDA_AB
    return "libmsda_b200 sm_100a nvcc " MSDA_STR(__CUDACC_VER_MAJOR__) "." MSDA_STR(__CUDACC_VER_MINOR__) " built " __DATE__ " +AB flavours";
#else
    return "libmsda_b200 sm_100a nvcc " MSDA_STR(__CUDACC_VER_MAJOR__) "." MSDA_STR(__CUDACC_VER_MINOR__) " built " __DATE__;
#endif
}

const char *msda_last_error(void) { return t_err; }

long long msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static int *tuning_slot(const char *key)
{
    if (!key) return nullptr;
    if (!strcmp(key, "fwd_variant")) return &tuning().fwd_variant;
    if (!strcmp(key, "bwd_variant")) return &tuning().bwd_variant;
    if (!strcmp(key, "fwd_pipe")) return &tuning().fwd_pipe;
    if (!strcmp(key, "bwd_pipe")) return &tuning().bwd_pipe;
    if (!strcmp(key, "bf16_direct")) return &tuning().bf16_direct;
    return nullptr;
}

int msda_set_tuning(const char *key, int value)
{
    int *slot = tuning_slot(key);
    if (!slot) return fail(MSDA_ERR_BAD_SHAPE, "msda_set_tuning: unknown key '%s'", key ? key : "(null)");
    *slot = value;
    return 0;
}

int msda_get_tuning(const char *key)
{
    int *slot = tuning_slot(key);
    return slot ? *slot : -1;
}

static DType dtype_of(int bits, int is_bf16) { return is_bf16 ? DType::BF16 : (bits == 64 ? DType::F64 : DType::F32); }

const char *msda_describe_forward(int dtype_bits, int is_bf16, int N, int M, int D, int L, int P, int Lq)
{
    return forward_kernel_name(dtype_of(dtype_bits, is_bf16), Dims{N, 1, M, D, L, Lq, P}, true);
}

const char *msda_describe_backward(int dtype_bits, int is_bf16, int N, int M, int D, int L, int P, int Lq)
{
    return backward_kernel_name(dtype_of(dtype_bits, is_bf16), Dims{N, 1, M, D, L, Lq, P}, true);
}

}  // extern "C"
