/*
 * msda_b200.h -- C ABI of libmsda_b200.so: MultiScaleDeformableAttention forward/backward
 * for NVIDIA B200 (sm_100a).  Plain pointers and sizes only: no torch / ATen / pybind types.
 *
 * This is the drop-in boundary for the reference's compiled extension
 * "MultiScaleDeformableAttention" (jskvrna/MonoSOWA, paths relative to
 * MonoDETR/lib/models/monodetr/ops/):
 *
 *   msda_forward_*   replaces  ms_deform_attn_forward   src/ms_deform_attn.h:20-38
 *                              -> ms_deform_attn_cuda_forward   src/cuda/ms_deform_attn_cuda.cu:20-80
 *                              -> ms_deformable_im2col_cuda     src/cuda/ms_deform_im2col_cuda.cuh:923-954
 *   msda_backward_*  replaces  ms_deform_attn_backward  src/ms_deform_attn.h:41-61
 *                              -> ms_deform_attn_cuda_backward  src/cuda/ms_deform_attn_cuda.cu:83-153
 *                              -> ms_deformable_col2im_cuda     src/cuda/ms_deform_im2col_cuda.cuh:956-1327
 *   (Python binding that sat on top: PYBIND11_MODULE in src/vision.cpp:13-16, called from
 *    functions/ms_deform_attn_func.py:25 and :35.)
 *
 * Differences from the reference boundary, all deliberate:
 *   - the caller owns every buffer; the library never allocates or frees device memory
 *     (the reference allocates outputs with at::zeros, ms_deform_attn_cuda.cu:54,121-123);
 *   - the whole batch is processed by one launch -- the reference's im2col_step chunk loop
 *     (ms_deform_attn_cuda.cu:61,131) is a host-side artefact with no numerical effect; the
 *     Python layer still validates batch % min(batch, im2col_step) == 0 like :52 does;
 *   - errors are returned, not printf'ed (the reference prints cudaGetLastError, cuh:948-952);
 *   - a bf16 variant exists (the reference dispatches float/double only, .cu:64,134).
 *
 * Tensor layouts (all contiguous, row-major, device memory unless noted):
 *   value            [N, S, M, D]          S = sum_l H_l*W_l
 *   spatial_shapes   [L, 2]  int64 (H, W)  DEVICE memory (read inside the kernels)
 *   level_start_index[L]     int64         DEVICE memory
 *   sampling_loc     [N, Lq, M, L, P, 2]   normalised (x, y); may lie outside [0, 1]
 *   attn_weight      [N, Lq, M, L, P]
 *   out / grad_out   [N, Lq, M, D]
 *   grad_value       like value; grad_loc like sampling_loc; grad_attn like attn_weight
 * Semantics: pixel coordinate = loc * size - 0.5 (grid_sample align_corners=False), bilinear,
 * zero padding per corner, samples outside (-1, size) skipped (cuh:285-291, 33-84).
 *
 * Alignment: every pointer must be aligned to its element size (else MSDA_ERR_MISALIGNED).  The fast kernels
 * (D in {16, 32, 64}) need 16-byte aligned tensor pointers, and the forward takes its widest loads (eight channels
 * per lane) when `value` is aligned to eight elements (32 bytes fp32, 16 bytes bf16); anything less aligned runs the
 * narrower or the generic kernels with identical results -- torch allocations are 512-byte aligned.
 *
 * Every entry point enqueues work on `stream` (a cudaStream_t passed as void*; NULL = the
 * legacy default stream) and returns immediately: no host synchronisation, no host reads of
 * device metadata, safe inside CUDA-graph capture.  Return value: MSDA_OK, a negative
 * MSDA_ERR_* argument error, or a positive cudaError_t from the launch.  After a non-zero
 * return msda_last_error() (thread-local) describes it.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 2
#define MSDA_MAX_LEVELS 16

enum {
    MSDA_OK = 0,
    MSDA_ERR_NULL_POINTER = -1, /* a required pointer is NULL (and the tensor is non-empty) */
    MSDA_ERR_BAD_SHAPE = -2,    /* negative size, L > MSDA_MAX_LEVELS, or a size overflow   */
    MSDA_ERR_MISALIGNED = -3,   /* a pointer is not aligned to its element size             */
    MSDA_ERR_UNSUPPORTED = -4,  /* fused entry points only: no fused kernel for this shape  */
};

/* ---- fp32: value/loc/attn/out all float ------------------------------------------------ */
int msda_forward_f32(const void *value, const int64_t *spatial_shapes,
                     const int64_t *level_start_index, const void *sampling_loc,
                     const void *attn_weight, void *out,
                     int N, int S, int M, int D, int L, int Lq, int P, void *stream);

/* grad_value is zero-filled by the library on `stream` before the scatter. */
int msda_backward_f32(const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, const void *grad_out,
                      void *grad_value, void *grad_loc, void *grad_attn,
                      int N, int S, int M, int D, int L, int Lq, int P, void *stream);

/* ---- fp64: everything double (keeps the reference's gradcheck contract, ops/test.py:63-86) */
int msda_forward_f64(const void *value, const int64_t *spatial_shapes,
                     const int64_t *level_start_index, const void *sampling_loc,
                     const void *attn_weight, void *out,
                     int N, int S, int M, int D, int L, int Lq, int P, void *stream);

int msda_backward_f64(const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, const void *grad_out,
                      void *grad_value, void *grad_loc, void *grad_attn,
                      int N, int S, int M, int D, int L, int Lq, int P, void *stream);

/* ---- bf16: value / out / grad_out / grad_value are bfloat16; sampling_loc, attn_weight, grad_loc
 *      and grad_attn stay float (bf16 cannot hold sub-pixel coordinates); arithmetic is fp32.
 *      The backward writes the bf16 gradient itself.  Short query sets scatter straight into it
 *      (REDG.E.ADD.BF16x4: a row receives a handful of additions); long query sets and odd shapes
 *      accumulate in fp32 and narrow once -- those need `scratch_f32`, a caller-owned DEVICE buffer of
 *      msda_backward_bf16_scratch_bytes(...) bytes (0 = not needed, pass NULL).
 *      `pointers_aligned16`: whether every tensor pointer of the call is 16-byte aligned. ---------- */
int msda_forward_bf16(const void *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const void *sampling_loc,
                      const void *attn_weight, void *out,
                      int N, int S, int M, int D, int L, int Lq, int P, void *stream);

int msda_backward_bf16(const void *value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const void *sampling_loc,
                       const void *attn_weight, const void *grad_out,
                       void *grad_value, void *grad_loc, void *grad_attn, void *scratch_f32,
                       int N, int S, int M, int D, int L, int Lq, int P, void *stream);
size_t msda_backward_bf16_scratch_bytes(int N, int S, int M, int D, int L, int Lq, int P,
                                        int pointers_aligned16);

/* ---- fused pre-processing -----------------------------------------
 * Same op, but the kernels consume the RAW outputs of the module's two Linears and build the
 * sampling locations and softmax weights in registers -- what the reference module does in PyTorch
 * at ops/modules/ms_deform_attn.py:145-155:
 *     attn = softmax(attn_logits over L*P)
 *     ref_dim 2:  loc = reference_points[l] + sampling_offsets / (W_l, H_l)                   (:149-152)
 *     ref_dim 6:  loc = ref[l][:2] + sampling_offsets / P * (ref[l][2]+ref[l][3], ref[l][4]+ref[l][5]) * 0.5
 *                                                                                             (:153-155)
 *   reference_points [N, Lq, L, ref_dim] float (not differentiated here: d loc / d ref[:2] = 1, so the
 *   Python layer sums grad_offsets when a caller needs it), sampling_offsets [N, Lq, M, L, P, 2] float,
 *   attn_logits [N, Lq, M, L*P] float; value / out / grad_out / grad_value float or bf16 as above.
 *   backward writes grad_value (zero-filled here), grad_offsets, grad_logits; the bf16 flavour takes
 *   the same `scratch_f32` as msda_backward_bf16.
 * Supported: D in {16, 32, 64}, L*P <= D, 16-byte aligned pointers; otherwise MSDA_ERR_UNSUPPORTED
 * is returned and nothing is launched (callers fall back to the unfused entry points). */
int msda_forward_fused_f32(const void *value, const int64_t *spatial_shapes,
                           const int64_t *level_start_index, const void *reference_points, int ref_dim,
                           const void *sampling_offsets, const void *attn_logits, void *out,
                           int N, int S, int M, int D, int L, int Lq, int P, void *stream);
int msda_forward_fused_bf16(const void *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const void *reference_points, int ref_dim,
                            const void *sampling_offsets, const void *attn_logits, void *out,
                            int N, int S, int M, int D, int L, int Lq, int P, void *stream);
int msda_backward_fused_f32(const void *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const void *reference_points, int ref_dim,
                            const void *sampling_offsets, const void *attn_logits,
                            const void *grad_out, void *grad_value, void *grad_offsets,
                            void *grad_logits, int N, int S, int M, int D, int L, int Lq, int P,
                            void *stream);
int msda_backward_fused_bf16(const void *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const void *reference_points, int ref_dim,
                             const void *sampling_offsets, const void *attn_logits,
                             const void *grad_out, void *grad_value, void *grad_offsets,
                             void *grad_logits, void *scratch_f32, int N, int S, int M, int D, int L, int Lq,
                             int P, void *stream);

/* ---- host-buffer step --------------------------------------------------------------------------
 * One forward + backward of a whole batch whose tensors live in HOST memory (pinned memory for
 * asynchronous copies): value / sampling_loc / attn_weight / grad_out in, out / grad_value /
 * grad_loc / grad_attn back, all with the layouts above (grad_value is bf16 for the bf16 flavour, as
 * in msda_backward_bf16).  spatial_shapes and level_start_index stay DEVICE pointers.  The reference
 * has no counterpart: its extension takes CUDA tensors only (ops/src/ms_deform_attn.h:29-38), so a
 * caller with host data pays cudaMemcpy + kernels + cudaMemcpy serially.  Here the batch is pipelined
 * in chunks of `images_per_chunk` images (images are independent, cuh:269) over two internal copy
 * streams and `stream`: H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1 overlap.
 *   workspace : DEVICE scratch of at least msda_host_step_workspace_bytes(...) bytes (three pipeline
 *               stages), 256-byte aligned, owned by the caller (the library never allocates device
 *               memory).  A larger workspace is used for a deeper ring, up to 16 stages (whole
 *               multiples of a third of the minimum): worth it when H2D and D2H run at different paces.
 * Nothing blocks the host; results are valid once `stream` has been synchronised.  The copy streams
 * are per device: calls on one device are serialised by that device's mutex and each call waits for the
 * previous call's last copy, so calls from different caller streams need SEPARATE workspaces only if they
 * are meant to overlap -- they never corrupt each other. */
int msda_host_step_f32(const void *h_value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const void *h_sampling_loc,
                       const void *h_attn_weight, const void *h_grad_out, void *h_out,
                       void *h_grad_value, void *h_grad_loc, void *h_grad_attn, void *workspace,
                       size_t workspace_bytes, int N, int S, int M, int D, int L, int Lq, int P,
                       int images_per_chunk, void *stream);
int msda_host_step_bf16(const void *h_value, const int64_t *spatial_shapes,
                        const int64_t *level_start_index, const void *h_sampling_loc,
                        const void *h_attn_weight, const void *h_grad_out, void *h_out,
                        void *h_grad_value, void *h_grad_loc, void *h_grad_attn, void *workspace,
                        size_t workspace_bytes, int N, int S, int M, int D, int L, int Lq, int P,
                        int images_per_chunk, void *stream);
size_t msda_host_step_workspace_bytes(int is_bf16, int S, int M, int D, int L, int Lq, int P,
                                      int images_per_chunk);

/* ---- introspection ---------------------------------------------------------------------- */
int msda_abi_version(void);            /* == MSDA_ABI_VERSION                                */
const char *msda_build_info(void);     /* "sm_100a nvcc <ver> <date>"                        */
const char *msda_last_error(void);     /* thread-local; "" when the last call succeeded      */
long long msda_launch_count(void);     /* kernels this library has launched so far (memsets excluded) */

/* Kernel-selection knobs, for benchmarking A/B runs and tests only (process-wide, not thread-safe
 * against concurrent launches).  Keys: "fwd_variant" (11 = record kernel, 99 = generic kernel;
 * 12 = tile kernel, -DMSDA_AB builds only), "bwd_variant" (11 = record kernel for any Lq, 21 = binned
 * kernel for any Lq, 99 = generic kernel; 20 = tile kernel, -DMSDA_AB builds only), "fwd_pipe" /
 * "bwd_pipe" (4 = four channels per lane in the record kernels, where the default would take eight; other values:
 * launch flavours of the -DMSDA_AB kernels, ignored by the shipped library).  value -1 restores the default.
 * Returns MSDA_OK or MSDA_ERR_BAD_SHAPE (unknown key). */
int msda_set_tuning(const char *key, int value);
int msda_get_tuning(const char *key);

/* Name of the kernel family the current heuristics pick for this problem: "fwd_rec_f32",
 * "bwd_bin_bf16", "fwd_generic_f64", ... (static storage; for logs, tests and bench.py).  Long query
 * sets (Lq >= 1024) with enough (image, head, chunk) work items take the binned backward, which combines
 * the coarse levels' grad_value contributions in shared memory. */
const char *msda_describe_forward(int dtype_bits, int is_bf16, int N, int M, int D, int L, int P, int Lq);
const char *msda_describe_backward(int dtype_bits, int is_bf16, int N, int M, int D, int L, int P, int Lq);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
