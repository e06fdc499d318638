/* monodetr_step_b200.h -- C ABI of libmonodetr_step_b200.so: device-side pieces of the MonoDETR training
 * step that the reference runs on the host between kernels.
 *
 * Not part of the MSDA operator (include/msda_b200.h); a separate library so that the operator's ABI stays
 * exactly what the reference's extension exported.  Same conventions: plain pointers and sizes, caller owns
 * all memory, the stream is passed in, no host synchronisation, 0 on success / negative argument error /
 * positive cudaError_t, detr_step_last_error() for the message.
 */
#ifndef MONODETR_STEP_B200_H_
#define MONODETR_STEP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DETR_STEP_MAX_IMAGES 256 /* per call; image sizes travel as kernel arguments (no device copy, no sync) */

enum {
    DETR_STEP_OK = 0,
    DETR_STEP_ERR_NULL_POINTER = -1,
    DETR_STEP_ERR_BAD_SHAPE = -2,
    DETR_STEP_ERR_UNSUPPORTED = -4 /* a sub-problem does not fit the kernel's shared memory: use the host path */
};

/* Group-wise Hungarian matching on the device.
 *
 * Replaces the host section of HungarianMatcher.forward -- MonoDETR/lib/models/monodetr/matcher.py:87-104:
 * there the cost matrix is copied to the host (`C.cpu()`, a device synchronisation per decoder layer) and
 * scipy.optimize.linear_sum_assignment runs once per image and query group (16 x 11 calls per layer).
 *
 *   cost        DEVICE float [B, Q, T]: matching cost of query q of image b against target t, where T is the
 *               number of targets of the WHOLE batch (matcher.py:86); image b owns the columns
 *               [sum(sizes[:b]), sum(sizes[:b+1])).
 *   sizes       HOST int [B]: targets per image (known on the host: len(t["boxes"])); read during the call.
 *   groups      query groups (matcher.py:93): the Q queries are split into `groups` consecutive blocks of
 *               Q / groups, each matched against the image's targets independently.
 *   out_query / out_target
 *               DEVICE int64 [groups * sum_b min(sizes[b], Q / groups)]: per image one block of
 *               groups * k_b pairs (k_b = min(sizes[b], Q / groups)), group after group, inside a group ordered
 *               by query index -- i.e. exactly the concatenation matcher.py:99-103 builds from scipy's
 *               (row_ind, col_ind); out_query already carries the group offset g * (Q / groups).
 * The assignment is the exact minimum (shortest augmenting paths with fp64 potentials, like scipy's
 * rectangular solver); for cost matrices without exact ties it is THE optimum, hence identical to scipy's.
 * One warp per (image, group) sub-problem. */
int detr_group_lsa_f32(const float *cost, const int *sizes, int B, int Q, int T, int groups,
                       int64_t *out_query, int64_t *out_target, void *stream);

/* Same, plus a DEVICE int `status` (may be NULL): bit 0 is OR-ed in when a sub-problem met only non-finite
 * reduced costs (NaN / Inf in `cost`) and fell back to an arbitrary free column.  scipy raises ValueError
 * ("matrix contains invalid numeric entries") there, which stops a diverged run; the caller of this library
 * reads `status` when it next synchronises and raises the same error (monosowa_b200/step_host/lsa.py does, one
 * matcher call late, without adding a synchronisation to the step). */
int detr_group_lsa_status_f32(const float *cost, const int *sizes, int B, int Q, int T, int groups,
                              int64_t *out_query, int64_t *out_target, int *status, void *stream);

/* FrozenBatchNorm2d (+ residual add) (+ ReLU) in one pass over a contiguous NCHW fp32 activation of `n` elements
 * (`hw` = H*W, `C` channels): replaces the element-wise kernels behind MonoDETR/lib/models/monodetr/backbone.py:55-65
 * (`x * scale + bias`) and torchvision's Bottleneck tail (`out += identity; relu(out)`).
 *   y = relu?( fadd( fadd( fmul(x, scale[c]), bias[c] ), residual? ) )     every step rounded, the reference's order:
 * results are bit-identical to the separate kernels.  `residual` may be NULL, `relu` is 0 / 1.
 * Backward: grad_x = fmul(m, scale[c]), grad_residual (may be NULL) = m, with m = grad_y where y > 0 (relu) else grad_y;
 * `y` is only read when relu != 0.  scale / bias are frozen buffers: no gradient.  Return 0, a negative
 * DETR_STEP_ERR_* code, or a cudaError_t from the launch. */
int detr_frozen_bn_act_f32(const float *x, const float *residual, const float *scale, const float *bias, float *y,
                           long long n, long long hw, int C, int relu, void *stream);
int detr_frozen_bn_act_backward_f32(const float *grad_y, const float *y, const float *scale, float *grad_x,
                                    float *grad_residual, long long n, long long hw, int C, int relu, void *stream);

const char *detr_step_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MONODETR_STEP_B200_H_ */
