// step_bn.cu -- FrozenBatchNorm2d (+ residual add) (+ ReLU) of the MonoDETR backbone as ONE pass (libmonodetr_step_b200.so).
//
// The reference's FrozenBatchNorm2d.forward (MonoDETR/lib/models/monodetr/backbone.py:55-65) is `x * scale + bias` on
// broadcast (1,C,1,1) tensors: PyTorch runs it as two element-wise kernels (mul, add), torchvision's Bottleneck then adds
// the identity and applies ReLU in two more (resnet.py Bottleneck.forward), and autograd mirrors them: at the KITTI input
// (16 x 3 x 384 x 1280) an activation of layer1 is 503 MB, so every pass costs ~0.15 ms and the 53 normalisations of
// ResNet-50 add up to ~18 ms of element-wise kernels per training step (profiles/r02_training_step.md).
// Here: y = relu?( fadd(fadd?(fmul(x, scale[c]), bias[c]), residual) ) with EVERY operation rounded separately, in the
// reference's order (mul, add bias, add identity, clamp) -- bit-identical results, one read of x (+ residual), one
// write of y.  Backward: grad_x = fmul(mask(grad_y), scale[c]), grad_residual = mask(grad_y), mask = (y > 0) as
// threshold_backward does; scale / bias are frozen buffers and receive no gradient.
// Layout: contiguous NCHW fp32, channel of flat element i = (i / HW) % C.  The Python layer
// (monosowa_b200/step_host/frozen_bn.py) falls back to the reference's own forward for anything else.
#include <cuda_runtime.h>
#include <stdint.h>

#include "monodetr_step_b200.h"

namespace {

__device__ __forceinline__ float bn1(float x, float s, float b) { return __fadd_rn(__fmul_rn(x, s), b); }
__device__ __forceinline__ float relu1(float v) { return v < 0.f ? 0.f : v; }       // NaN stays NaN, like clamp_min_(0)

template <bool RELU, bool RES, int V>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float *__restrict__ x, const float *__restrict__ res, const float *__restrict__ scale,
                  const float *__restrict__ bias, float *__restrict__ y, const long n, const long hw, const int C)
{
    const long stride = (long)gridDim.x * blockDim.x * V;
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        const int c = (int)((i / hw) % C);
        const float s = __ldg(scale + c), b = __ldg(bias + c);
        if constexpr (V == 4) {
            const float4 v = *reinterpret_cast<const float4 *>(x + i);
            float4 o = make_float4(bn1(v.x, s, b), bn1(v.y, s, b), bn1(v.z, s, b), bn1(v.w, s, b));
            if constexpr (RES) {
                const float4 r = *reinterpret_cast<const float4 *>(res + i);
                o = make_float4(__fadd_rn(o.x, r.x), __fadd_rn(o.y, r.y), __fadd_rn(o.z, r.z), __fadd_rn(o.w, r.w));
            }
            if constexpr (RELU) o = make_float4(relu1(o.x), relu1(o.y), relu1(o.z), relu1(o.w));
            *reinterpret_cast<float4 *>(y + i) = o;
        } else {
            float o = bn1(x[i], s, b);
            if constexpr (RES) o = __fadd_rn(o, res[i]);
            if constexpr (RELU) o = relu1(o);
            y[i] = o;
        }
    }
}

template <bool RELU, bool RES, int V>
__global__ void __launch_bounds__(256)
bn_act_bwd_kernel(const float *__restrict__ gy, const float *__restrict__ y, const float *__restrict__ scale,
                  float *__restrict__ gx, float *__restrict__ gres, const long n, const long hw, const int C)
{
    const long stride = (long)gridDim.x * blockDim.x * V;
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        const int c = (int)((i / hw) % C);
        const float s = __ldg(scale + c);
        if constexpr (V == 4) {
            float4 g = *reinterpret_cast<const float4 *>(gy + i);
            if constexpr (RELU) {
                const float4 o = *reinterpret_cast<const float4 *>(y + i);
                g = make_float4(o.x > 0.f ? g.x : 0.f, o.y > 0.f ? g.y : 0.f, o.z > 0.f ? g.z : 0.f, o.w > 0.f ? g.w : 0.f);
            }
            if constexpr (RES) *reinterpret_cast<float4 *>(gres + i) = g;
            *reinterpret_cast<float4 *>(gx + i) = make_float4(__fmul_rn(g.x, s), __fmul_rn(g.y, s), __fmul_rn(g.z, s), __fmul_rn(g.w, s));
        } else {
            float g = gy[i];
            if constexpr (RELU) g = y[i] > 0.f ? g : 0.f;
            if constexpr (RES) gres[i] = g;
            gx[i] = __fmul_rn(g, s);
        }
    }
}

int grid_for(long n, int v)
{
    const long blocks = (n / v + 255) / 256;
    return (int)(blocks < 1 ? 1 : (blocks > 148L * 16 ? 148L * 16 : blocks));
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int detr_frozen_bn_act_f32(const float *x, const float *residual, const float *scale, const float *bias, float *y,
                           long long n, long long hw, int C, int relu, void *stream)
{
    if (n < 0 || hw <= 0 || C <= 0) return DETR_STEP_ERR_BAD_SHAPE;
    if (n == 0) return 0;
    if (!x || !scale || !bias || !y) return DETR_STEP_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (hw % 4 == 0) && aligned16(x) && aligned16(y) && (!residual || aligned16(residual));
#define BN_FWD(RELU, RES, V) \
    bn_act_fwd_kernel<RELU, RES, V><<<grid_for(n, V), 256, 0, st>>>(x, residual, scale, bias, y, (long)n, (long)hw, C)
    if (v4) {
        if (relu && residual) BN_FWD(true, true, 4); else if (relu) BN_FWD(true, false, 4);
        else if (residual) BN_FWD(false, true, 4); else BN_FWD(false, false, 4);
    } else {
        if (relu && residual) BN_FWD(true, true, 1); else if (relu) BN_FWD(true, false, 1);
        else if (residual) BN_FWD(false, true, 1); else BN_FWD(false, false, 1);
    }
#undef BN_FWD
    return (int)cudaGetLastError();
}

int detr_frozen_bn_act_backward_f32(const float *grad_y, const float *y, const float *scale, float *grad_x,
                                    float *grad_residual, long long n, long long hw, int C, int relu, void *stream)
{
    if (n < 0 || hw <= 0 || C <= 0) return DETR_STEP_ERR_BAD_SHAPE;
    if (n == 0) return 0;
    if (!grad_y || !scale || !grad_x || (relu && !y)) return DETR_STEP_ERR_NULL_POINTER;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v4 = (hw % 4 == 0) && aligned16(grad_y) && aligned16(grad_x) && (!relu || aligned16(y)) &&
                    (!grad_residual || aligned16(grad_residual));
#define BN_BWD(RELU, RES, V) \
    bn_act_bwd_kernel<RELU, RES, V><<<grid_for(n, V), 256, 0, st>>>(grad_y, y, scale, grad_x, grad_residual, (long)n, (long)hw, C)
    if (v4) {
        if (relu && grad_residual) BN_BWD(true, true, 4); else if (relu) BN_BWD(true, false, 4);
        else if (grad_residual) BN_BWD(false, true, 4); else BN_BWD(false, false, 4);
    } else {
        if (relu && grad_residual) BN_BWD(true, true, 1); else if (relu) BN_BWD(true, false, 1);
        else if (grad_residual) BN_BWD(false, true, 1); else BN_BWD(false, false, 1);
    }
#undef BN_BWD
    return (int)cudaGetLastError();
}

}  // extern "C"
