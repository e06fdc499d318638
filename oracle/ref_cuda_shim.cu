/*
 * ref_cuda_shim.cu -- C-ABI doorway onto the REFERENCE's own CUDA kernels, for use as the
 * on-GPU baseline ("the reference kernel recompiled for sm_100a") and as a
 * second parity checker.  TEST/BENCH INFRASTRUCTURE ONLY; never loaded by monosowa_b200/.
 *
 * No reference source is copied: the reference header is #included from where it lies
 * (/root/reference/MonoDETR/lib/models/monodetr/ops/src/cuda/ms_deform_im2col_cuda.cuh,
 * host dispatchers at :923-954 and :956-1327) at build time, and only the resulting
 * shared object (oracle/_ref/libmsda_ref_sm100.so, git-ignored) travels to the GPU box.
 * The reference's ATen host wrapper (ms_deform_attn_cuda.cu) is not used -- it does not
 * compile on torch >= 2 (value.type() dispatch) -- so this shim calls the templated
 * launchers directly with raw pointers, exactly as that wrapper does at :65-72 / :135-147,
 * including the zero-fill of the outputs it performs with at::zeros (:54, :121-123).
 */
#include <cstdint>
#include <cuda_runtime.h>
#include "cuda/ms_deform_im2col_cuda.cuh"

#define REF_ENTRY(SUFFIX, T)                                                                  \
extern "C" int msda_ref_forward_##SUFFIX(const void *value, const int64_t *shapes,            \
        const int64_t *lsi, const void *loc, const void *attn, void *out, int N, int S,       \
        int M, int D, int L, int Lq, int P, void *stream)                                     \
{                                                                                             \
    cudaStream_t st = (cudaStream_t)stream;                                                   \
    cudaMemsetAsync(out, 0, sizeof(T) * (size_t)N * Lq * M * D, st);                          \
    ms_deformable_im2col_cuda<T>(st, (const T *)value, shapes, lsi, (const T *)loc,           \
                                 (const T *)attn, N, S, M, D, L, Lq, P, (T *)out);            \
    return (int)cudaGetLastError();                                                           \
}                                                                                             \
extern "C" int msda_ref_backward_##SUFFIX(const void *value, const int64_t *shapes,           \
        const int64_t *lsi, const void *loc, const void *attn, const void *grad_out,          \
        void *grad_value, void *grad_loc, void *grad_attn, int N, int S, int M, int D,        \
        int L, int Lq, int P, void *stream)                                                   \
{                                                                                             \
    cudaStream_t st = (cudaStream_t)stream;                                                   \
    cudaMemsetAsync(grad_value, 0, sizeof(T) * (size_t)N * S * M * D, st);                    \
    cudaMemsetAsync(grad_loc, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2, st);             \
    cudaMemsetAsync(grad_attn, 0, sizeof(T) * (size_t)N * Lq * M * L * P, st);                \
    ms_deformable_col2im_cuda<T>(st, (const T *)grad_out, (const T *)value, shapes, lsi,      \
                                 (const T *)loc, (const T *)attn, N, S, M, D, L, Lq, P,       \
                                 (T *)grad_value, (T *)grad_loc, (T *)grad_attn);             \
    return (int)cudaGetLastError();                                                           \
}

REF_ENTRY(f32, float)
REF_ENTRY(f64, double)
