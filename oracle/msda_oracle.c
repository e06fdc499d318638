/*
 * msda_oracle.c -- CPU restatement of the reference MSDA algorithm.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (monosowa_b200/) never links, imports or falls back to anything in oracle/.
 *
 * What it restates (reference = jskvrna/MonoSOWA, paths relative to
 * MonoDETR/lib/models/monodetr/ops/):
 *   - sample coordinate + in-range test ....... src/cuda/ms_deform_im2col_cuda.cuh:285-291
 *   - bilinear gather with zero padding ....... src/cuda/ms_deform_im2col_cuda.cuh:33-84
 *   - weighted sum over levels x points ....... src/cuda/ms_deform_im2col_cuda.cuh:272-297
 *   - grad_value / grad_loc / grad_attn ....... src/cuda/ms_deform_im2col_cuda.cuh:87-159, 347-401
 *   - tensor layouts / index arithmetic ....... src/cuda/ms_deform_attn_cuda.cu:40-77
 * It is written as plain per-sample loops (one (n,q,m) triple at a time, all channels),
 * not as the reference's one-thread-per-channel kernels.
 *
 * Parity pin: checked in tests/test_oracle.py against tests/golden/*.npz, which were
 * produced by importing the reference's own ms_deform_attn_core_pytorch
 * (functions/ms_deform_attn_func.py:41-61) in the build container
 * (tests/golden/gen_golden.py).  See DESIGN.md "Oracle".
 *
 * Two arithmetic flavours per entry point:
 *   *_f64 : everything in double (the parity yardstick).
 *   *_f32 : the same loops in float, compiled with -ffp-contract=off, so that floor()
 *           decisions follow fp32 arithmetic exactly as the reference kernel's
 *           scalar_t=float instantiation does: the pixel coordinate is ONE fused
 *           multiply-add, fma(loc, H, -0.5) -- what nvcc makes of cuh:285-286 (FFMA / DFMA
 *           in the SASS of the compiled reference); everything else is not contracted.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_FMA_f64(a, b, c) fma((a), (b), (c))
#define ORACLE_FMA_f32(a, b, c) fmaf((a), (b), (c))

#define DEFINE_ORACLE(SUFFIX, real)                                                          \
                                                                                             \
/* one bilinear tap set; returns 1 when the sample is inside the (-1,H)x(-1,W) window */     \
static int tap_##SUFFIX(real loc_x, real loc_y, int H, int W,                                \
                        int *y0, int *x0, real *ly, real *lx)                                \
{                                                                                            \
    /* cuh:285-286 : pixel coordinate, align_corners=False convention */                     \
    real py = ORACLE_FMA_##SUFFIX(loc_y, (real)H, (real)-0.5);                               \
    real px = ORACLE_FMA_##SUFFIX(loc_x, (real)W, (real)-0.5);                               \
    /* cuh:288 */                                                                            \
    if (!(py > (real)-1 && px > (real)-1 && py < (real)H && px < (real)W)) return 0;         \
    real fy = floor(py), fx = floor(px);                                                     \
    *y0 = (int)fy; *x0 = (int)fx;                                                            \
    *ly = py - fy; *lx = px - fx;                                                            \
    return 1;                                                                                \
}                                                                                            \
                                                                                             \
void msda_oracle_forward_##SUFFIX(const real *value, const int64_t *shapes,                  \
                                  const int64_t *lsi, const real *loc, const real *attn,     \
                                  real *out, int N, int S, int M, int D, int L, int Lq,      \
                                  int P)                                                     \
{                                                                                            \
    const long nqm = (long)N * Lq * M;                                                       \
    _Pragma("omp parallel for schedule(static)")                                             \
    for (long t = 0; t < nqm; ++t) {                                                         \
        const int m = (int)(t % M);                                                          \
        const long n = t / ((long)Lq * M);                                                   \
        real *o = out + t * D;                                                               \
        for (int c = 0; c < D; ++c) o[c] = 0;                                                \
        const real *lp = loc + t * (long)L * P * 2;                                          \
        const real *ap = attn + t * (long)L * P;                                             \
        for (int l = 0; l < L; ++l) {                                                        \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                    \
            const real *vl = value + ((n * S + lsi[l]) * M + m) * (long)D;                   \
            const long xs = (long)M * D, ys = xs * W;                                        \
            for (int p = 0; p < P; ++p, lp += 2, ++ap) {                                     \
                int y0, x0; real ly, lx;                                                     \
                if (!tap_##SUFFIX(lp[0], lp[1], H, W, &y0, &x0, &ly, &lx)) continue;         \
                const real hy = 1 - ly, hx = 1 - lx;                                         \
                const real w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;       \
                const int oky0 = y0 >= 0, oky1 = y0 + 1 <= H - 1;                            \
                const int okx0 = x0 >= 0, okx1 = x0 + 1 <= W - 1;                            \
                const real *b = vl + y0 * ys + x0 * xs;                                      \
                const real a = *ap;                                                          \
                for (int c = 0; c < D; ++c) {                                                \
                    real v00 = (oky0 && okx0) ? b[c] : 0;                                    \
                    real v01 = (oky0 && okx1) ? b[xs + c] : 0;                               \
                    real v10 = (oky1 && okx0) ? b[ys + c] : 0;                               \
                    real v11 = (oky1 && okx1) ? b[ys + xs + c] : 0;                          \
                    o[c] += (w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11) * a;             \
                }                                                                            \
            }                                                                                \
        }                                                                                    \
    }                                                                                        \
}                                                                                            \
                                                                                             \
/* grad_value must be zero-filled by the caller (ms_deform_attn_cuda.cu:121).  The outer  */ \
/* loop is over images so that grad_value scatter needs no atomics: one thread per image. */ \
void msda_oracle_backward_##SUFFIX(const real *value, const int64_t *shapes,                 \
                                   const int64_t *lsi, const real *loc, const real *attn,    \
                                   const real *grad_out, real *grad_value, real *grad_loc,   \
                                   real *grad_attn, int N, int S, int M, int D, int L,       \
                                   int Lq, int P)                                            \
{                                                                                            \
    _Pragma("omp parallel for schedule(dynamic,1) collapse(2)")                              \
    for (long n = 0; n < N; ++n)                                                             \
    for (int m = 0; m < M; ++m)                                                              \
    for (long q = 0; q < Lq; ++q) {                                                          \
        const long t = (n * Lq + q) * M + m;                                                 \
        const real *g = grad_out + t * D;                                                    \
        const real *lp = loc + t * (long)L * P * 2;                                          \
        const real *ap = attn + t * (long)L * P;                                             \
        real *gl = grad_loc + t * (long)L * P * 2;                                           \
        real *ga = grad_attn + t * (long)L * P;                                              \
        for (int l = 0; l < L; ++l) {                                                        \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                    \
            const long base = ((n * S + lsi[l]) * M + m) * (long)D;                          \
            const long xs = (long)M * D, ys = xs * W;                                        \
            for (int p = 0; p < P; ++p, lp += 2, ++ap, gl += 2, ++ga) {                      \
                int y0, x0; real ly, lx;                                                     \
                gl[0] = 0; gl[1] = 0; ga[0] = 0;                     /* cuh:365-367 */       \
                if (!tap_##SUFFIX(lp[0], lp[1], H, W, &y0, &x0, &ly, &lx)) continue;         \
                const real hy = 1 - ly, hx = 1 - lx;                                         \
                const real w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;       \
                const int oky0 = y0 >= 0, oky1 = y0 + 1 <= H - 1;                            \
                const int okx0 = x0 >= 0, okx1 = x0 + 1 <= W - 1;                            \
                const long o00 = base + y0 * ys + x0 * xs;                                   \
                const real a = *ap;                                                          \
                real sx = 0, sy = 0, sa = 0;                                                 \
                for (int c = 0; c < D; ++c) {                                                \
                    const real tg = g[c] * a;                        /* cuh:111 */           \
                    real v00 = 0, v01 = 0, v10 = 0, v11 = 0;                                 \
                    if (oky0 && okx0) { v00 = value[o00 + c];                                \
                                        grad_value[o00 + c] += w00 * tg; }                   \
                    if (oky0 && okx1) { v01 = value[o00 + xs + c];                           \
                                        grad_value[o00 + xs + c] += w01 * tg; }              \
                    if (oky1 && okx0) { v10 = value[o00 + ys + c];                           \
                                        grad_value[o00 + ys + c] += w10 * tg; }              \
                    if (oky1 && okx1) { v11 = value[o00 + ys + xs + c];                      \
                                        grad_value[o00 + ys + xs + c] += w11 * tg; }         \
                    /* cuh:119-158: d(bilinear)/dx and /dy, then chain through W, H */       \
                    const real dgx = -hy * v00 + hy * v01 - ly * v10 + ly * v11;             \
                    const real dgy = -hx * v00 - lx * v01 + hx * v10 + lx * v11;             \
                    sx += (real)W * dgx * tg;                                                \
                    sy += (real)H * dgy * tg;                                                \
                    sa += g[c] * (w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11);            \
                }                                                                            \
                gl[0] = sx; gl[1] = sy; ga[0] = sa;                                          \
            }                                                                                \
        }                                                                                    \
    }                                                                                        \
}

DEFINE_ORACLE(f64, double)
DEFINE_ORACLE(f32, float)

int msda_oracle_abi_version(void) { return 1; }
