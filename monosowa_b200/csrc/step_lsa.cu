// step_lsa.cu -- group-wise Hungarian matching on the device (include/monodetr_step_b200.h).
//
// Replaces the host section of the reference matcher (MonoDETR/lib/models/monodetr/matcher.py:87-104):
// `C.cpu()` (a device synchronisation per decoder layer) followed by scipy.optimize.linear_sum_assignment
// once per image and query group.  Here one warp solves one (image, group) sub-problem with the classic
// shortest-augmenting-path Hungarian method (row by row, dual potentials u / v, O(n^2 m) with the column
// scans spread over the 32 lanes), entirely in shared memory, in fp64 like scipy's solver.  The smaller
// side of the rectangular problem provides the rows, so every row gets a column.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "monodetr_step_b200.h"

namespace {

thread_local char t_err[256] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

struct ImageTable {
    int toff[DETR_STEP_MAX_IMAGES];     // first cost column of the image
    int nt[DETR_STEP_MAX_IMAGES];       // its number of targets
    int ooff[DETR_STEP_MAX_IMAGES];     // first output pair of the image
};

constexpr unsigned kFull = 0xffffffffu;

// rows i in [1, n], columns j in [1, m]; `rows_are_targets` tells how (i, j) maps onto (query, target)
__global__ void __launch_bounds__(32)
group_lsa_kernel(const float *__restrict__ cost, const ImageTable tab, const int Q, const int T, const int groups,
                 const int nmax, const int mmax, int64_t *__restrict__ out_query, int64_t *__restrict__ out_target,
                 int *__restrict__ status)
{
    extern __shared__ double sm[];
    double *u = sm;                                  // [nmax + 1] row potentials
    double *v = u + (nmax + 1);                      // [mmax + 1] column potentials
    double *minv = v + (mmax + 1);                   // [mmax + 1] reduced cost of the best edge into each free column
    int *p = reinterpret_cast<int *>(minv + (mmax + 1));   // [mmax + 1] row matched to column j (0 = free)
    int *way = p + (mmax + 1);                       // [mmax + 1] previous column on the alternating path
    int *used = way + (mmax + 1);                    // [mmax + 1] column is in the tree
    int *rowcol = used + (mmax + 1);                 // [nmax + 1] column matched to row i (output pass)

    const int lane = threadIdx.x;
    const int b = blockIdx.x / groups, g = blockIdx.x % groups;
    const int nq = Q / groups, nt = tab.nt[b];
    const int k = min(nq, nt);
    if (k == 0) return;
    const bool rows_are_targets = nt <= nq;
    const int n = rows_are_targets ? nt : nq, m = rows_are_targets ? nq : nt;
    const float *cb = cost + ((long)b * Q + (long)g * nq) * T + tab.toff[b];      // cb[query * T + target]
    auto edge = [&](int i, int j) -> double {         // 1-based row / column
        return rows_are_targets ? (double)cb[(long)(j - 1) * T + (i - 1)] : (double)cb[(long)(i - 1) * T + (j - 1)];
    };

    for (int j = lane; j <= m; j += 32) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
    for (int i = lane; i <= n; i += 32) u[i] = 0.0;
    __syncwarp();

    for (int i = 1; i <= n; ++i) {
        for (int j = lane; j <= m; j += 32) { minv[j] = CUDART_INF; used[j] = 0; }
        if (lane == 0) p[0] = i;
        __syncwarp();
        int j0 = 0;
        for (;;) {
            if (lane == 0) used[j0] = 1;
            __syncwarp();
            const int i0 = p[j0];
            const double ui0 = u[i0];
            double delta = CUDART_INF;
            int j1 = 0x7fffffff;
            for (int j = lane + 1; j <= m; j += 32) {
                if (used[j]) continue;
                const double cur = edge(i0, j) - ui0 - v[j];
                if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                if (minv[j] < delta) { delta = minv[j]; j1 = j; }          // ascending j: first minimum wins
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const double od = __shfl_xor_sync(kFull, delta, off);
                const int oj = __shfl_xor_sync(kFull, j1, off);
                if (od < delta || (od == delta && oj < j1)) { delta = od; j1 = oj; }
            }
            // the scan above wrote minv[] / way[] from lane (j - 1) % 32; the loops below touch minv[j] from lane
            // j % 32: shuffles order nothing in memory, so the warp is fenced here (independent thread scheduling)
            __syncwarp();
            if (j1 == 0x7fffffff) {                   // only non-finite costs left: take any free column ...
                for (int j = 1; j <= m; ++j)
                    if (!used[j]) { j1 = j; break; }
                delta = 0.0;
                if (lane == 0) {
                    way[j1] = j0;
                    if (status) atomicOr(status, 1);  // ... and say so: scipy raises ValueError on NaN / Inf costs
                }
            }
            for (int j = lane; j <= m; j += 32) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            __syncwarp();
            j0 = j1;
            if (p[j0] == 0) break;
        }
        if (lane == 0) {                              // flip the alternating path
            do {
                const int jp = way[j0];
                p[j0] = p[jp];
                j0 = jp;
            } while (j0 != 0);
        }
        __syncwarp();
    }

    // pairs ordered by query index (scipy returns row_ind sorted; rows there are the queries)
    const long obase = (long)groups * tab.ooff[b] + (long)g * k;
    if (rows_are_targets) {
        int written = 0;
        for (int j0 = 1; j0 <= m; j0 += 32) {
            const int j = j0 + lane;
            const int row = (j <= m) ? p[j] : 0;
            const unsigned has = __ballot_sync(kFull, row != 0);
            if (row != 0) {
                const int pos = written + __popc(has & ((1u << lane) - 1u));
                out_query[obase + pos] = (long)g * nq + (j - 1);
                out_target[obase + pos] = row - 1;
            }
            written += __popc(has);
        }
    } else {
        for (int j = lane + 1; j <= m; j += 32)
            if (p[j] != 0) rowcol[p[j]] = j;
        __syncwarp();
        for (int i = lane + 1; i <= n; i += 32) {
            out_query[obase + i - 1] = (long)g * nq + (i - 1);
            out_target[obase + i - 1] = rowcol[i] - 1;
        }
    }
}

}  // namespace

extern "C" {

int detr_group_lsa_status_f32(const float *cost, const int *sizes, int B, int Q, int T, int groups, int64_t *out_query,
                              int64_t *out_target, int *status, void *stream);

int detr_group_lsa_f32(const float *cost, const int *sizes, int B, int Q, int T, int groups, int64_t *out_query,
                       int64_t *out_target, void *stream)
{
    return detr_group_lsa_status_f32(cost, sizes, B, Q, T, groups, out_query, out_target, nullptr, stream);
}

int detr_group_lsa_status_f32(const float *cost, const int *sizes, int B, int Q, int T, int groups, int64_t *out_query,
                              int64_t *out_target, int *status, void *stream)
{
    if (B < 0 || Q < 0 || T < 0 || groups < 1 || Q % groups != 0)
        return fail(DETR_STEP_ERR_BAD_SHAPE, "detr_group_lsa_f32: bad shape (B=%d Q=%d T=%d groups=%d)", B, Q, T, groups);
    if (B > DETR_STEP_MAX_IMAGES)
        return fail(DETR_STEP_ERR_UNSUPPORTED, "detr_group_lsa_f32: B=%d exceeds DETR_STEP_MAX_IMAGES", B);
    if (B == 0) { t_err[0] = 0; return 0; }
    if (!sizes) return fail(DETR_STEP_ERR_NULL_POINTER, "detr_group_lsa_f32: sizes is NULL");
    ImageTable tab;
    const int nq = Q / groups;
    int toff = 0, ooff = 0, nmax = 0, mmax = 0;
    for (int b = 0; b < B; ++b) {
        if (sizes[b] < 0) return fail(DETR_STEP_ERR_BAD_SHAPE, "detr_group_lsa_f32: sizes[%d] < 0", b);
        tab.toff[b] = toff; tab.nt[b] = sizes[b]; tab.ooff[b] = ooff;
        toff += sizes[b];
        const int n = sizes[b] < nq ? sizes[b] : nq, m = sizes[b] < nq ? nq : sizes[b];
        ooff += n;
        if (n > 0) { nmax = n > nmax ? n : nmax; mmax = m > mmax ? m : mmax; }
    }
    if (toff != T) return fail(DETR_STEP_ERR_BAD_SHAPE, "detr_group_lsa_f32: sum(sizes)=%d but T=%d", toff, T);
    if (ooff == 0) { t_err[0] = 0; return 0; }
    if (!cost || !out_query || !out_target) return fail(DETR_STEP_ERR_NULL_POINTER, "detr_group_lsa_f32: NULL device pointer");
    const size_t smem = (size_t)(nmax + 1) * (8 + 4) + (size_t)(mmax + 1) * (8 + 8 + 4 + 4 + 4) + 8;
    if (smem > 48 * 1024)
        return fail(DETR_STEP_ERR_UNSUPPORTED, "detr_group_lsa_f32: sub-problem %d x %d needs %zu B of shared memory", nmax, mmax, smem);
    group_lsa_kernel<<<(unsigned)(B * groups), 32, smem, (cudaStream_t)stream>>>(cost, tab, Q, T, groups, nmax, mmax,
                                                                               out_query, out_target, status);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(t_err, sizeof(t_err), "detr_group_lsa_f32: %s (cudaError %d)", cudaGetErrorString(e), (int)e);
        return (int)e;
    }
    t_err[0] = 0;
    return 0;
}

const char *detr_step_last_error(void) { return t_err; }

}  // extern "C"
