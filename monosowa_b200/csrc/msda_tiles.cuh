// msda_tiles.cuh -- query tiles for long query sets (the encoder: queries ARE the pixel pyramid).
// Used by the tile kernels of the measurement build only (-DMSDA_AB; DESIGN.md 4.3, profiles/r02_tile_kernels.md).
//
// Why: both directions of the op are bound by 128-byte ROWS moved between L2 and the SMs,
// not by HBM.  In MonoDETR's encoder (reference depthaware_transformer.py:363-376) query q is pixel q of the
// flattened pyramid and samples every level around its own normalised position.  A CTA that owns 256
// CONSECUTIVE queries owns a 160 x 1.6 pixel strip: its samples spread over a band of the whole image width
// and almost nothing is shared between its queries.  A CTA that owns a 2-D TILE of the image -- the
// 12 x 16 block of base-level pixels plus the pixels of the coarser levels whose centres fall into the same
// image region -- has all its samples inside a compact window per level: the rows it gathers are reused from
// L1, and the grad_value contributions of ALL levels meet often enough inside the window to be combined in
// the SM (msda_backward_tiled.cu).
//
// The plan is derived on the device from spatial_shapes (the host never reads them; launches stay
// CUDA-graph capturable), so the kernels run a persistent grid and every CTA walks work items
// item = blockIdx.x, blockIdx.x + gridDim.x, ...  with  item -> (image n, head m, tile t).
// Tiling is a performance hint only: any query set is processed correctly.  If the queries are not the
// pixel pyramid (sum_l H_l*W_l != Lq) tiles are runs of consecutive queries.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kTileH = 12;          // tile of the base (largest) level, in pixels
constexpr int kTileW = 16;
constexpr int kTileMaxDim = 16384;  // sanity bound on H, W (keeps all integer arithmetic below 2^31)

struct TilePlan {
    int grid_mode;                  // 1: queries are the pixel pyramid described by spatial_shapes
    int base;                       // tiling base level (largest H*W)
    int nty, ntx;                   // tiles along y / x of the base level
    int n_tiles;                    // tiles per (image, head)
    int linear_q;                   // linear mode: queries per tile
    int qstart[MSDA_MAX_LEVELS];    // grid mode: query index of the first pixel of each level
};

// the sub-rectangle of level l's pixel grid that belongs to a tile, and the running query count
struct TileRect {
    int y0, x0, h, w;
    int qbase;                      // local index of the rectangle's first query inside the tile
};

struct TileItem {
    int n, m, ty, tx;
    int nq;                         // queries of the tile
    int q0;                         // linear mode: first query
    TileRect r[MSDA_MAX_LEVELS];
};

__host__ __device__ __forceinline__ int floor_div(long a, long b)      // b > 0
{
    long q = a / b;
    if ((a % b != 0) && (a < 0)) --q;
    return (int)q;
}

// first pixel of a level (extent `size`) whose centre lies at or beyond tile boundary t of the base level
// (extent `bsize`, tile extent `tdim`): smallest y with (y + 0.5) / size >= t * tdim / bsize.
__host__ __device__ __forceinline__ int tile_lo(int t, int n_t, int tdim, int size, int bsize)
{
    if (t <= 0) return 0;
    if (t >= n_t) return size;
    const long num = 2L * t * tdim * size - bsize;              // y >= num / (2 * bsize)
    int y = floor_div(num + 2L * bsize - 1, 2L * bsize);        // ceil
    return y < 0 ? 0 : (y > size ? size : y);
}

// thread 0 of the CTA builds the plan (s_lv staged before); callers __syncthreads() afterwards
__device__ __forceinline__ void make_tile_plan(TilePlan &p, const LevelInfo *s_lv, const Dims &d, int linear_q)
{
    long sum = 0;
    long best_area = -1;
    bool sane = d.L >= 1;
    p.base = 0;
    for (int l = 0; l < d.L; ++l) {
        const int H = s_lv[l].H, W = s_lv[l].W;
        if (H <= 0 || W <= 0 || H > kTileMaxDim || W > kTileMaxDim) sane = false;
        p.qstart[l] = (int)sum;
        const long area = (long)H * W;
        sum += sane ? area : 0;
        if (sane && area > best_area) { best_area = area; p.base = l; }
    }
    p.grid_mode = (sane && sum == (long)d.Lq) ? 1 : 0;
    p.linear_q = linear_q;
    if (p.grid_mode) {
        p.nty = (s_lv[p.base].H + kTileH - 1) / kTileH;
        p.ntx = (s_lv[p.base].W + kTileW - 1) / kTileW;
        p.n_tiles = p.nty * p.ntx;
    } else {
        p.nty = 1;
        p.ntx = (d.Lq + linear_q - 1) / linear_q;
        p.n_tiles = p.ntx;
    }
}

// Decode work item -> (n, m, tile) and the tile's per-level rectangles.  Executed by one thread.
__device__ __forceinline__ void make_tile_item(TileItem &it, const TilePlan &p, const LevelInfo *s_lv, const Dims &d,
                                               long item)
{
    const int t = (int)(item % p.n_tiles);
    const long r = item / p.n_tiles;
    it.m = (int)(r % d.M);
    it.n = (int)(r / d.M);
    it.ty = t / p.ntx;
    it.tx = t - it.ty * p.ntx;
    if (!p.grid_mode) {
        it.q0 = t * p.linear_q;
        it.nq = min(p.linear_q, d.Lq - it.q0);
        return;
    }
    it.q0 = 0;
    const int Hb = s_lv[p.base].H, Wb = s_lv[p.base].W;
    int count = 0;
    for (int l = 0; l < d.L; ++l) {
        const int H = s_lv[l].H, W = s_lv[l].W;
        TileRect rc;
        rc.y0 = tile_lo(it.ty, p.nty, kTileH, H, Hb);
        rc.x0 = tile_lo(it.tx, p.ntx, kTileW, W, Wb);
        rc.h = tile_lo(it.ty + 1, p.nty, kTileH, H, Hb) - rc.y0;
        rc.w = tile_lo(it.tx + 1, p.ntx, kTileW, W, Wb) - rc.x0;
        rc.qbase = count;
        count += rc.h * rc.w;
        it.r[l] = rc;
    }
    it.nq = count;
}

// local query index of a tile -> global query index (i < it.nq)
__device__ __forceinline__ int tile_query(const TileItem &it, const TilePlan &p, const LevelInfo *s_lv, int L, int i)
{
    if (!p.grid_mode) return it.q0 + i;
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < L; ++k)
        if (i >= it.r[k].qbase) l = k;
    // qbase is non-decreasing and i < nq, so the LAST level with qbase <= i is the (non-empty) owner
    const TileRect rc = it.r[l];
    const int j = i - rc.qbase;
    const int y = j / rc.w;
    const int x = j - y * rc.w;
    return p.qstart[l] + (rc.y0 + y) * s_lv[l].W + rc.x0 + x;
}

// persistent grid: `ctas_per_sm` CTAs for every SM of the current device (SM count cached per device)
inline int persistent_grid(int ctas_per_sm)
{
    static int sm_count[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148 * ctas_per_sm;
    if (sm_count[dev] <= 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        sm_count[dev] = sms;
    }
    return sm_count[dev] * ctas_per_sm;
}

}  // namespace msda
