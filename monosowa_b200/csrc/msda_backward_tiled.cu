// msda_backward_tiled.cu -- MSDA backward for long query sets (the encoder): the grad_value contributions of
// ALL feature levels are combined inside the SM before they leave it.
//
// Why: the direct backward (bwd_rec_kernel, msda_backward.cu) sends one 128-byte
// reduction line to L2 per (sample, corner), and L2 absorbs only ~49 G such lines per second chip-wide --
// at the KITTI encoder shape that alone is 1.4 ms.  Contributions can only be combined where they MEET, and
// they meet inside a CTA only if the CTA's queries are neighbours in the image in BOTH directions: a CTA owns a
// 2-D image tile (msda_tiles.cuh; 12 x 16 base-level pixels plus the coarser levels' pixels of the same region,
// ~255 queries), whose samples fall into a compact window of every level (tile + a 6-pixel halo).
//
// Per work item (image, head, tile), per batch of G samples of every query (G = D/4 lanes; at D = 32, P = 4 a
// batch is two levels):
//   phase A  the record kernel's work: gathers, partial dots, grad_loc / grad_attn.  A sample whose base-corner
//            cell lies inside its level's window is not reduced to L2: the owning lane parks a 16-byte entry
//            {a, lx, ly, cell | rank} in shared memory and counts it in a histogram over the window cells (the
//            base corner (y0, x0) ranges over the (H+1) x (W+1) lattice [-1,H-1] x [-1,W-1]).  Strays (outside
//            every window: large offsets, non-pyramid query sets) take the direct REDG path.
//   phase B  exclusive scan of the histogram, counting-sort permutation of the entry indices.
//   phase C  every lane group walks a chunk of the sorted entries: all entries of a cell share their four corner
//            pixels, so the group reads each entry's grad_out row once (L1 -- phase A just read it), accumulates the
//            four corner rows ((wy*wx)*a)*g in registers and carries the corner column that two neighbouring cells
//            share: two REDG lines per cell plus two per run.
// Windows are sized on the device from spatial_shapes (the host never reads them); queries that are not the
// pixel pyramid are processed in runs of consecutive queries with whole-level windows for the coarse levels.
// Arithmetic: same products as the direct path, ((wy*wx)*a)*g, accumulated per cell corner in fp32 registers
// before one global reduction per corner (the reference accumulates every contribution with a global atomic,
// ms_deform_im2col_cuda.cuh:125-152) -- covered by the backward tolerance, which already allows for atomic
// ordering.  grad_loc / grad_attn are bitwise identical to bwd_rec_kernel (same phase-A arithmetic; tested).
//
// STATUS (round 2, measured on B200 at configs[1] and the configs[4] shapes, profiles/r02_tile_kernels.md): the
// reductions that reach L2 fall from 46.7 M (binned kernel) to 8-15 M lines and the L1 hit rate of the gathers
// rises from 42 % to 57-70 %, but the kernel takes 1.56-1.63 ms against the binned kernel's 1.44-1.46 ms: it is
// bound by instruction issue (1.22 G warp instructions, 65 % issue-active with 24 warps/SM; phase C alone is 50 % of
// them at ~1.5 entries per cell on the fine levels), not by L2 any more.  It is therefore NOT part of the product
// library: it is compiled only into the measurement build (-DMSDA_AB, libmsda_b200_ab.so; bwd_variant = 20), where
// the parity tests keep it covered.
#include "msda_common.cuh"
#include "msda_records.cuh"
#include "msda_tiles.cuh"

namespace msda {

#ifdef MSDA_AB
namespace {

constexpr int kTileThreads = 256;
constexpr int kMaxBins = 1280;        // window cells over the levels of one batch
constexpr int kHalo = 6;              // window = tile region + kHalo pixels on every side (per level)
constexpr unsigned kNoEntry = 0xffffffffu;

template <int D>
struct TileCfg {
    static constexpr int G = D / kChannelsPerLane;
    static constexpr int QPW = 32 / G;
    static constexpr int WARPS = kTileThreads / 32;
    static constexpr int QPI = WARPS * QPW;                        // queries per pass of the CTA's warps
    static constexpr int QC = D <= 32 ? 256 : 128;                 // queries per round (entry capacity)
    static constexpr int HIST_HALVES = kMaxBins + 2;               // 16-bit counters, packed two per word
    static constexpr int HIST_WORDS = (HIST_HALVES + 1) / 2;
    static constexpr int ENT_BYTES = QC * G * 16;
    static constexpr int HIST_BYTES = ((HIST_WORDS * 4 + 15) / 16) * 16;
    static constexpr int REC_BYTES = WARPS * RecordLayout<G>::WARP_WORDS * 4;
    static constexpr int BIN_BYTES = kMaxBins * 8;                 // per cell: pixel index of its base corner, W | flags
    static constexpr int SMEM = ENT_BYTES + HIST_BYTES + REC_BYTES + BIN_BYTES;
    static_assert(REC_BYTES >= QC * G * 2, "the sorted index array aliases the record area");
    static_assert(QC % QPI == 0, "a round must be a whole number of passes");
    static_assert(QC * G < 65536, "entry indices are 16-bit");
};

// window of one level on its base-corner lattice; nb = 0: the level has no window in this batch
struct LevelWin {
    int cy0, cx0, wh, ww, binbase, nb;
};

// per-cell flags (bits 16.. of the second word of the cell table; the low 16 bits hold W)
constexpr unsigned kCellY0 = 1u << 16, kCellY1 = 1u << 17, kCellX0 = 1u << 18, kCellX1 = 1u << 19;
constexpr unsigned kCellRowEnd = 1u << 20;        // last cell of its window row: the next cell does not share a corner column
constexpr int kChunk = 8;                         // phase C: sorted entries per lane-group task

// window cells [lo, hi] along one axis: the base corners of samples whose pixel coordinate lies within kHalo
// pixels of the tile's extent [t0, t1) (in base-level pixels) mapped to this level
__device__ __forceinline__ void window_range(int t0, int t1, int size, int bsize, int &lo, int &hi)
{
    // pixel coordinate p = u * size - 0.5 for u in [t0 / bsize, t1 / bsize); base corner = floor(p +- halo)
    lo = floor_div(2L * t0 * size - bsize - 2L * kHalo * bsize, 2L * bsize) + 1;
    hi = floor_div(2L * t1 * size - bsize + 2L * kHalo * bsize, 2L * bsize) + 1;
    lo = max(lo, 0);
    hi = min(hi, size);                                   // lattice coordinate of corner size-1 is `size`
}

template <typename VT, int D, bool FUSED, int MINB>
__global__ void __launch_bounds__(kTileThreads, MINB)
bwd_tile_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
                const float *__restrict__ loc, const float *__restrict__ attn, const VT *__restrict__ grad_out,
                float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn,
                const Dims d, const float *__restrict__ ref, const int ref_dim, const int first_binned)
{
    using C = TileCfg<D>;
    using RL = RecordLayout<C::G>;
    constexpr int G = C::G, QPW = C::QPW, QPI = C::QPI, QC = C::QC;

    extern __shared__ __align__(16) unsigned char smem[];
    uint4 *s_ent = reinterpret_cast<uint4 *>(smem);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + C::ENT_BYTES);
    uint32_t *s_rec = reinterpret_cast<uint32_t *>(smem + C::ENT_BYTES + C::HIST_BYTES);
    int2 *s_cell = reinterpret_cast<int2 *>(smem + C::ENT_BYTES + C::HIST_BYTES + C::REC_BYTES);
    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    __shared__ TilePlan s_plan;
    __shared__ TileItem s_item;
    __shared__ LevelWin s_win[MSDA_MAX_LEVELS];
    __shared__ int s_q[QC];                              // global query index of the round's local queries
    __shared__ int s_warp_tot[kTileThreads / 32];
    __shared__ int s_nbins;

    const int tid = threadIdx.x;
    stage_levels(s_lv, shapes, lsi, d.L);
    if (tid == 0) make_tile_plan(s_plan, s_lv, d, QC);
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int gl = lane % G, k = lane / G;
    const int LP = d.L * d.P;
    const int n_batches = (LP + G - 1) / G;
    const int xs = d.M * D;
    uint32_t *grp = s_rec + warp * RL::WARP_WORDS + k * RL::GROUP_WORDS;
    unsigned short *hh = reinterpret_cast<unsigned short *>(s_hist);
    unsigned short *s_idx = reinterpret_cast<unsigned short *>(s_rec);    // phases B/C: the record area is free
    const long n_items = (long)d.N * d.M * s_plan.n_tiles;

    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();                                        // everyone is done with the previous item
        if (tid == 0) make_tile_item(s_item, s_plan, s_lv, d, item);
        __syncthreads();
        const int n = s_item.n, m = s_item.m;
        const long img = ((long)n * d.S * d.M + m) * D + gl * kChannelsPerLane;
        const VT *vimg = value + img;
        float *gvimg = grad_value + img;
        const long qrow0 = (long)n * d.Lq;                      // (qrow0 + q) * M + m = the (query, head) row

        for (int r0 = 0; r0 < s_item.nq; r0 += QC) {            // rounds of at most QC queries (entry capacity)
            const int nqr = min(QC, s_item.nq - r0);
            for (int i = tid; i < nqr; i += kTileThreads) s_q[i] = tile_query(s_item, s_plan, s_lv, d.L, r0 + i);

            for (int batch = 0; batch < n_batches; ++batch) {
                const int b0 = batch * G;
                const int lfirst = b0 / d.P, llast = min(d.L - 1, (b0 + G - 1) / d.P);
                // ---- windows of this batch's levels, histogram reset -------------------------------------
                __syncthreads();                                // phase C of the previous batch is over
                for (int i = tid; i < C::HIST_WORDS; i += kTileThreads) s_hist[i] = 0u;
                if (tid == 0) {
                    int nbins = 0;
                    for (int l = llast; l >= lfirst; --l) {     // coarsest first: the densest windows get the budget
                        const int H = s_lv[l].H, W = s_lv[l].W;
                        LevelWin wn{0, 0, 0, 0, nbins, 0};
                        if (s_plan.grid_mode) {
                            const int Hb = s_lv[s_plan.base].H, Wb = s_lv[s_plan.base].W;
                            int ylo, yhi, xlo, xhi;
                            window_range(s_item.ty * kTileH, min((s_item.ty + 1) * kTileH, Hb), H, Hb, ylo, yhi);
                            window_range(s_item.tx * kTileW, min((s_item.tx + 1) * kTileW, Wb), W, Wb, xlo, xhi);
                            wn.cy0 = ylo; wn.cx0 = xlo; wn.wh = yhi - ylo + 1; wn.ww = xhi - xlo + 1;
                        } else if (H > 0 && W > 0 && H <= kTileMaxDim && W <= kTileMaxDim && (long)H * W <= (long)QC * d.P) {
                            wn.wh = H + 1; wn.ww = W + 1;       // whole lattice of a coarse level
                        }
                        const long cells = (long)wn.wh * wn.ww;
                        if (wn.wh > 0 && wn.ww > 0 && nbins + cells <= kMaxBins) {
                            wn.nb = (int)cells;
                            nbins += wn.nb;
                        }
                        s_win[l] = wn;
                    }
                    s_nbins = nbins;
                }
                __syncthreads();
                // cell table: everything phase C needs to flush a cell, so that it never divides or searches
                for (int c = tid; c < s_nbins; c += kTileThreads) {
                    int l = lfirst;
                    while (l < llast && (unsigned)(c - s_win[l].binbase) >= (unsigned)s_win[l].nb) ++l;
                    const LevelWin wn = s_win[l];
                    const LevelInfo li = s_lv[l];
                    const int cc = c - wn.binbase;
                    const int by = cc / wn.ww, bx = cc - by * wn.ww;
                    const int cy = wn.cy0 + by, cx = wn.cx0 + bx;                          // lattice coordinates
                    unsigned f = (unsigned)li.W;
                    if (cy - 1 >= 0) f |= kCellY0;
                    if (cy <= li.H - 1) f |= kCellY1;
                    if (cx - 1 >= 0) f |= kCellX0;
                    if (cx <= li.W - 1) f |= kCellX1;
                    if (bx + 1 == wn.ww) f |= kCellRowEnd;
                    s_cell[c] = make_int2(li.start + (cy - 1) * li.W + (cx - 1), (int)f);  // pixel (y0, x0)
                }

                // ---- phase A ---------------------------------------------------------------------------
                const int sidx = b0 + gl;
                const int l = min(sidx / d.P, d.L - 1);
                SampleIn in_next{0.f, 0.f, 0.f, 0.f, 0.f};       // unfused: the next pass's sample, fetched a pass ahead
                if constexpr (!FUSED) {
                    const int ql = warp * QPW + k;
                    if (ql < nqr && sidx < LP) in_next = fetch_sample(true, loc, attn, ((qrow0 + s_q[ql]) * d.M + m) * LP + sidx);
                }
                for (int i0 = warp * QPW; i0 < nqr; i0 += QPI) {
                    const int ql = i0 + k;
                    const bool qvalid = ql < nqr;
                    const long qm = (qrow0 + s_q[qvalid ? ql : 0]) * d.M + m;
                    const bool has = qvalid && sidx < LP;

                    float g[4];
                    Vec4<VT>::load(grad_out + qm * D + gl * kChannelsPerLane, g);      // read again by phase C: keep it cached
                    SampleIn in;
                    if constexpr (FUSED) {
                        float aw[kMaxBatches];
                        group_softmax<G>(attn, qm * LP, LP, gl, qvalid, aw);
                        const float a = batch == 0 ? aw[0] : (batch == 1 ? aw[1] : (batch == 2 ? aw[2] : aw[3]));
                        in = fetch_sample_fused<stream_policy<G, VT>()>(has, loc, ref, ref_dim, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv, l, d.P, a);
                    } else {
                        in = in_next;
                    }
                    int4 off;
                    float4 wa;
                    const SampleGeom gm = sample_geometry(has && d.S > 0, in, s_lv, l, xs, off, wa);
                    bool stray = false;
                    if (qvalid) {
                        uint4 e = make_uint4(0u, 0u, 0u, kNoEntry);
                        if (gm.live && gm.a != 0.f) {
                            const LevelWin wn = s_win[l];
                            const int by = gm.cy - wn.cy0, bx = gm.cx - wn.cx0;
                            if (wn.nb > 0 && sidx >= first_binned && (unsigned)by < (unsigned)wn.wh && (unsigned)bx < (unsigned)wn.ww) {
                                // inside the window: park the sample instead of sending four reduction lines to L2
                                const int bin = wn.binbase + by * wn.ww + bx;
                                const unsigned sh = (bin & 1) * 16;
                                const unsigned old = atomicAdd(&s_hist[bin >> 1], 1u << sh);
                                e = make_uint4(__float_as_uint(gm.a), __float_as_uint(gm.lx), __float_as_uint(gm.ly),
                                               (unsigned)bin | (((old >> sh) & 0xffffu) << 16));
                                wa = make_float4(0.f, 0.f, 0.f, 0.f);
                            } else {
                                stray = true;
                            }
                        }
                        s_ent[ql * G + gl] = e;
                    }
                    if (!gm.live) off.x = -1;                    // consumers skip the gathers of this sample
                    *reinterpret_cast<int4 *>(grp + gl * 4) = off;
                    *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + gl * 4) = wa;
                    const unsigned smask = __ballot_sync(kFullMask, stray) >> (k * G);   // bit s: sample s of MY group is a stray
                    __syncwarp();
                    if constexpr (!FUSED) {                      // in flight while this pass gathers
                        const int qn = ql + QPI;
                        const bool hn = qn < nqr && sidx < LP;
                        in_next = fetch_sample(hn, loc, attn, ((qrow0 + s_q[hn ? qn : 0]) * d.M + m) * LP + sidx);
                    }

                    // The G samples of the batch are consumed in two halves so that only 2*G partial dot products
                    // are live at a time.  After the reduce-scatter of a half, lane j holds corner pair (j & 1) of the
                    // half's sample j / 2; the owner of sample s then pulls its four totals from lanes 2*(s % (G/2))
                    // and +1 of half s / (G/2).
                    constexpr int GH = G / 2;
                    float th[2][2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float t[4 * GH];
#pragma unroll
                        for (int u = 0; u < GH; ++u) {
                            const int s = h * GH + u;
                            t[4 * u] = t[4 * u + 1] = t[4 * u + 2] = t[4 * u + 3] = 0.f;
                            const int4 o4 = *reinterpret_cast<const int4 *>(grp + s * 4);
                            if (o4.x >= 0) {
                                float v00[4], v01[4], v10[4], v11[4];
                                Vec4<VT>::template gather<0>(vimg + o4.x, v00);
                                Vec4<VT>::template gather<0>(vimg + o4.y, v01);
                                Vec4<VT>::template gather<0>(vimg + o4.z, v10);
                                Vec4<VT>::template gather<0>(vimg + o4.w, v11);
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    t[4 * u] += g[c] * v00[c];
                                    t[4 * u + 1] += g[c] * v01[c];
                                    t[4 * u + 2] += g[c] * v10[c];
                                    t[4 * u + 3] += g[c] * v11[c];
                                }
                            }
                            if ((smask >> s) & 1u) {             // stray: direct reductions, zero weights skipped
                                const float4 w4 = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
                                if (w4.x != 0.f) red_add_f32x4(gvimg + o4.x, w4.x * g[0], w4.x * g[1], w4.x * g[2], w4.x * g[3]);
                                if (w4.y != 0.f) red_add_f32x4(gvimg + o4.y, w4.y * g[0], w4.y * g[1], w4.y * g[2], w4.y * g[3]);
                                if (w4.z != 0.f) red_add_f32x4(gvimg + o4.z, w4.z * g[0], w4.z * g[1], w4.z * g[2], w4.z * g[3]);
                                if (w4.w != 0.f) red_add_f32x4(gvimg + o4.w, w4.w * g[0], w4.w * g[1], w4.w * g[2], w4.w * g[3]);
                            }
                        }
                        group_reduce_scatter<G, 4 * GH>(t, gl);
                        th[h][0] = t[0];
                        th[h][1] = t[1];
                    }
                    __syncwarp();

                    float t[4];
                    {
                        const int src = (lane & ~(G - 1)) | (2 * (gl % GH));
                        const bool second = gl >= GH;
                        const float a0 = __shfl_sync(kFullMask, th[0][0], src), a1 = __shfl_sync(kFullMask, th[0][1], src);
                        const float a2 = __shfl_sync(kFullMask, th[0][0], src + 1), a3 = __shfl_sync(kFullMask, th[0][1], src + 1);
                        const float c0 = __shfl_sync(kFullMask, th[1][0], src), c1 = __shfl_sync(kFullMask, th[1][1], src);
                        const float c2 = __shfl_sync(kFullMask, th[1][0], src + 1), c3 = __shfl_sync(kFullMask, th[1][1], src + 1);
                        t[0] = second ? c0 : a0; t[1] = second ? c1 : a1; t[2] = second ? c2 : a2; t[3] = second ? c3 : a3;
                    }
                    if (has) {
                        float gx = 0.f, gy = 0.f, ga = 0.f;
                        if (gm.live) {
                            // corners outside the map were read from a clamped address; the reference counts 0
                            const float t00 = (gm.vmask & 1u) ? t[0] : 0.f, t01 = (gm.vmask & 2u) ? t[1] : 0.f;
                            const float t10 = (gm.vmask & 4u) ? t[2] : 0.f, t11 = (gm.vmask & 8u) ? t[3] : 0.f;
                            ga = gm.w00 * t00 + gm.w01 * t01 + gm.w10 * t10 + gm.w11 * t11;
                            gx = gm.Wf * gm.a * (gm.hy * (t01 - t00) + gm.ly * (t11 - t10));
                            gy = gm.Hf * gm.a * (gm.hx * (t10 - t00) + gm.lx * (t11 - t01));
                        }
                        const long si = qm * LP + sidx;
                        if constexpr (FUSED) {
                            const float2 go = gm.live ? fused_offset_grad(ref_dim, gx, gy, gm.Wf, gm.Hf, in.ex, in.ey, d.P)
                                                      : make_float2(0.f, 0.f);
                            __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), go);
                            grad_attn[si] = ga;                  // d out / d a; the softmax backward follows below
                        } else {
                            __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), make_float2(gx, gy));
                            __stcs(grad_attn + si, ga);
                        }
                    }
                }

                const int nbins = s_nbins;
                if (nbins == 0) continue;                        // block-uniform: nothing was parked
                __syncthreads();

                // ---- phase B: histogram -> exclusive offsets (in place), then the counting-sort permutation ----
                {
                    constexpr int BPT = (C::HIST_HALVES + kTileThreads - 1) / kTileThreads;
                    int c[BPT], sum = 0;
#pragma unroll
                    for (int i = 0; i < BPT; ++i) {
                        const int idx = tid * BPT + i;
                        c[i] = idx < C::HIST_HALVES ? hh[idx] : 0;
                        sum += c[i];
                    }
                    int incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(kFullMask, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (lane == 31) s_warp_tot[warp] = incl;
                    __syncthreads();
                    int base = incl - sum;
                    for (int w2 = 0; w2 < warp; ++w2) base += s_warp_tot[w2];
#pragma unroll
                    for (int i = 0; i < BPT; ++i) {
                        const int idx = tid * BPT + i;
                        if (idx < C::HIST_HALVES) hh[idx] = (unsigned short)base;
                        base += c[i];
                    }
                }
                __syncthreads();
                for (int e = tid; e < nqr * G; e += kTileThreads) {
                    const unsigned pk = s_ent[e].w;
                    if (pk != kNoEntry) s_idx[hh[pk & 0xffffu] + (pk >> 16)] = (unsigned short)e;
                }
                __syncthreads();

                // ---- phase C: every lane group takes kChunk consecutive entries of the SORTED list --------------------
                // Entries of one cell are neighbours in the list and share their four corner pixels: the group reads each
                // entry's grad_out row once (L1 -- phase A just read it; a pixel-owner formulation reads it four times --
                // measured: more L1 data-pipe wavefronts than the reductions it replaces), accumulates the four corner rows
                // in registers and flushes them when the cell changes.  Consecutive cells of a window row share a corner
                // column, which is carried over: two REDG lines per cell plus two per run of adjacent cells.  All groups of
                // a warp walk the same number of entries.  Measured alternatives (profiles/r02_phase_c_formulations.md):
                // one group per run of CELLS (four diverged groups per warp, a division per cell: 45 % of the kernel's
                // instructions); one warp per chunk with one group per CORNER (convergent, but one entry at a time per
                // warp: +23 % instructions, 1.92 instead of 1.62 ms).
                {
                    const int total = hh[nbins];
                    const VT *gq_head = grad_out + (qrow0 * d.M + m) * D + gl * kChannelsPerLane;
                    for (int t0 = warp * QPW * kChunk; t0 < total; t0 += QPI * kChunk) {
                        const int ibeg = t0 + k * kChunk, iend = min(ibeg + kChunk, total);
                        float l0[4] = {0.f, 0.f, 0.f, 0.f}, l1[4] = {0.f, 0.f, 0.f, 0.f};      // corners (y0, x0), (y1, x0)
                        float r0a[4] = {0.f, 0.f, 0.f, 0.f}, r1a[4] = {0.f, 0.f, 0.f, 0.f};    // corners (y0, x1), (y1, x1)
                        bool lt = false, rt = false;                                           // accumulators hold something
                        int cur = -1;
                        int2 cell = make_int2(0, 0);
                        auto flush_left = [&]() {
                            const unsigned f = (unsigned)cell.y;
                            if (lt && (f & kCellX0)) {
                                float *row = gvimg + (long)cell.x * xs;
                                if (f & kCellY0) red_add_f32x4(row, l0[0], l0[1], l0[2], l0[3]);
                                if (f & kCellY1) red_add_f32x4(row + (long)(f & 0xffffu) * xs, l1[0], l1[1], l1[2], l1[3]);
                            }
                        };
                        auto flush_right = [&]() {
                            const unsigned f = (unsigned)cell.y;
                            if (rt && (f & kCellX1)) {
                                float *row = gvimg + (long)(cell.x + 1) * xs;
                                if (f & kCellY0) red_add_f32x4(row, r0a[0], r0a[1], r0a[2], r0a[3]);
                                if (f & kCellY1) red_add_f32x4(row + (long)(f & 0xffffu) * xs, r1a[0], r1a[1], r1a[2], r1a[3]);
                            }
                        };
                        // software pipeline: the next entry and its grad_out row are in flight while this one accumulates
                        uint4 en = make_uint4(0u, 0u, 0u, 0u);
                        float gt[4] = {0.f, 0.f, 0.f, 0.f};
                        if (ibeg < iend) {
                            const int e = s_idx[ibeg];
                            en = s_ent[e];
                            Vec4<VT>::load(gq_head + (long)s_q[e / G] * xs, gt);
                        }
#pragma unroll 1
                        for (int i = ibeg; i < t0 + (k + 1) * kChunk; ++i) {
                            const bool valid = i < iend;
                            uint4 en_n = en;
                            float gn[4] = {0.f, 0.f, 0.f, 0.f};
                            if (i + 1 < iend) {
                                const int e = s_idx[i + 1];
                                en_n = s_ent[e];
                                Vec4<VT>::load(gq_head + (long)s_q[e / G] * xs, gn);
                            }
                            if (valid) {
                                const int bin = (int)(en.w & 0xffffu);
                                if (bin != cur) {
                                    if (cur >= 0) {
                                        flush_left();
                                        if (bin == cur + 1 && !((unsigned)cell.y & kCellRowEnd)) {       // shares a corner column
#pragma unroll
                                            for (int k2 = 0; k2 < 4; ++k2) { l0[k2] = r0a[k2]; l1[k2] = r1a[k2]; r0a[k2] = 0.f; r1a[k2] = 0.f; }
                                            lt = rt;
                                            rt = false;
                                        } else {
                                            flush_right();
#pragma unroll
                                            for (int k2 = 0; k2 < 4; ++k2) l0[k2] = l1[k2] = r0a[k2] = r1a[k2] = 0.f;
                                            lt = rt = false;
                                        }
                                    }
                                    cur = bin;
                                    cell = s_cell[bin];
                                }
                                const float a = __uint_as_float(en.x), lx = __uint_as_float(en.y), ly = __uint_as_float(en.z);
                                const float hy = 1.f - ly, hx = 1.f - lx;
                                const float w00 = (hy * hx) * a, w01 = (hy * lx) * a, w10 = (ly * hx) * a, w11 = (ly * lx) * a;
#pragma unroll
                                for (int k2 = 0; k2 < 4; ++k2) {
                                    l0[k2] += w00 * gt[k2];
                                    r0a[k2] += w01 * gt[k2];
                                    l1[k2] += w10 * gt[k2];
                                    r1a[k2] += w11 * gt[k2];
                                }
                                lt = rt = true;
                            }
                            en = en_n;
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) gt[k2] = gn[k2];
                        }
                        if (cur >= 0) {
                            flush_left();
                            flush_right();
                        }
                    }
                }
            }

            if constexpr (FUSED) {
                // softmax backward over the L*P samples of every (query, head) of the round:
                // grad_logit_i = a_i * (g_i - sum_j a_j g_j), with g = d out / d a parked in grad_attn by phase A
                __syncthreads();                                // phase A's global writes are visible to the CTA
                for (int i0 = warp * QPW; i0 < nqr; i0 += QPI) {
                    const int ql = i0 + k;
                    const bool qvalid = ql < nqr;
                    const long qm = (qrow0 + s_q[qvalid ? ql : 0]) * d.M + m;
                    float aw[kMaxBatches], gg[kMaxBatches];
                    group_softmax<G>(attn, qm * LP, LP, gl, qvalid, aw);
                    float dot = 0.f;
#pragma unroll
                    for (int b = 0; b < kMaxBatches; ++b) {
                        const int sidx = b * G + gl;
                        gg[b] = (qvalid && sidx < LP) ? __ldcg(grad_attn + qm * LP + sidx) : 0.f;
                        dot += aw[b] * gg[b];
                    }
                    dot = group_allreduce_sum<G>(dot);
#pragma unroll
                    for (int b = 0; b < kMaxBatches; ++b) {
                        const int sidx = b * G + gl;
                        if (qvalid && sidx < LP) grad_attn[qm * LP + sidx] = aw[b] * (gg[b] - dot);
                    }
                }
            }
            __syncthreads();                                    // s_q is rewritten by the next round
        }
    }
}

template <typename VT, int D, bool FUSED>
int run_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
             const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
             cudaStream_t st)
{
    using C = TileCfg<D>;
    constexpr int MINB = 3;
    auto kern = bwd_tile_kernel<VT, D, FUSED, MINB>;
    static bool prepared[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !prepared[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) prepared[dev] = true;
    }
    kern<<<persistent_grid(MINB), kTileThreads, C::SMEM, st>>>((const VT *)value, shapes, lsi, (const float *)loc,
                                                                (const float *)attn, (const VT *)grad_out, (float *)gv,
                                                                (float *)gl, (float *)ga, d, (const float *)ref, ref_dim,
                                                                tuning().bwd_pipe > 0 ? tuning().bwd_pipe : 0);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename VT, bool FUSED>
int dispatch_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                  const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
                  cudaStream_t st)
{
    switch (d.D) {
    case 16: return run_tile<VT, 16, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 32: return run_tile<VT, 32, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 64: return run_tile<VT, 64, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    }
    return kUnsupported;
}

}  // namespace

#endif  // MSDA_AB

// bwd_variant 20 selects this kernel (any Lq) in -DMSDA_AB builds; see msda_backward.cu for the default choice
bool tiled_backward_applies(const Dims &d, DType dt, bool vec_ok)
{
#ifdef MSDA_AB
    if (!vec_ok || dt == DType::F64 || !(d.D == 16 || d.D == 32 || d.D == 64)) return false;
    if ((long)d.S * d.M * d.D >= (1L << 31) || (long)d.Lq * d.M * d.D >= (1L << 31) || d.L * d.P < 1) return false;
    return tuning().bwd_variant == 20;
#else
    (void)d; (void)dt; (void)vec_ok;
    return false;
#endif
}

// grad_value must already be zero-filled.  `ref` != nullptr selects the fused pre-processing flavour
// (loc = raw offsets, attn = raw logits, ref_dim = 2 or 6).
int launch_backward_tiled(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                          const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
                          const void *ref, int ref_dim, cudaStream_t st)
{
#ifdef MSDA_AB
    if (ref) {
        if (dt == DType::F32) return dispatch_tile<float, true>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
        return dispatch_tile<__nv_bfloat16, true>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    }
    if (dt == DType::F32) return dispatch_tile<float, false>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, nullptr, 2, st);
    return dispatch_tile<__nv_bfloat16, false>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, nullptr, 2, st);
#else
    (void)dt; (void)value; (void)shapes; (void)lsi; (void)loc; (void)attn; (void)grad_out; (void)gv; (void)gl; (void)ga;
    (void)d; (void)ref; (void)ref_dim; (void)st;
    return kUnsupported;
#endif
}

}  // namespace msda
