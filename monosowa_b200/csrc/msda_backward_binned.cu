// msda_backward_binned.cu -- MSDA backward for long query sets: the grad_value contributions of the
// COARSE feature levels are combined inside the SM before they leave it.
//
// Why: bwd_rec_kernel is bound by the reductions that leave the SM -- one 128-byte
// REDG line per (sample, corner), ~5.3 cycles each.  On a feature pyramid half of those lines go to the
// two coarsest levels, which are tiny: at the KITTI shape levels 2 and 3 hold 600 pixels per head but
// receive 8 of the 16 samples of every query, so a block of 256 queries sends ~7000 reduction lines to
// at most 600 distinct rows.  This kernel keeps those contributions on chip:
//
//   phase A  the record kernel's work (msda_backward.cu) for a chunk of QC consecutive queries of one
//            head: gathers, partial dots, grad_loc / grad_attn, and the REDGs of the FINE levels.  For a
//            sample on a coarse ("binned") level the owning lane instead drops a 16-byte entry
//            {a, lx, ly, cell | rank} into shared memory and counts it in a histogram over the base-corner
//            cells of the level -- a (H+1)x(W+1) lattice, the corner (y0, x0) ranges over [-1,H-1]x[-1,W-1].
//   phase B  exclusive scan of the histogram, then a counting-sort permutation of the entry indices.
//   phase C  one lane group per non-empty cell: all samples of a cell share their four corner pixels, so the
//            group reads each sample's grad_out row once (L1/L2 -- phase A just read it), accumulates the four corner
//            rows ((wy*wx)*a)*g in registers and walks runs of consecutive cells, carrying the corner column two
//            neighbouring cells share: two REDG lines per cell plus two per run.
//
// Which levels are binned is decided on the device from spatial_shapes (the host never reads them):
// the coarsest levels while their cells fit kMaxBins, their samples fit kBinSamples per query and the
// level has at most QC*P pixels (>= 4 updates per row on average).  If none qualifies the kernel is
// the record kernel plus one block barrier.  Arithmetic: same products as the direct path,
// ((wy*wx)*a)*g, accumulated per cell corner in fp32 registers before one global reduction per corner
// (the reference accumulates every contribution with a global atomic, ms_deform_im2col_cuda.cuh:125-152)
// -- covered by the backward tolerance, which already allows for atomic ordering.
#include "msda_common.cuh"
#include "msda_records.cuh"

namespace msda {

namespace {

constexpr int kBinSamples = 8;      // entry slots per (query, head): samples on binned levels
constexpr int kMaxBins = 704;       // base-corner cells over all binned levels
constexpr int kBinThreads = 256;
constexpr unsigned kNoEntry = 0xffffffffu;
constexpr int kBinMinQueries = 1024;  // below this the chunks are too few / too short to pay for the two extra phases

// CPL: channels per lane -- 4, or 8 for D = 64 (8 lanes per (query, head) like D = 32: see bwd_rec_kernel for the lane /
// reduction-chunk layout and why D = 32 stays at 4)
template <int D, int QCQ = 0, int CPL = kChannelsPerLane>
struct BinCfg {
    static constexpr int G = D / CPL;
    static constexpr int NCH = CPL / 4;                           // 16-byte reduction chunks per lane
    static constexpr int QPW = 32 / G;
    static constexpr int QPI = (kBinThreads / 32) * QPW;          // queries per pass of the CTA's warps
    static constexpr int QC = QCQ > 0 ? QCQ : (G <= 8 ? 256 : 128);    // queries per CTA
    static constexpr int HIST_HALVES = kMaxBins + 2;              // 16-bit counters, packed two per word
    static constexpr int HIST_WORDS = (HIST_HALVES + 1) / 2;
    static constexpr int G_BYTES = 0;                             // the chunk's grad_out rows are re-read from L1/L2
    static constexpr int ENT_BYTES = QC * kBinSamples * 16;
    static constexpr int HIST_BYTES = ((HIST_WORDS * 4 + 15) / 16) * 16;
    static constexpr int REC_BYTES = (kBinThreads / 32) * RecordLayout<G>::WARP_WORDS * 4;
    static constexpr int SMEM = G_BYTES + ENT_BYTES + HIST_BYTES + REC_BYTES;
    static_assert(REC_BYTES >= QC * kBinSamples * 2, "the sorted index array aliases the record area");
    static_assert(QC % QPI == 0, "chunk must be a whole number of passes");
    static_assert(QC * kBinSamples < 65536, "entry indices are 16-bit");
};

struct BinPlan {
    int lb;                          // first binned level (levels lb..L-1 are binned); L if none
    int lbP;                         // lb * P: first binned sample index of a query
    int nbs;                         // binned samples per query = (L - lb) * P
    int nbins;                       // cells over all binned levels
    int binbase[MSDA_MAX_LEVELS];    // first cell of each binned level
};

template <typename VT, int D, bool FUSED, int QCQ = 0, int MINB = 3, int CPL = kChannelsPerLane>
__global__ void __launch_bounds__(kBinThreads, MINB)
bwd_bin_kernel(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ loc, const float *__restrict__ attn, const VT *__restrict__ grad_out,
               float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn,
               const Dims d, const float *__restrict__ ref, const int ref_dim, const int max_binned)
{
    constexpr int LDQ = 1, LOADH = CPL == 4 ? 0 : 1, RUN = 4;
    using C = BinCfg<D, QCQ, CPL>;
    using RL = RecordLayout<C::G>;
    using V = VecN<VT, CPL>;
    constexpr int G = C::G, QPW = C::QPW, QPI = C::QPI, QC = C::QC, NCH = C::NCH;
    constexpr int LD = (LDQ <= G / 2) ? LDQ : G / 2;     // samples whose loads are issued together

    extern __shared__ __align__(16) unsigned char smem[];
    uint4 *s_ent = reinterpret_cast<uint4 *>(smem + C::G_BYTES);
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem + C::G_BYTES + C::ENT_BYTES);
    uint32_t *s_rec = reinterpret_cast<uint32_t *>(smem + C::G_BYTES + C::ENT_BYTES + C::HIST_BYTES);
    __shared__ LevelInfo s_lv[MSDA_MAX_LEVELS];
    __shared__ BinPlan s_plan;
    __shared__ int s_warp_tot[kBinThreads / 32];
    __shared__ int s_next;

    const int tid = threadIdx.x;
    for (int i = tid; i < C::HIST_WORDS; i += kBinThreads) s_hist[i] = 0u;
    stage_levels(s_lv, shapes, lsi, d.L);
    if (tid == 0) {
        BinPlan p;
        p.lb = d.L;
        int nbins = 0;
        for (int l = d.L - 1; l >= 0; --l) {
            const int H = s_lv[l].H, W = s_lv[l].W;
            if (H < 0 || W < 0 || H > 8192 || W > 8192) break;
            const int cells = (H + 1) * (W + 1);
            if ((d.L - l) * d.P > max_binned || nbins + cells > kMaxBins || H * W > QC * d.P) break;
            p.binbase[l] = nbins;
            nbins += cells;
            p.lb = l;
        }
        p.lbP = p.lb * d.P;
        p.nbs = (d.L - p.lb) * d.P;
        p.nbins = nbins;
        s_plan = p;
        s_next = 0;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int gl = lane % G, k = lane / G;
    const int n_chunks = (d.Lq + QC - 1) / QC;
    const int m = (int)(blockIdx.x % d.M);
    const long cr = blockIdx.x / d.M;
    const int q0 = (int)(cr % n_chunks) * QC;
    const int n = (int)(cr / n_chunks);

    const int LP = d.L * d.P;
    const int lbP = s_plan.lbP;
    const long img = ((long)n * d.S * d.M + m) * D;
    const VT *vimg = value + img + gl * CPL;             // gathers: CPL contiguous channels
    float *gvimg = grad_value + img + gl * 4;            // reductions: chunk j = channels [j*4G + 4*gl, +4)
    const int xs = d.M * D;
    uint32_t *grp = s_rec + warp * RL::WARP_WORDS + k * RL::GROUP_WORDS;

    // ---- phase A --------------------------------------------------------------------------------
    for (int it = 0; it < QC / QPI; ++it) {
        const int qw = it * QPI + warp * QPW;            // first local query of this warp's pass
        if (q0 + qw >= d.Lq) break;                      // whole warp past the end
        const int ql = qw + k;
        const bool qvalid = q0 + ql < d.Lq;
        const long qm = ((long)n * d.Lq + (qvalid ? q0 + ql : q0)) * d.M + m;

        float g[CPL], gr[NCH][4];
        V::load(grad_out + qm * D + gl * CPL, g);                              // read again by phase C: keep it cached
        if constexpr (NCH == 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) gr[0][c] = g[c];
        } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) Vec4<VT>::load(grad_out + qm * D + j * 4 * G + gl * 4, gr[j]);
        }
        auto reduce_row = [&](int off, float wgt) {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                red_add_f32x4(gvimg + off + j * 4 * G, wgt * gr[j][0], wgt * gr[j][1], wgt * gr[j][2], wgt * gr[j][3]);
        };

        float aw[kMaxBatches];
        float pa[kMaxBatches], pg[kMaxBatches];
        if constexpr (FUSED) {
            group_softmax<G>(attn, qm * LP, LP, gl, qvalid, aw);
#pragma unroll
            for (int b = 0; b < kMaxBatches; ++b) pa[b] = pg[b] = 0.f;
        }
        auto fetch = [&](int sidx) -> SampleIn {
            const bool has = qvalid && sidx < LP;
            if constexpr (FUSED) {
                const int l = has ? sidx / d.P : 0;
                const SampleIn r = fetch_sample_fused<stream_policy<G, VT>()>(has, loc, ref, ref_dim, qm * LP + sidx, (qm / d.M) * d.L + l, s_lv, l,
                                                      d.P, aw[0]);
                aw[0] = aw[1]; aw[1] = aw[2]; aw[2] = aw[3];
                return r;
            } else {
                return fetch_sample<stream_policy<G, VT>()>(has, loc, attn, qm * LP + sidx);
            }
        };
        SampleIn in = fetch(gl);
        for (int b0 = 0; b0 < LP; b0 += G) {
            const int sidx = b0 + gl;
            const bool has = qvalid && sidx < LP;
            const int l = sidx / d.P;
            int4 roff;
            float4 rwa;
            const SampleGeom gm = sample_geometry(has && d.S > 0, in, s_lv, l, xs, roff, rwa);
            if (!gm.live) roff.x = -1;                           // consumers skip the gathers of this sample
            *reinterpret_cast<int4 *>(grp + gl * 4) = roff;
            *reinterpret_cast<float4 *>(grp + RL::WEIGHTS + gl * 4) = rwa;
            const float a_cur = in.a, ex_cur = in.ex, ey_cur = in.ey;
            if (has && sidx >= lbP) {
                // binned level: park the sample instead of sending its four reduction lines to L2
                uint4 e = make_uint4(0u, 0u, 0u, kNoEntry);
                if (gm.live && gm.a != 0.f) {
                    const int bin = s_plan.binbase[l] + gm.cy * (s_lv[l].W + 1) + gm.cx;
                    const unsigned sh = (bin & 1) * 16;
                    const unsigned old = atomicAdd(&s_hist[bin >> 1], 1u << sh);
                    e = make_uint4(__float_as_uint(gm.a), __float_as_uint(gm.lx), __float_as_uint(gm.ly),
                                   (unsigned)bin | (((old >> sh) & 0xffffu) << 16));
                }
                s_ent[ql * kBinSamples + (sidx - lbP)] = e;
            }
            __syncwarp();
            in = fetch(sidx + G);

            constexpr int GH = G / 2;
            float th[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float t[4 * GH];
                // LD samples are loaded together (4*LD independent LDG.128 in flight per lane) before any of them is
                // consumed.  The ncu source view shows 44 % of the stall samples as long-scoreboard waits on the first
                // FFMA of each sample, yet LD = 1 is the measured optimum: LD = 2 / 4 are 25-90 % SLOWER with or without
                // L1 allocation (profiles/r01b_sweep_load_grouping.jsonl) -- the latency is queueing in a saturated
                // memory system, not exposed idle time, and deeper bursts only lengthen the queues.
#pragma unroll
                for (int u0 = 0; u0 < GH; u0 += LD) {
                    int4 off[LD];
                    float4 wa[LD];
                    float v[LD][4][CPL];
#pragma unroll
                    for (int j = 0; j < LD; ++j) {
                        const int s = h * GH + u0 + j;
                        off[j] = *reinterpret_cast<const int4 *>(grp + s * 4);
                        wa[j] = *reinterpret_cast<const float4 *>(grp + RL::WEIGHTS + s * 4);
#pragma unroll
                        for (int c = 0; c < CPL; ++c) v[j][0][c] = v[j][1][c] = v[j][2][c] = v[j][3][c] = 0.f;
                        // outside the window / past L*P: nothing is read (the reference's branch, cuh:288, 365-367)
                        if (off[j].x >= 0) {
                            V::template gather<LOADH>(vimg + off[j].x, v[j][0]);
                            V::template gather<LOADH>(vimg + off[j].y, v[j][1]);
                            V::template gather<LOADH>(vimg + off[j].z, v[j][2]);
                            V::template gather<LOADH>(vimg + off[j].w, v[j][3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < LD; ++j) {
                        const int u = u0 + j, s = h * GH + u;
                        t[4 * u] = t[4 * u + 1] = t[4 * u + 2] = t[4 * u + 3] = 0.f;
                        if (d.S > 0) {
#pragma unroll
                            for (int c = 0; c < CPL; ++c) {
                                t[4 * u] += g[c] * v[j][0][c];
                                t[4 * u + 1] += g[c] * v[j][1][c];
                                t[4 * u + 2] += g[c] * v[j][2][c];
                                t[4 * u + 3] += g[c] * v[j][3][c];
                            }
                        }
                        if (b0 + s < lbP) {                  // fine level (warp-uniform): direct reductions
                            if (wa[j].x != 0.f) reduce_row(off[j].x, wa[j].x);
                            if (wa[j].y != 0.f) reduce_row(off[j].y, wa[j].y);
                            if (wa[j].z != 0.f) reduce_row(off[j].z, wa[j].z);
                            if (wa[j].w != 0.f) reduce_row(off[j].w, wa[j].w);
                        }
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
                const float b0_ = __shfl_sync(kFullMask, th[1][0], src), b1 = __shfl_sync(kFullMask, th[1][1], src);
                const float b2 = __shfl_sync(kFullMask, th[1][0], src + 1), b3 = __shfl_sync(kFullMask, th[1][1], src + 1);
                t[0] = second ? b0_ : a0; t[1] = second ? b1 : a1; t[2] = second ? b2 : a2; t[3] = second ? b3 : a3;
            }
            if (has) {
                float gx = 0.f, gy = 0.f, ga = 0.f;
                if (gm.live) {
                    const float t00 = (gm.vmask & 1u) ? t[0] : 0.f, t01 = (gm.vmask & 2u) ? t[1] : 0.f;
                    const float t10 = (gm.vmask & 4u) ? t[2] : 0.f, t11 = (gm.vmask & 8u) ? t[3] : 0.f;
                    ga = gm.w00 * t00 + gm.w01 * t01 + gm.w10 * t10 + gm.w11 * t11;
                    gx = gm.Wf * gm.a * (gm.hy * (t01 - t00) + gm.ly * (t11 - t10));
                    gy = gm.Hf * gm.a * (gm.hx * (t10 - t00) + gm.lx * (t11 - t01));
                }
                const long si = qm * LP + sidx;
                if constexpr (FUSED) {
                    const float2 go = gm.live ? fused_offset_grad(ref_dim, gx, gy, gm.Wf, gm.Hf, ex_cur, ey_cur, d.P)
                                              : make_float2(0.f, 0.f);
                    __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), go);
                    pa[0] = pa[1]; pa[1] = pa[2]; pa[2] = pa[3]; pa[3] = a_cur;
                    pg[0] = pg[1]; pg[1] = pg[2]; pg[2] = pg[3]; pg[3] = ga;
                } else {
                    __stcs(reinterpret_cast<float2 *>(grad_loc + 2 * si), make_float2(gx, gy));
                    __stcs(grad_attn + si, ga);
                }
            } else if constexpr (FUSED) {
                pa[0] = pa[1]; pa[1] = pa[2]; pa[2] = pa[3]; pa[3] = 0.f;
                pg[0] = pg[1]; pg[1] = pg[2]; pg[2] = pg[3]; pg[3] = 0.f;
            }
        }
        if constexpr (FUSED) {
            float dot = 0.f;
#pragma unroll
            for (int b = 0; b < kMaxBatches; ++b) dot += pa[b] * pg[b];
            dot = group_allreduce_sum<G>(dot);
            const int nb = (LP + G - 1) / G;
#pragma unroll
            for (int b = 0; b < kMaxBatches; ++b) {
                const int sidx = (b - (kMaxBatches - nb)) * G + gl;
                if (b >= kMaxBatches - nb && qvalid && sidx < LP) grad_attn[qm * LP + sidx] = pa[b] * (pg[b] - dot);
            }
        }
    }

    const int nbs = s_plan.nbs;
    if (nbs == 0) return;                                // block-uniform: nothing was binned
    __syncthreads();

    // ---- phase B: histogram -> exclusive offsets (in place), then the counting-sort permutation --------
    unsigned short *hh = reinterpret_cast<unsigned short *>(s_hist);
    {
        constexpr int BPT = (C::HIST_HALVES + kBinThreads - 1) / kBinThreads;
        int c[BPT], sum = 0;
#pragma unroll
        for (int i = 0; i < BPT; ++i) {
            const int idx = tid * BPT + i;
            c[i] = idx < C::HIST_HALVES ? hh[idx] : 0;
            sum += c[i];
        }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(kFullMask, incl, off);
            if (lane >= off) incl += v;
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
    unsigned short *s_idx = reinterpret_cast<unsigned short *>(s_rec);    // phase A is over: reuse the record area
    {
        const int nq = min(QC, d.Lq - q0);
        for (int e = tid; e < nq * kBinSamples; e += kBinThreads) {
            if ((e & (kBinSamples - 1)) >= nbs) continue;
            const unsigned pk = s_ent[e].w;
            if (pk != kNoEntry) s_idx[hh[pk & 0xffffu] + (pk >> 16)] = (unsigned short)e;
        }
    }
    __syncthreads();

    // ---- phase C: one lane group per run of RUN consecutive base-corner cells -------------------------
    // Every sample of a cell shares its four corner pixels: the group reads each sample's grad_out row ONCE (a
    // pixel-owner formulation reads it four times -- measured: more L1 data-pipe wavefronts than the reductions it
    // replaces) and forms the four corner rows in registers.  Neighbouring cells of a lattice row share a corner
    // column (the right corners of cell bx are the left corners of cell bx + 1), so a group that walks RUN cells of a
    // row carries the right accumulators over and sends two REDG lines per cell plus two per run instead of four
    // per cell: 35 % fewer flush lines at RUN = 4 (bwd_variant 25 = RUN 1 for A/B; the run time is unchanged within
    // noise, profiles/r01b_sweep_cell_runs.jsonl -- the flush is not what paces the kernel).
    const unsigned gmask = (G == 32) ? kFullMask : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const int nbins = s_plan.nbins, lb = s_plan.lb;
    const VT *gq_global = grad_out + (((long)n * d.Lq + q0) * d.M + m) * D + gl * 4;
    auto flush = [&](float *p, const float (&acc)[NCH][4]) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) red_add_f32x4(p + j * 4 * G, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    };
    for (;;) {
        int c0 = 0;
        if (gl == 0) c0 = atomicAdd(&s_next, RUN);
        c0 = __shfl_sync(gmask, c0, 0, G);
        if (c0 >= nbins) break;
        const int c1 = min(c0 + RUN, nbins);
        float l0[NCH][4], l1[NCH][4];                                          // corners (y0, x0), (y1, x0)
        float r0[NCH][4], r1[NCH][4];                                          // corners (y0, x1), (y1, x1)
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) l0[j][k2] = l1[j][k2] = r0[j][k2] = r1[j][k2] = 0.f;
        bool lt = false, rt = false;                                           // accumulators hold something
        for (int c = c0; c < c1; ++c) {
            int l = lb;
            while (l < d.L - 1 && c < s_plan.binbase[l]) ++l;        // binbase grows towards the finer levels
            const LevelInfo li = s_lv[l];
            const int cc = c - s_plan.binbase[l];
            const int by = cc / (li.W + 1), bx = cc - by * (li.W + 1);
            const int i0 = hh[c], i1 = hh[c + 1];
#pragma unroll 2
            for (int i = i0; i < i1; ++i) {
                const int e = s_idx[i];
                const uint4 en = s_ent[e];
                const float a = __uint_as_float(en.x), lx = __uint_as_float(en.y), ly = __uint_as_float(en.z);
                const float hy = 1.f - ly, hx = 1.f - lx;
                const float w00 = (hy * hx) * a, w01 = (hy * lx) * a, w10 = (ly * hx) * a, w11 = (ly * lx) * a;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    float gt[4];
                    Vec4<VT>::load(gq_global + (long)(e / kBinSamples) * xs + j * 4 * G, gt);
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        l0[j][k2] += w00 * gt[k2];
                        r0[j][k2] += w01 * gt[k2];
                        l1[j][k2] += w10 * gt[k2];
                        r1[j][k2] += w11 * gt[k2];
                    }
                }
            }
            if (i1 > i0) lt = rt = true;
            // corners outside the map receive nothing (zero padding, ms_deform_im2col_cuda.cuh:125-152)
            const int y0 = by - 1, x0 = bx - 1;
            const bool y0ok = y0 >= 0, y1ok = by <= li.H - 1;
            float *row = gvimg + (li.start + y0 * li.W + x0) * xs;             // pixel (y0, x0)
            const int ys = li.W * xs;
            if (lt && x0 >= 0) {                                               // the left column is complete now
                if (y0ok) flush(row, l0);
                if (y1ok) flush(row + ys, l1);
            }
            if (c + 1 < c1 && bx < li.W) {                                     // next cell continues this lattice row
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) { l0[j][k2] = r0[j][k2]; l1[j][k2] = r1[j][k2]; r0[j][k2] = 0.f; r1[j][k2] = 0.f; }
                lt = rt;
                rt = false;
            } else {
                if (rt && bx <= li.W - 1) {
                    if (y0ok) flush(row + xs, r0);
                    if (y1ok) flush(row + ys + xs, r1);
                }
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) l0[j][k2] = l1[j][k2] = r0[j][k2] = r1[j][k2] = 0.f;
                lt = rt = false;
            }
        }
    }
}

template <typename VT, int D, bool FUSED, int CPL, int MINB>
int launch_bin(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
               const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim, cudaStream_t st)
{
    using C = BinCfg<D, 0, CPL>;
    auto kern = bwd_bin_kernel<VT, D, FUSED, 0, MINB, CPL>;
    static bool prepared[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64 || !prepared[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) prepared[dev] = true;
    }
    const long n_chunks = (d.Lq + C::QC - 1) / C::QC;
    const long grid = (long)d.N * d.M * n_chunks;
    if (grid > 0x7fffffffL) return kUnsupported;
    kern<<<(unsigned)grid, kBinThreads, C::SMEM, st>>>((const VT *)value, shapes, lsi, (const float *)loc,
                                                        (const float *)attn, (const VT *)grad_out, (float *)gv,
                                                        (float *)gl, (float *)ga, d, (const float *)ref, ref_dim, kBinSamples);
    count_launch();
    return (int)cudaGetLastError();
}

// D = 64 takes 8 channels per lane (8-lane groups, 256-query chunks like D = 32) when `value` is aligned to 8 elements
// and, fused, the L*P logits fit kMaxBatches batches of 8 lanes
template <typename VT, int D, bool FUSED>
bool bin_wide(const void *value, const Dims &d)
{
    return D == 64 && tuning().bwd_pipe != 4 && reinterpret_cast<uintptr_t>(value) % (8 * sizeof(VT)) == 0 &&
           (!FUSED || d.L * d.P <= kMaxBatches * (D / 8));
}

template <typename VT, int D, bool FUSED>
int run_bin(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
            const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim, cudaStream_t st)
{
    // chunk size / register cap / where the chunk's grad_out rows live were swept on B200
    // (profiles/r01b_sweep_binned_flavours*.jsonl): 256 queries at 80 registers (3 CTAs/SM) with grad_out re-read
    // through L1/L2 in phase C wins -- parking the rows in shared memory (+32 KB per CTA) shrinks L1 to ~28 KB and
    // costs 3 %; 64 registers (4 CTAs/SM), 128 registers (2 CTAs/SM) and 128/160/320/512-query chunks are 1-10 % slower
    if constexpr (D == 64) {
        if (bin_wide<VT, D, FUSED>(value, d)) {
            // profiles/r02_bwd_bin_d64_interleaved.jsonl (configs[1] shape with 4 heads of 64, timed alternately): fp32
            // 1.47 ms at 2 CTAs/SM (128 registers, no spills) against 1.57 at 3 CTAs/SM (80 registers, 96 B spilled) and
            // 1.58 for the 4-channel flavour; bf16 1.48 (3 CTAs/SM) against 1.65
#ifdef MSDA_AB
            if (tuning().bwd_pipe == 83)
                return launch_bin<VT, D, FUSED, 8, 3>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
#endif
            return launch_bin<VT, D, FUSED, 8, 2>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
        }
    }
#ifdef MSDA_AB
    if constexpr (D == 32)                       // A/B: 2 CTAs/SM, 128 registers (the fused flavour spills ~100 B at 80)
        if (tuning().bwd_pipe == 92)
            return launch_bin<VT, D, FUSED, 4, 2>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
#endif
    return launch_bin<VT, D, FUSED, 4, 3>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
}

template <typename VT, bool FUSED>
int dispatch_bin(const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc, const void *attn,
                 const void *grad_out, void *gv, void *gl, void *ga, const Dims &d, const void *ref, int ref_dim,
                 cudaStream_t st)
{
    switch (d.D) {
    case 16: return run_bin<VT, 16, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 32: return run_bin<VT, 32, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    case 64: return run_bin<VT, 64, FUSED>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    }
    return kUnsupported;
}

}  // namespace

// bwd_variant: -1 default (binned kernel for long query sets); 21 force it for any Lq; 11/20/99 never.
bool binned_backward_applies(const Dims &d, DType dt, bool vec_ok)
{
    const int v = tuning().bwd_variant;
    if (!vec_ok || dt == DType::F64 || !(d.D == 16 || d.D == 32 || d.D == 64)) return false;
    if ((long)d.S * d.M * d.D >= (1L << 31) || d.L * d.P < 1) return false;
    if (v == 21) return true;
    // Three barrier-separated phases per CTA need several waves of CTAs to overlap each other: measured on B200
    // (profiles/r02_binned_crossover.jsonl, r02_kernel_family_table.jsonl) the binned kernel wins from ~2560 work items
    // (KITTI batch 8: 0.75 vs 0.78 ms; batch 16: 1.43 vs 1.54; Waymo batch 4: 2.01 vs 2.09), is level around 1300-2000
    // and loses below (KITTI batch 2: 0.26 vs 0.21 ms; KITTI-360 / 640x960 batch 4: 0.47 vs 0.43 ms).
    const long chunk = 256;                          // queries per CTA (D = 64 without 32-byte aligned values: 128)
    return v == -1 && d.Lq >= kBinMinQueries && (long)d.N * d.M * ((d.Lq + chunk - 1) / chunk) >= 2560;
}

// grad_value (fp32) must already be zero-filled.  `ref` != nullptr selects the fused pre-processing flavour
// (loc = raw offsets, attn = raw logits, ref_dim = 2 or 6).
int launch_backward_binned(DType dt, const void *value, const int64_t *shapes, const int64_t *lsi, const void *loc,
                           const void *attn, const void *grad_out, void *gv, void *gl, void *ga, const Dims &d,
                           const void *ref, int ref_dim, cudaStream_t st)
{
    if (ref) {
        if (dt == DType::F32) return dispatch_bin<float, true>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
        return dispatch_bin<__nv_bfloat16, true>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, ref, ref_dim, st);
    }
    if (dt == DType::F32) return dispatch_bin<float, false>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, nullptr, 2, st);
    return dispatch_bin<__nv_bfloat16, false>(value, shapes, lsi, loc, attn, grad_out, gv, gl, ga, d, nullptr, 2, st);
}

}  // namespace msda
