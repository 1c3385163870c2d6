// sb_pipeline.cu -- fused per-tile stain pipeline for sm_100a.
//
// One thread-block cluster (1, 2, 4 or 8 CTAs) owns one tile and runs the whole dependency chain of
//   ExtractiveStainNormalizer.transform  (stainlib/normalization/normalizer.py:39-50)
// on it without leaving the kernel; clusters are persistent and stride over the batch:
//
//   Macenko (macenko_stain_extractor.py:7-44)
//     A    tissue mask + masked OD moments (n, sum od, sum od x od)     -> 3x3 covariance, fp64 Jacobi eigenvectors
//     B0   angle keys of a 1-in-16 sample -> 4096-bin histogram          -> a key bracket around each angular percentile
//     B1'  one full pass: count the keys below each bracket, list the keys inside -> exact order statistics -> stain matrix
//          (on a miss, and for clusters: B1 full 4096-bin histogram + B2 11-bit refinement -- same keys, same result)
//   Vahadane (vahadane_stain_extractor.py:19-43; spams.trainDL restated as a deterministic dictionary iteration)
//     V0   tissue mask (kept as one bit per pixel)
//     V    sparse-code the tissue pixels, accumulate A = sum aa^T, B = sum xa^T, update D: sample passes to a residual of
//          1e-4, then full passes to 2e-6, the fixed-point map Anderson-accelerated (memory 4)
//   common (stain_utils.py:69-78, normalizer.py:46-50)
//     C0 + C1'  the same sampled-bracket selection for the 99th percentile of each closed-form non-negative LASSO
//          concentration over ALL pixels (fallback C1 + C2: two-level histograms)
//     D    recombine with the target matrix: K4 (sb_recombine.cu on the ring of sb_ring.cuh), launched right behind this
//          kernel on the same stream with the per-tile statistics computed here (sb_normalize = this kernel + K4)
//
// Every pass re-reads the tile with 16-byte vector loads.  Instruction issue, not HBM, bounds this kernel, so the
// passes are built to minimise instructions: the lookup table holds one {od, gamma} pair per lane in 256-byte rows (one
// PRMT makes the offset, one conflict-free LDS.64 returns density and linearised value; the Macenko passes recompute
// the tissue mask from the gammas instead of storing it), the moments are predicated adds, the concentration pass runs
// the compare-free LASSO on the packed f32x2 pipe, the rare pixels that need an exact key go through per-warp
// compaction queues, and the ragged last group is peeled off so the main loops carry no validity checks.  Cross-CTA
// reductions (moments, histograms) go through distributed shared memory; every CTA of a cluster redundantly evaluates
// the small serial steps (eigenvectors, selections) so no broadcast is needed and results are bit-identical.
// The slide-level (multi-tile, multi-rank) fit passes at the end of the file produce the same statistics as sums that
// add across tiles, launches and ranks.
#include "sb_kernels.h"
#include "sb_pipe_common.cuh"

namespace sb {

template <int METHOD>
__global__ void __launch_bounds__(NT, 2) tile_pipeline_kernel(PipeArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* od_rep = smem_raw;                                        // 64 KB: OD table + mask bits
    PipeShared* sh = reinterpret_cast<PipeShared*>(smem_raw + OD_REP_BYTES);
    const int S = a.cluster_size;
    const int crank = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int cluster_id = blockIdx.x / S, n_clusters = gridDim.x / S;
    const int npx = a.npx;
    const int G = (npx + GROUP_PX - 1) / GROUP_PX;
    // a cluster splits the tile at unit boundaries (see for_each_unit)
    const int UK = unit_groups(G), U = (G + UK * NT - 1) / (UK * NT);
    const int ub = (U * crank) / S, ue = (U * (crank + 1)) / S;
    const int gb = min(ub * UK * NT, G), ge = min(ue * UK * NT, G);
    const size_t tile_bytes = (size_t)npx * 3;
    const bool aligned = a.aligned != 0;
    const uint32_t lane_off = (threadIdx.x & 31) << 3;      // {od, gamma} pairs: 8 bytes per lane
    const YCoef yc{a.ycoef[0], a.ycoef[1], a.ycoef[2], a.ybound};
    // Vahadane keeps the tissue mask of its tile for all dictionary passes: 16 bits per group in the idle histogram
    // buffer when the CTA's share fits (262,144 pixels), else in a global scratch row (L2-resident, one 2-byte load per group)
    const bool cache_smem = (ge - gb) <= MASK_CAP_GROUPS;
    const bool cache_mask = cache_smem || a.mask_scratch != nullptr;

    const int n_todo = a.tile_count ? min(*a.tile_count, a.B) : a.B;
    if (cluster_id >= n_todo) return;                      // (whole clusters leave together; the fallback launch of the streaming path is usually empty)
    fill_odg_rep(od_rep, a.tab.od, a.tab.gamma, NT);
    if (threadIdx.x < 10) sh->acc64[threadIdx.x] = 0ull;
    __syncthreads();
    int pbuf = 0;   // parity of sh->part

    for (int t_i = cluster_id; t_i < n_todo; t_i += n_clusters) {
        const int tile = a.tile_list ? a.tile_list[t_i] : t_i;
        const uint8_t* __restrict__ tin = a.in + (size_t)tile * tile_bytes;
        unsigned n_tissue = 0;
        if (threadIdx.x == 0) { sh->flags = 0; }

        if (METHOD == SB_METHOD_MACENKO) {
            // ------------------------------------------------------------------ A: mask + moments
            long long acc[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] = 0;
            unsigned cnt = 0;
            for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int) {
                constexpr bool TAIL = decltype(tail)::value;
                float f[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) f[i] = 0.f;
                for_each_px_odg(od_rep, lane_off, w, [&](int i, float2 r, float2 g, float2 b) {
                    float y = tissue_y(yc, r.y, g.y, b.y);
                    if (TAIL && i >= nvalid) y = yc.bound;
                    accum_if_tissue(y, yc.bound, r.x, g.x, b.x, f, cnt);
                });
                // the fp32 sums of ONE 16-pixel group enter the fixed-point accumulators: independent of the thread layout
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += to_fix(f[i], FIX_MOMENT);
            });
            block_reduce10(sh, pbuf, acc, cnt);
            tile_sync(S);
            cluster_total10(sh, pbuf, S, 1.0 / (double)FIX_MOMENT);
            pbuf ^= 1;
            n_tissue = (unsigned)sh->tot[9];
            if (threadIdx.x == 0) {
                const double n = sh->tot[9];
                int flags = 0;
                if (n < 1.0) flags |= SB_STATUS_EMPTY_MASK;
                else if (n < 2.0) flags |= SB_STATUS_FEW_TISSUE;
                if (!flags) {
                    const double* t = sh->tot;
                    const double inv = 1.0 / (n - 1.0);
                    double c[6];
                    c[0] = (t[3] - t[0] * t[0] / n) * inv; c[1] = (t[4] - t[0] * t[1] / n) * inv; c[2] = (t[5] - t[0] * t[2] / n) * inv;
                    c[3] = (t[6] - t[1] * t[1] / n) * inv; c[4] = (t[7] - t[1] * t[2] / n) * inv; c[5] = (t[8] - t[2] * t[2] / n) * inv;
                    double wv[3], v[3][3];
                    jacobi_eig3(c, wv, v);
                    int i1 = 0;
                    if (wv[1] > wv[i1]) i1 = 1;
                    if (wv[2] > wv[i1]) i1 = 2;
                    int i2 = -1;
                    for (int k = 0; k < 3; ++k) if (k != i1 && (i2 < 0 || wv[k] > wv[i2])) i2 = k;
                    const double s1 = v[0][i1] < 0 ? -1.0 : 1.0, s2 = v[0][i2] < 0 ? -1.0 : 1.0;
                    bool ok = true;
                    for (int k = 0; k < 3; ++k) {
                        const double x1 = s1 * v[k][i1], x2 = s2 * v[k][i2];
                        sh->V[k] = (float)x1; sh->V[3 + k] = (float)x2;
                        sh->D[k] = x1; sh->D[3 + k] = x2;        // keep the fp64 eigenvectors
                        ok = ok && isfinite(x1) && isfinite(x2);
                    }
                    if (!ok) flags |= SB_STATUS_DEGENERATE;
                }
                sh->flags = flags;
            }
            __syncthreads();
            if (sh->flags == 0) {
                const float v00 = sh->V[0], v01 = sh->V[1], v02 = sh->V[2], v10 = sh->V[3], v11 = sh->V[4], v12 = sh->V[5];
                unsigned p_lo[2], p_hi[2];
                { double fr; percentile_index(n_tissue, 100.0 - a.ang_pct, p_lo[0], p_hi[0], fr); percentile_index(n_tissue, a.ang_pct, p_lo[1], p_hi[1], fr); }
                bool sampled = false;
                if (S == 1 && n_tissue >= 16384u) {
                    // ---------------------------------------------------------- B0: angle keys of a 1-in-16 sample
                    if (threadIdx.x == 0) { sh->s_cnt = 0; sh->l_len[0] = sh->l_len[1] = 0; sh->l_below[0] = sh->l_below[1] = 0; sh->s_ok = 0; sh->wq_overflow = 0; }
                    zero_hist(sh);
                    unsigned scnt = 0;
                    for_each_sample_group(tin, npx, gb, ge, aligned, [&](auto, const uint32_t (&w)[12], int, int) {
                        for_each_px_odg(od_rep, lane_off, w, [&](int, float2 r, float2 g, float2 b) {
                            const float px = fmaf(b.x, v02, fmaf(g.x, v01, r.x * v00));
                            const float py = fmaf(b.x, v12, fmaf(g.x, v11, r.x * v10));
                            if (tissue_y(yc, r.y, g.y, b.y) < yc.bound) { ++scnt; atomicAdd(&sh->hist[angle_key(px, py) >> L2_BITS], 1u); }
                        });
                    });
                    scnt = warp_sum_u(scnt);
                    if ((threadIdx.x & 31) == 0 && scnt) atomicAdd(&sh->s_cnt, scnt);
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const unsigned n_s = sh->s_cnt;
                        if (n_s >= 1024u) {
                            plan_bracket(n_tissue, n_s, p_lo[0], sh->q_rank[0], sh->q_rank[1], a.bracket_sigmas, a.bracket_pad);
                            plan_bracket(n_tissue, n_s, p_lo[1], sh->q_rank[2], sh->q_rank[3], a.bracket_sigmas, a.bracket_pad);
                            for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
                            sh->s_ok = 1;
                        }
                    }
                    __syncthreads();
                    if (sh->s_ok) {
                        select_ranks<L1_BINS>(sh, sh->hist, 1, sh->q_rank, 4, sh->q_bin, sh->q_rem);
                        if (threadIdx.x == 0) {
                            for (int j = 0; j < 2; ++j) {
                                // (a rank clamped to the sample's first / last element brackets at that element's bin: keys outside
                                //  are still counted exactly, and the validation below catches a target rank that fell outside)
                                sh->brk_a[j] = sh->q_bin[2 * j] << L2_BITS;
                                sh->brk_b[j] = (sh->q_bin[2 * j + 1] + 1u) << L2_BITS;
                            }
                            // A pixel in the half-plane x > 0 whose diamond coordinate d = y/(x+|y|) lies safely between the
                            // low bracket and the high bracket (64 key units of slack, the key arithmetic errs by < 1) is
                            // "above bracket 0, below bracket 1" without computing its key.
                            const double d_lo = diamond_from_key((double)sh->brk_b[0] + 64.0), d_hi = diamond_from_key((double)sh->brk_a[1] - 64.0);
                            const bool usable = sh->brk_a[1] >= 64u && d_lo > -0.999 && d_hi < 0.999 && d_lo < d_hi;
                            sh->fast_lo[0] = usable ? float_above(d_lo) : INFINITY;
                            sh->fast_hi[0] = usable ? float_below(d_hi) : -INFINITY;
                        }
                        __syncthreads();
                        // ------------------------------------------------------ B1': count below / collect inside the brackets
                        const unsigned ka0 = sh->brk_a[0], kb0 = sh->brk_b[0], ka1 = sh->brk_a[1], kb1 = sh->brk_b[1];
                        const float f_lo = sh->fast_lo[0], f_hi = sh->fast_hi[0];
                        const KeyList list0{sh->hist, ka0, kb0 - ka0 > LIST_SPAN}, list1{sh->hist + L1_BINS, ka1, kb1 - ka1 > LIST_SPAN};
                        unsigned below0 = 0, below1 = 0;
                        // exact treatment of one pixel (packed RGB): key, below counters, list appends
                        auto exact_px = [&](bool has, uint32_t rgb) {
                            const float o0 = od_lookup(od_rep, rgb, lane_off, 0), o1 = od_lookup(od_rep, rgb, lane_off, 1), o2 = od_lookup(od_rep, rgb, lane_off, 2);
                            const float px = fmaf(o2, v02, fmaf(o1, v01, o0 * v00));
                            const float py = fmaf(o2, v12, fmaf(o1, v11, o0 * v10));
                            const uint32_t key = angle_key(px, py);
                            if (has) {
                                if (key < ka0) ++below0;
                                else if (key < kb0) { const unsigned idx = atomicAdd(&sh->l_len[0], 1u); if (idx < list0.cap()) list0.put(idx, key); }
                                if (key < ka1) ++below1;
                                else if (key < kb1) { const unsigned idx = atomicAdd(&sh->l_len[1], 1u); if (idx < list1.cap()) list1.put(idx, key); }
                            }
                        };
                        WarpQueue wq{sh->wq[threadIdx.x >> 5], 0u};
                        auto drain_px = [&](bool has, uint32_t pos) { exact_px(has, load_px(tin, pos)); };
                        const float nf_lo = -f_lo;
                        for_each_group_uniform(tin, npx, gb, ge, aligned, [&](const uint32_t (&w)[12], bool active, int g) {
                            // inactive lanes hold an all-white group: no tissue bit is set for it
                            uint32_t mbits = 0, fastbits = 0;
                            for_each_px_odg(od_rep, lane_off, w, [&](int i, float2 r, float2 gg, float2 b) {
                                const float px = fmaf(b.x, v02, fmaf(gg.x, v01, r.x * v00));
                                const float py = fmaf(b.x, v12, fmaf(gg.x, v11, r.x * v10));
                                const float sd = px + fabsf(py);
                                // inside (f_lo, f_hi) in the half-plane x > 0  <=>  min(py - f_lo sd, f_hi sd - py, px) > 0
                                const float mn = min3f(fmaf(nf_lo, sd, py), fmaf(f_hi, sd, -py), px);
                                mbits |= set_lt(tissue_y(yc, r.y, gg.y, b.y), yc.bound) & (1u << i);
                                fastbits |= set_gt(mn, 0.f) & (1u << i);
                            });
                            if (!active) mbits = 0;
                            below1 += __popc(mbits & fastbits);
                            wq_push_flagged(wq, mbits & ~fastbits, (unsigned)g * GROUP_PX, &sh->wq_overflow);
                            wq_drain(wq, false, drain_px);
                        });
                        wq_drain(wq, true, drain_px);
                        if ((npx % GROUP_PX) != 0 && ge == G && threadIdx.x == 0) {
                            // ragged last group: plain exact path
                            uint32_t w[12];
                            int nvalid;
                            load_group<true>(tin, npx, G - 1, false, w, nvalid);
                            const uint32_t mbits = mask16<true>(od_rep, lane_off, w, yc, nvalid);
#pragma unroll
                            for (int i = 0; i < GROUP_PX; ++i)
                                if (mbits & (1u << i)) exact_px(true, pixel_word(w, i));
                        }
                        below0 = warp_sum_u(below0); below1 = warp_sum_u(below1);
                        if ((threadIdx.x & 31) == 0) { if (below0) atomicAdd(&sh->l_below[0], below0); if (below1) atomicAdd(&sh->l_below[1], below1); }
                        __syncthreads();
                        if (threadIdx.x == 0) {
                            bool ok = sh->wq_overflow == 0;
                            for (int j = 0; j < 2; ++j)
                                ok = ok && sh->l_len[j] <= (j ? list1 : list0).cap() && sh->l_below[j] <= p_lo[j] && p_hi[j] < sh->l_below[j] + sh->l_len[j];
                            sh->s_ok = ok ? 1 : 0;
                        }
                        __syncthreads();
                        if (sh->s_ok) {
                            {
                                const unsigned r_lo[2] = {p_lo[0] - sh->l_below[0], p_lo[1] - sh->l_below[1]};
                                const unsigned r_hi[2] = {p_hi[0] - sh->l_below[0], p_hi[1] - sh->l_below[1]};
                                unsigned keys[4];
                                list_select_pairs(sh, list0, list1, r_lo, r_hi, keys);
                                if (threadIdx.x < 4) sh->okey[threadIdx.x] = keys[threadIdx.x];
                            }
                            __syncthreads();
                            sampled = true;
                        }
                    }
                }
                if (!sampled) {
                    // ---------------------------------------------------------- B1: angle histogram (tissue pixels)
                    zero_hist(sh);
                    for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int g) {
                        const uint32_t mbits = mask16<decltype(tail)::value>(od_rep, lane_off, w, yc, nvalid);
                        for_each_px_od(od_rep, lane_off, w, [&](int i, float o0, float o1, float o2) {
                            const float px = fmaf(o2, v02, fmaf(o1, v01, o0 * v00));
                            const float py = fmaf(o2, v12, fmaf(o1, v11, o0 * v10));
                            const uint32_t key = angle_key(px, py);
                            if (mbits & (1u << i)) atomicAdd(&sh->hist[key >> L2_BITS], 1u);
                        });
                    });
                    if (threadIdx.x == 0) {
                        unsigned lo, hi; double fr;
                        percentile_index(n_tissue, 100.0 - a.ang_pct, lo, hi, fr);
                        sh->q_rank[0] = lo; sh->q_rank[1] = hi;
                        percentile_index(n_tissue, a.ang_pct, lo, hi, fr);
                        sh->q_rank[2] = lo; sh->q_rank[3] = hi;
                        for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
                    }
                    __syncthreads();
                    tile_sync(S);
                    select_ranks<L1_BINS>(sh, sh->hist, S, sh->q_rank, 4, sh->q_bin, sh->q_rem);
                    if (threadIdx.x == 0) { const unsigned src[4] = {0, 0, 0, 0}; plan_level2(sh, src); }
                    tile_sync(S);
                    // -------------------------------------------------------------- B2: 11-bit refinement
                    zero_hist(sh);
                    {
                        const int nd = sh->n_distinct;
                        const unsigned b0 = sh->d_bin[0], b1 = sh->d_bin[1], b2 = sh->d_bin[2], b3 = sh->d_bin[3];
                        for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int g) {
                            const uint32_t mbits = mask16<decltype(tail)::value>(od_rep, lane_off, w, yc, nvalid);
                            for_each_px_od(od_rep, lane_off, w, [&](int i, float o0, float o1, float o2) {
                                const float px = fmaf(o2, v02, fmaf(o1, v01, o0 * v00));
                                const float py = fmaf(o2, v12, fmaf(o1, v11, o0 * v10));
                                const uint32_t key = angle_key(px, py);
                                const uint32_t bin = key >> L2_BITS, low = key & (L2_BINS - 1);
                                if (mbits & (1u << i)) {
                                    if (bin == b0) atomicAdd(&sh->hist[low], 1u);
                                    if (nd > 1 && bin == b1) atomicAdd(&sh->hist[L2_BINS + low], 1u);
                                    if (nd > 2 && bin == b2) atomicAdd(&sh->hist[2 * L2_BINS + low], 1u);
                                    if (nd > 3 && bin == b3) atomicAdd(&sh->hist[3 * L2_BINS + low], 1u);
                                }
                            });
                        });
                    }
                    __syncthreads();
                    tile_sync(S);
                    for (int q = 0; q < 4; ++q)
                        select_ranks<L2_BINS>(sh, sh->hist + sh->q_hist[q] * L2_BINS, S, &sh->q_rem[q], 1, &sh->q_key[q], &sh->q_tmp[q]);
                    if (threadIdx.x < 4) sh->okey[threadIdx.x] = (sh->q_bin[threadIdx.x] << L2_BITS) | sh->q_key[threadIdx.x];
                    __syncthreads();
                }
                // the fp64 transcendentals of this step are latency-bound on one thread: four threads take one atan2 each,
                // two threads one sincos each
                if (threadIdx.x < 4) sh->ang[threadIdx.x] = angle_from_key(sh->okey[threadIdx.x]);
                __syncthreads();
                if (threadIdx.x < 2) {
                    unsigned lo, hi; double fr;
                    percentile_index(n_tissue, threadIdx.x == 0 ? 100.0 - a.ang_pct : a.ang_pct, lo, hi, fr);
                    const double phi = lerp_np(sh->ang[2 * threadIdx.x], sh->ang[2 * threadIdx.x + 1], fr);
                    sh->cs[2 * threadIdx.x] = cos(phi); sh->cs[2 * threadIdx.x + 1] = sin(phi);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    const double c1 = sh->cs[0], s1 = sh->cs[1], c2 = sh->cs[2], s2 = sh->cs[3];
                    double v1[3], v2[3];
                    for (int k = 0; k < 3; ++k) {
                        v1[k] = sh->D[k] * c1 + sh->D[3 + k] * s1;
                        v2[k] = sh->D[k] * c2 + sh->D[3 + k] * s2;
                    }
                    const bool first = v1[0] > v2[0];
                    const double* h = first ? v1 : v2;
                    const double* e = first ? v2 : v1;
                    const double nh = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), ne = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                    bool ok = true;
                    for (int k = 0; k < 3; ++k) {
                        sh->Msrc[k] = h[k] / nh; sh->Msrc[3 + k] = e[k] / ne;
                        ok = ok && isfinite(sh->Msrc[k]) && isfinite(sh->Msrc[3 + k]);
                    }
                    if (!ok) sh->flags |= SB_STATUS_DEGENERATE;
                }
                tile_sync(S);
            }
        } else {
            // ------------------------------------------------------------------ Vahadane: full-batch dictionary learning
            if (threadIdx.x == 0) {
                const double r0[3] = {0.65, 0.70, 0.29}, r1[3] = {0.07, 0.99, 0.11};
                const double n0 = sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]), n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
                for (int k = 0; k < 3; ++k) { sh->D[k] = r0[k] / n0; sh->D[3 + k] = r1[k] / n1; }
                make_dict_lasso_consts(sh->D, a.dl_lambda, sh->lk);
            }
            __syncthreads();
            // V0: tissue mask -> one bit per pixel (reused by every iteration); tissue counts of the tile and of its sample
            {
                unsigned cnt_tissue = 0, cnt_sample = 0;
                const int nfull = npx / GROUP_PX;
                for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int g) {
                    const uint32_t mbits = mask16<decltype(tail)::value>(od_rep, lane_off, w, yc, nvalid);
                    cnt_tissue += __popc(mbits);
                    if (is_sample_group(g, nfull)) cnt_sample += __popc(mbits);
                    if (cache_mask) *(cache_smem ? mask_slot(sh->hist, g - gb) : a.mask_scratch + (size_t)tile * G + g) = (unsigned short)mbits;
                });
                long long acc[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] = 0;
                acc[0] = (long long)cnt_sample;
                block_reduce10(sh, pbuf, acc, cnt_tissue);
                tile_sync(S);
                cluster_total10(sh, pbuf, S, 1.0);
                pbuf ^= 1;
            }
            n_tissue = (unsigned)sh->tot[9];
            const bool use_sample = a.dl_sample_iters > 0 && sh->tot[0] >= 1024.0;
            __syncthreads();
            if (threadIdx.x == 0 && n_tissue < 1u) sh->flags |= SB_STATUS_EMPTY_MASK;
            __syncthreads();
            // phase 0: warm start on the 1-in-16 sample; phase 1: full passes.  Without a usable sample: 4 more full passes.
            for (int phase = use_sample ? 0 : 1; phase < 2 && sh->flags == 0; ++phase) {
                const int n_it = phase == 0 ? a.dl_sample_iters : a.dl_iters + ((a.dl_sample_iters > 0 && !use_sample) ? 4 : 0);
                // The full passes inherit the difference history of the sample passes: the sample map has (nearly) the same
                // Jacobian, so the first full steps are already quasi-Newton steps; only the residual bookkeeping restarts.
                __syncthreads();     // every thread has read the previous phase's dl_stop before it is cleared
                if (threadIdx.x == 0) { if (phase == 0 || !use_sample) aa_reset(sh->aa); else aa_carry(sh->aa); sh->dl_stop = 0; }
                for (int it = 0; it < n_it; ++it) {
                    const LassoK lk = sh->lk;
                    // Sparse-code the tissue pixels of a group and add their a a^T / x a^T terms.  Atoms on the unit sphere (the
                    // usual case: the norm constraint is active) take the compare-free LASSO two pixels at a time on the packed
                    // f32x2 pipe with packed accumulators; atoms inside the ball take the general KKT form.
                    // Partial sums stay in fp32 for one UNIT of the thread's share (<= 256 pixels, non-negative terms) and are
                    // flushed -- warp shuffle tree, one fixed-point atomic per warp -- at the unit boundary (for_each_unit): the
                    // fp64 / int64 accumulators would not fit the register budget of this loop, and the unit structure makes
                    // the sums independent of the cluster size.
                    float2 f[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) f[i] = make_float2(0.f, 0.f);
                    auto flush = [&]() {
#pragma unroll
                        for (int i = 0; i < 9; ++i) {
                            float v = f[i].x + f[i].y;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&sh->acc64[i], (unsigned long long)to_fix(v, FIX_DL));
                            f[i] = make_float2(0.f, 0.f);
                        }
                    };
                    auto accumulate_unit = [&](auto unit) {
                        return [&](auto tail, const uint32_t (&w)[12], int nvalid, int g) {
                            constexpr int LM = decltype(unit)::value;
                            const uint32_t mbits = cache_mask ? *(cache_smem ? mask_slot(sh->hist, g - gb) : a.mask_scratch + (size_t)tile * G + g)
                                                              : mask16<decltype(tail)::value>(od_rep, lane_off, w, yc, nvalid);
                            for_each_pair_od(od_rep, lane_off, w, [&](int i, float2 o0, float2 o1, float2 o2) {
                                float2 c0, c1;
                                lasso2_unit_pair<LM>(lk, o0, o1, o2, c0, c1);
                                const bool ma = (mbits & (1u << i)) != 0, mb = (mbits & (2u << i)) != 0;
                                c0.x = ma ? c0.x : 0.f; c1.x = ma ? c1.x : 0.f;
                                c0.y = mb ? c0.y : 0.f; c1.y = mb ? c1.y : 0.f;
                                f[0] = __ffma2_rn(c0, c0, f[0]); f[1] = __ffma2_rn(c0, c1, f[1]); f[2] = __ffma2_rn(c1, c1, f[2]);
                                f[3] = __ffma2_rn(o0, c0, f[3]); f[4] = __ffma2_rn(o1, c0, f[4]); f[5] = __ffma2_rn(o2, c0, f[5]);
                                f[6] = __ffma2_rn(o0, c1, f[6]); f[7] = __ffma2_rn(o1, c1, f[7]); f[8] = __ffma2_rn(o2, c1, f[8]);
                            });
                        };
                    };
                    auto run_pass = [&](auto&& body) {
                        if (phase == 0) for_each_sample_group_flush(tin, npx, gb, ge, aligned, body, flush);
                        else for_each_unit(tin, npx, UK, ub, ue, U, aligned, body, flush);
                    };
                    // (make_dict_lasso_consts normalises the atoms: every iterate takes the packed compare-free solver; the sums
                    //  are in its beta space and are scaled back below)
                    if (lk.g01 >= 0.f) run_pass(accumulate_unit(LassoMode<LASSO_UNIT_POS>{}));
                    else run_pass(accumulate_unit(LassoMode<LASSO_UNIT_NEG>{}));
                    __syncthreads();
                    if (threadIdx.x < 10) {
                        sh->part[pbuf][threadIdx.x] = threadIdx.x < 9 ? (long long)sh->acc64[threadIdx.x] : 0ll;
                        sh->acc64[threadIdx.x] = 0ull;
                    }
                    tile_sync(S);
                    cluster_total10(sh, pbuf, S, 1.0 / (double)FIX_DL);
                    pbuf ^= 1;
                    if (threadIdx.x == 0) {
                        { double sc[2]; dict_scales(sh->D, sc); dict_scale_sums(sh->tot, sc); }
                        const double* t = sh->tot;
                        // Mairal et al. 2010 Alg. 2, one block-coordinate sweep; D rows = atoms.
                        double FD[6];
                        for (int k = 0; k < 6; ++k) FD[k] = sh->D[k];
                        const double Aj[2][2] = {{t[0], t[1]}, {t[1], t[2]}};
                        for (int j = 0; j < 2; ++j) {
                            if (Aj[j][j] > 1e-12) {
                                double u[3], nrm = 0.0;
                                for (int k = 0; k < 3; ++k) {
                                    const double Da = FD[k] * Aj[0][j] + FD[3 + k] * Aj[1][j];
                                    u[k] = (t[3 + 3 * j + k] - Da) / Aj[j][j] + FD[3 * j + k];
                                    u[k] = u[k] > 0.0 ? u[k] : 0.0;
                                    nrm += u[k] * u[k];
                                }
                                nrm = sqrt(nrm);
                                const double sc = 1.0 / (nrm > 1.0 ? nrm : 1.0);
                                for (int k = 0; k < 3; ++k) FD[3 * j + k] = u[k] * sc;
                            }
                        }
                        // The sample only has to deliver a starting point within its own sampling error (~3e-3) and a
                        // difference history that is still far above the fp32 noise of the sums: stop it at 1e-4.
                        double rn2 = 0.0;
                        for (int k = 0; k < 6; ++k) rn2 += (FD[k] - sh->D[k]) * (FD[k] - sh->D[k]);
                        const double tol = phase == 0 ? DL_SAMPLE_TOL : DL_FULL_TOL;
                        if (a.dl_anderson > 0 && rn2 < tol * tol) sh->dl_stop = 1;      // this step is still applied, then the phase ends
                        aa_step(sh->aa, a.dl_anderson, sh->D, FD);
                        make_dict_lasso_consts(sh->D, a.dl_lambda, sh->lk);
                    }
                    __syncthreads();
                    if (sh->dl_stop) break;
                }
            }
            if (threadIdx.x == 0 && sh->flags == 0) {
                // vahadane_stain_extractor.py:38-43: H first, rows normalised
                const bool swap = sh->D[0] < sh->D[3];
                const double* h = swap ? sh->D + 3 : sh->D;
                const double* e = swap ? sh->D : sh->D + 3;
                const double nh = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), ne = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
                double m[6];
                bool ok = true;
                for (int k = 0; k < 3; ++k) { m[k] = h[k] / nh; m[3 + k] = e[k] / ne; ok = ok && isfinite(m[k]) && isfinite(m[3 + k]); }
                for (int k = 0; k < 6; ++k) sh->Msrc[k] = m[k];
                if (!ok) sh->flags |= SB_STATUS_DEGENERATE;
            }
            tile_sync(S);
        }

        // stain matrix out
        if (threadIdx.x == 0 && crank == 0 && a.M_out) {
            for (int k = 0; k < 6; ++k) a.M_out[(size_t)tile * 6 + k] = sh->flags ? __longlong_as_double(0x7ff8000000000000LL) : sh->Msrc[k];
        }

        if (a.mode >= PIPE_FIT && sh->flags == 0) {
            if (threadIdx.x == 0) {
                make_lasso_consts(sh->Msrc, a.lasso_lambda, sh->lk);
                sh->s_cnt = 0; sh->l_len[0] = sh->l_len[1] = 0; sh->l_below[0] = sh->l_below[1] = 0; sh->s_ok = 0; sh->wq_overflow = 0;
            }
            zero_hist(sh);
            const LassoK lk = sh->lk;
            unsigned c_lo, c_hi;
            { double fr; percentile_index((unsigned)npx, a.conc_pct, c_lo, c_hi, fr); }
            bool sampled = false;
            if (S == 1 && npx >= 32768) {
                // -------------------------------------------------------------- C0: concentration keys of a 1-in-16 sample
                unsigned scnt = 0;
                for_each_sample_group(tin, npx, gb, ge, aligned, [&](auto, const uint32_t (&w)[12], int, int) {
                    scnt += GROUP_PX;
                    for_each_px_od(od_rep, lane_off, w, [&](int, float o0, float o1, float o2) {
                        float c0, c1;
                        lasso2(lk, o0, o1, o2, c0, c1);
                        atomicAdd(&sh->hist[conc_key(c0) >> L2_BITS], 1u);
                        atomicAdd(&sh->hist[L1_BINS + (conc_key(c1) >> L2_BITS)], 1u);
                    });
                });
                scnt = warp_sum_u(scnt);
                if ((threadIdx.x & 31) == 0 && scnt) atomicAdd(&sh->s_cnt, scnt);
                __syncthreads();
                if (threadIdx.x == 0) {
                    const unsigned n_s = sh->s_cnt;
                    if (n_s >= 1024u) {
                        plan_bracket((unsigned)npx, n_s, c_lo, sh->q_rank[0], sh->q_rank[1], a.bracket_sigmas, a.bracket_pad);
                        sh->q_rank[2] = sh->q_rank[0]; sh->q_rank[3] = sh->q_rank[1];
                        for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
                        sh->s_ok = 1;
                    }
                }
                __syncthreads();
                if (sh->s_ok) {
                    select_ranks<L1_BINS>(sh, sh->hist, 1, sh->q_rank, 2, sh->q_bin, sh->q_rem);
                    select_ranks<L1_BINS>(sh, sh->hist + L1_BINS, 1, sh->q_rank + 2, 2, sh->q_bin + 2, sh->q_rem + 2);
                    if (threadIdx.x < 2) {
                        const unsigned j = threadIdx.x;
                        sh->brk_a[j] = sh->q_bin[2 * j] << L2_BITS;
                        sh->brk_b[j] = (sh->q_bin[2 * j + 1] + 1u) << L2_BITS;
                        // concentrations safely below / above the bracket (64 key units of slack) need no exact key
                        sh->fast_lo[j] = sh->brk_a[j] >= 64u ? float_below(conc_from_key(sh->brk_a[j] - 64u)) : -1.f;
                        sh->fast_hi[j] = sh->brk_b[j] + 64u < (1u << KEY_BITS) ? float_above(conc_from_key(sh->brk_b[j] + 64u)) : INFINITY;
                    }
                    __syncthreads();
                    // ---------------------------------------------------------- C1': count below / collect inside the brackets
                    const unsigned ka0 = sh->brk_a[0], kb0 = sh->brk_b[0], ka1 = sh->brk_a[1], kb1 = sh->brk_b[1];
                    const float lo0 = sh->fast_lo[0], hi0 = sh->fast_hi[0], lo1 = sh->fast_lo[1], hi1 = sh->fast_hi[1];
                    const KeyList list0{sh->hist, ka0, kb0 - ka0 > LIST_SPAN}, list1{sh->hist + L1_BINS, ka1, kb1 - ka1 > LIST_SPAN};
                    unsigned below0 = 0, below1 = 0;
                    auto bracket_pass = [&](auto unit) {
                        constexpr int LM = decltype(unit)::value;
                        // exact treatment of one pixel (packed RGB): a stain whose concentration the float test could not
                        // classify gets its exact key.  The concentrations come from the SAME packed arithmetic as in the
                        // main loop (the pixel is duplicated into both halves), so the same float test decides in both places.
                        auto exact_px = [&](bool has, uint32_t v, bool recount) {
                            const float o0 = od_lookup(od_rep, v, lane_off, 0), o1 = od_lookup(od_rep, v, lane_off, 1), o2 = od_lookup(od_rep, v, lane_off, 2);
                            float2 cc0, cc1;
                            lasso2_unit_pair<LM>(lk, dup(o0), dup(o1), dup(o2), cc0, cc1);
                            const float c0 = cc0.x, c1 = cc1.x;
                            const uint32_t k0 = conc_key(c0), k1 = conc_key(c1);
                            if (has && (recount || (!(c0 < lo0) && !(c0 > hi0)))) {
                                if (k0 < ka0) ++below0;
                                else if (k0 < kb0) { const unsigned idx = atomicAdd(&sh->l_len[0], 1u); if (idx < list0.cap()) list0.put(idx, k0); }
                            }
                            if (has && (recount || (!(c1 < lo1) && !(c1 > hi1)))) {
                                if (k1 < ka1) ++below1;
                                else if (k1 < kb1) { const unsigned idx = atomicAdd(&sh->l_len[1], 1u); if (idx < list1.cap()) list1.put(idx, k1); }
                            }
                        };
                        auto drain_px = [&](bool has, uint32_t pos) { exact_px(has, load_px(tin, pos), false); };
                        WarpQueue wq{sh->wq[threadIdx.x >> 5], 0u};
                        for_each_group_uniform(tin, npx, gb, ge, aligned, [&](const uint32_t (&w)[12], bool active, int g) {
                            unsigned cb0 = 0, cb1 = 0, slow = 0;
                            for_each_pair_od(od_rep, lane_off, w, [&](int i, float2 o0, float2 o1, float2 o2) {
                                float2 c0, c1;
                                lasso2_unit_pair<LM>(lk, o0, o1, o2, c0, c1);
                                // all-ones masks: below the bracket / above it, per stain; a pixel is "slow" when either stain is in neither
                                const uint32_t l0a = set_lt(c0.x, lo0), h0a = set_gt(c0.x, hi0), l1a = set_lt(c1.x, lo1), h1a = set_gt(c1.x, hi1);
                                const uint32_t l0b = set_lt(c0.y, lo0), h0b = set_gt(c0.y, hi0), l1b = set_lt(c1.y, lo1), h1b = set_gt(c1.y, hi1);
                                cb0 -= l0a; cb0 -= l0b; cb1 -= l1a; cb1 -= l1b;
                                slow |= ~((l0a | h0a) & (l1a | h1a)) & (1u << i);
                                slow |= ~((l0b | h0b) & (l1b | h1b)) & (2u << i);
                            });
                            if (active) { below0 += cb0; below1 += cb1; } else slow = 0;
                            wq_push_flagged(wq, slow, (unsigned)g * GROUP_PX, &sh->wq_overflow);
                            wq_drain(wq, false, drain_px);
                        });
                        wq_drain(wq, true, drain_px);
                        if ((npx % GROUP_PX) != 0 && ge == G && threadIdx.x == 0) {
                            uint32_t w[12];
                            int nvalid;
                            load_group<true>(tin, npx, G - 1, false, w, nvalid);
#pragma unroll
                            for (int i = 0; i < GROUP_PX; ++i)
                                if (i < nvalid) exact_px(true, pixel_word(w, i), true);
                        }
                    };
                    const int lm = lasso_mode_of(lk.rg00, lk.rg11, lk.g01);
                    if (lm == LASSO_UNIT_POS) bracket_pass(LassoMode<LASSO_UNIT_POS>{});
                    else if (lm == LASSO_UNIT_NEG) bracket_pass(LassoMode<LASSO_UNIT_NEG>{});
                    else if (threadIdx.x == 0) sh->wq_overflow = 1;   // non-unit rows: robust path
                    below0 = warp_sum_u(below0); below1 = warp_sum_u(below1);
                    if ((threadIdx.x & 31) == 0) { if (below0) atomicAdd(&sh->l_below[0], below0); if (below1) atomicAdd(&sh->l_below[1], below1); }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        bool ok = sh->wq_overflow == 0;
                        for (int j = 0; j < 2; ++j)
                            ok = ok && sh->l_len[j] <= (j ? list1 : list0).cap() && sh->l_below[j] <= c_lo && c_hi < sh->l_below[j] + sh->l_len[j];
                        sh->s_ok = ok ? 1 : 0;
                    }
                    __syncthreads();
                    if (sh->s_ok) {
                        {
                            const unsigned r_lo[2] = {c_lo - sh->l_below[0], c_lo - sh->l_below[1]};
                            const unsigned r_hi[2] = {c_hi - sh->l_below[0], c_hi - sh->l_below[1]};
                            unsigned keys[4];
                            list_select_pairs(sh, list0, list1, r_lo, r_hi, keys);
                            if (threadIdx.x < 4) sh->okey[threadIdx.x] = keys[threadIdx.x];
                        }
                        __syncthreads();
                        sampled = true;
                    }
                }
                if (!sampled) zero_hist(sh);
            }
            if (!sampled) {
                // -------------------------------------------------------------- C1: concentration histograms (all pixels)
                for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int) {
                    constexpr bool TAIL = decltype(tail)::value;
                    for_each_px_od(od_rep, lane_off, w, [&](int i, float o0, float o1, float o2) {
                        float c0, c1;
                        lasso2(lk, o0, o1, o2, c0, c1);
                        if (!TAIL || i < nvalid) {
                            atomicAdd(&sh->hist[conc_key(c0) >> L2_BITS], 1u);
                            atomicAdd(&sh->hist[L1_BINS + (conc_key(c1) >> L2_BITS)], 1u);
                        }
                    });
                });
                if (threadIdx.x == 0) {
                    unsigned lo, hi; double fr;
                    percentile_index((unsigned)npx, a.conc_pct, lo, hi, fr);
                    sh->q_rank[0] = lo; sh->q_rank[1] = hi; sh->q_rank[2] = lo; sh->q_rank[3] = hi;
                    for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
                }
                __syncthreads();
                tile_sync(S);
                select_ranks<L1_BINS>(sh, sh->hist, S, sh->q_rank, 2, sh->q_bin, sh->q_rem);
                select_ranks<L1_BINS>(sh, sh->hist + L1_BINS, S, sh->q_rank + 2, 2, sh->q_bin + 2, sh->q_rem + 2);
                if (threadIdx.x == 0) { const unsigned src[4] = {0, 0, 1, 1}; plan_level2(sh, src); }
                tile_sync(S);
                // ------------------------------------------------------------------ C2: refinement
                zero_hist(sh);
                {
                    const int nd = sh->n_distinct;
                    const unsigned b0 = sh->d_bin[0], b1 = sh->d_bin[1], b2 = sh->d_bin[2], b3 = sh->d_bin[3];
                    const unsigned s0 = sh->d_src[0], s1 = sh->d_src[1], s2 = sh->d_src[2], s3 = sh->d_src[3];
                    for_each_group<true>(tin, npx, gb, ge, aligned, [&](auto tail, const uint32_t (&w)[12], int nvalid, int) {
                        constexpr bool TAIL = decltype(tail)::value;
                        for_each_px_od(od_rep, lane_off, w, [&](int i, float o0, float o1, float o2) {
                            float c0, c1;
                            lasso2(lk, o0, o1, o2, c0, c1);
                            const uint32_t k0 = conc_key(c0), k1 = conc_key(c1);
                            if (!TAIL || i < nvalid) {
                                { const uint32_t k = s0 ? k1 : k0; if ((k >> L2_BITS) == b0) atomicAdd(&sh->hist[k & (L2_BINS - 1)], 1u); }
                                if (nd > 1) { const uint32_t k = s1 ? k1 : k0; if ((k >> L2_BITS) == b1) atomicAdd(&sh->hist[L2_BINS + (k & (L2_BINS - 1))], 1u); }
                                if (nd > 2) { const uint32_t k = s2 ? k1 : k0; if ((k >> L2_BITS) == b2) atomicAdd(&sh->hist[2 * L2_BINS + (k & (L2_BINS - 1))], 1u); }
                                if (nd > 3) { const uint32_t k = s3 ? k1 : k0; if ((k >> L2_BITS) == b3) atomicAdd(&sh->hist[3 * L2_BINS + (k & (L2_BINS - 1))], 1u); }
                            }
                        });
                    });
                }
                __syncthreads();
                tile_sync(S);
                for (int q = 0; q < 4; ++q)
                    select_ranks<L2_BINS>(sh, sh->hist + sh->q_hist[q] * L2_BINS, S, &sh->q_rem[q], 1, &sh->q_key[q], &sh->q_tmp[q]);
                if (threadIdx.x < 4) sh->okey[threadIdx.x] = (sh->q_bin[threadIdx.x] << L2_BITS) | sh->q_key[threadIdx.x];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                double cv[4];
                for (int q = 0; q < 4; ++q) cv[q] = conc_from_key(sh->okey[q]);
                unsigned lo, hi; double fr;
                percentile_index((unsigned)npx, a.conc_pct, lo, hi, fr);
                sh->maxC[0] = lerp_np(cv[0], cv[1], fr);
                sh->maxC[1] = lerp_np(cv[2], cv[3], fr);
                if (crank == 0 && a.maxC_out) { a.maxC_out[(size_t)tile * 2] = sh->maxC[0]; a.maxC_out[(size_t)tile * 2 + 1] = sh->maxC[1]; }
            }
            tile_sync(S);
        } else if (a.mode >= PIPE_FIT && threadIdx.x == 0 && crank == 0 && a.maxC_out) {
            a.maxC_out[(size_t)tile * 2] = a.maxC_out[(size_t)tile * 2 + 1] = __longlong_as_double(0x7ff8000000000000LL);
        }

        if (threadIdx.x == 0 && crank == 0 && a.status) a.status[tile] = sh->flags;
        tile_sync(S);   // protects sh->flags / histograms of the next tile from slow peers
    }
}

template <int METHOD>
static int launch_tile_pipeline_t(const PipeArgs& a, int num_sms, cudaStream_t stream) {
    static DeviceOnce once;      // per instantiation
    const size_t smem = OD_REP_BYTES + sizeof(PipeShared);
    {
        cudaError_t e = ensure_dyn_smem(once, tile_pipeline_kernel<METHOD>, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int S = a.cluster_size;
    int ctas_per_sm = 2;
    int n_clusters = (num_sms * ctas_per_sm) / S;
    if (n_clusters > a.B) n_clusters = a.B;
    if (n_clusters < 1) n_clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * S);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, tile_pipeline_kernel<METHOD>, a);
}

// One instantiation per extraction method: the Macenko and Vahadane chains share only the concentration passes, and a
// single kernel holding both is allocated registers for the worse of the two.
int launch_tile_pipeline(const PipeArgs& a, int num_sms, cudaStream_t stream) {
    return a.method == SB_METHOD_MACENKO ? launch_tile_pipeline_t<SB_METHOD_MACENKO>(a, num_sms, stream)
                                         : launch_tile_pipeline_t<SB_METHOD_VAHADANE>(a, num_sms, stream);
}

// ------------------------------------------------------------------------------------------ slide-level (multi-tile) fit
// ExtractiveStainNormalizer.fit (normalizer.py:27-36) of a SET of tiles treated as one image (a whole slide, possibly
// sharded over ranks): the statistics of the tile chain above are produced as batch-wide sums that add across CTAs,
// launches and ranks -- moments, then two-level key histograms for the exact order statistics -- and the small serial
// steps (covariance, eigenvectors, rank selection, stain matrix) run on the host between the passes
// (stainlib_b200/normalization/slide_fit.py).  Five streaming passes over the target tiles; fit is once per slide.
constexpr int SLIDE_NT = 256;
constexpr int SLIDE_CHUNK_GROUPS = 4096;      // work item = 65,536 pixels of one tile

template <int PASS>
__global__ void __launch_bounds__(SLIDE_NT, 2) slide_pass_kernel(SlideArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* od_rep = smem_raw;
    unsigned* hist = reinterpret_cast<unsigned*>(smem_raw + OD_REP_BYTES);          // 8192 counters
    __shared__ long long red[SLIDE_NT / 32][10];
    const uint32_t lane_off = (threadIdx.x & 31) << 3;
    const YCoef yc{a.ycoef[0], a.ycoef[1], a.ycoef[2], a.ybound};
    for (int i = threadIdx.x; i < 256 * 32; i += SLIDE_NT)
        *reinterpret_cast<float2*>(od_rep + (i >> 5) * OD_ROW_BYTES + (i & 31) * 8) = make_float2(a.tab.od[i >> 5], (float)a.tab.gamma[i >> 5]);
    for (int i = threadIdx.x; i < 2 * L1_BINS; i += SLIDE_NT) hist[i] = 0;
    __syncthreads();
    const int npx = a.npx;
    const int G = (npx + GROUP_PX - 1) / GROUP_PX;
    const int cpt = (G + SLIDE_CHUNK_GROUPS - 1) / SLIDE_CHUNK_GROUPS;
    const long long total = (long long)cpt * a.B;
    const float v00 = a.V[0], v01 = a.V[1], v02 = a.V[2], v10 = a.V[3], v11 = a.V[4], v12 = a.V[5];
    const LassoK lk = a.lk;
    const unsigned b0 = a.bins[0], b1 = a.bins[1], b2 = a.bins[2], b3 = a.bins[3];
    // fixed-point partial sums (the group sums of pass A / the group sums of a dictionary pass): the caller adds the rows of
    // all CTAs, launches and ranks as INTEGERS, so the totals do not depend on how the slide was sharded
    long long acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0;
    unsigned long long cnt = 0;
    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int tile = (int)(item / cpt);
        const int g0 = (int)(item % cpt) * SLIDE_CHUNK_GROUPS;
        const int g1 = min(g0 + SLIDE_CHUNK_GROUPS, G);
        const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
        for (int g = g0 + threadIdx.x; g < g1; g += SLIDE_NT) {
            uint32_t w[12];
            int nvalid;
            load_group<false>(tin, npx, g, a.aligned != 0, w, nvalid);
            if (PASS == 0) {
                float f[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) f[i] = 0.f;
                unsigned c = 0;
                for_each_px_odg(od_rep, lane_off, w, [&](int i, float2 r, float2 gg, float2 b) {
                    float y = tissue_y(yc, r.y, gg.y, b.y);
                    if (i >= nvalid) y = yc.bound;
                    accum_if_tissue(y, yc.bound, r.x, gg.x, b.x, f, c);
                });
                cnt += c;
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += to_fix(f[i], FIX_MOMENT);
            } else if (PASS == 5 || PASS == 6) {
                // Vahadane dictionary pass: sparse codes of the tissue pixels (PASS 6: of the 1-in-16 sample groups of
                // every tile) under the dictionary in lk; sums of a a^T (3) and x a^T (6), and the pixel count
                if (PASS == 6 && !is_sample_group(g, npx / GROUP_PX)) continue;
                float f[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) f[i] = 0.f;
                unsigned c = 0;
                for_each_px_odg(od_rep, lane_off, w, [&](int i, float2 r, float2 gg, float2 b) {
                    float c0, c1;
                    lasso2(lk, r.x, gg.x, b.x, c0, c1);
                    const bool m = (i < nvalid) & (tissue_y(yc, r.y, gg.y, b.y) < yc.bound);
                    c0 = m ? c0 : 0.f; c1 = m ? c1 : 0.f;
                    c += m ? 1u : 0u;
                    f[0] = fmaf(c0, c0, f[0]); f[1] = fmaf(c0, c1, f[1]); f[2] = fmaf(c1, c1, f[2]);
                    f[3] = fmaf(r.x, c0, f[3]); f[4] = fmaf(gg.x, c0, f[4]); f[5] = fmaf(b.x, c0, f[5]);
                    f[6] = fmaf(r.x, c1, f[6]); f[7] = fmaf(gg.x, c1, f[7]); f[8] = fmaf(b.x, c1, f[8]);
                });
                cnt += c;
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += to_fix(f[i], FIX_DL);
            } else if (PASS == 1 || PASS == 2) {
                for_each_px_odg(od_rep, lane_off, w, [&](int i, float2 r, float2 gg, float2 b) {
                    const float px = fmaf(b.x, v02, fmaf(gg.x, v01, r.x * v00));
                    const float py = fmaf(b.x, v12, fmaf(gg.x, v11, r.x * v10));
                    if (i < nvalid && tissue_y(yc, r.y, gg.y, b.y) < yc.bound) {
                        const uint32_t key = angle_key(px, py);
                        const uint32_t bin = key >> L2_BITS, low = key & (L2_BINS - 1);
                        if (PASS == 1) atomicAdd(&hist[bin], 1u);
                        else {
                            if (bin == b0) atomicAdd(&hist[low], 1u);
                            if (bin == b1) atomicAdd(&hist[L2_BINS + low], 1u);
                            if (bin == b2) atomicAdd(&hist[2 * L2_BINS + low], 1u);
                            if (bin == b3) atomicAdd(&hist[3 * L2_BINS + low], 1u);
                        }
                    }
                });
            } else {
                for_each_px_od(od_rep, lane_off, w, [&](int i, float o0, float o1, float o2) {
                    float c0, c1;
                    lasso2(lk, o0, o1, o2, c0, c1);
                    if (i < nvalid) {
                        const uint32_t k0 = conc_key(c0), k1 = conc_key(c1);
                        if (PASS == 3) { atomicAdd(&hist[k0 >> L2_BITS], 1u); atomicAdd(&hist[L1_BINS + (k1 >> L2_BITS)], 1u); }
                        else {
                            if ((k0 >> L2_BITS) == b0) atomicAdd(&hist[k0 & (L2_BINS - 1)], 1u);
                            if ((k0 >> L2_BITS) == b1) atomicAdd(&hist[L2_BINS + (k0 & (L2_BINS - 1))], 1u);
                            if ((k1 >> L2_BITS) == b2) atomicAdd(&hist[2 * L2_BINS + (k1 & (L2_BINS - 1))], 1u);
                            if ((k1 >> L2_BITS) == b3) atomicAdd(&hist[3 * L2_BINS + (k1 & (L2_BINS - 1))], 1u);
                        }
                    }
                });
            }
        }
        if (PASS >= 1 && PASS <= 4) {
            // shared counters are 32-bit: a CTA flushes after every work item (65,536 pixels)
            __syncthreads();
            for (int i = threadIdx.x; i < 2 * L1_BINS; i += SLIDE_NT) {
                const unsigned v = hist[i];
                if (v) { atomicAdd(&a.hist[i], (unsigned long long)v); hist[i] = 0; }
            }
            __syncthreads();
        }
    }
    if (PASS == 0 || PASS >= 5) {
        // per-CTA partial sums, added on the host in CTA order (fixed order for a fixed grid)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = warp_sum_ll(acc[i]);
        const long long c = warp_sum_ll((long long)cnt);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) red[warp][i] = acc[i];
            red[warp][9] = c;
        }
        __syncthreads();
        if (threadIdx.x < 10) {
            long long t = 0;
            for (int wv = 0; wv < SLIDE_NT / 32; ++wv) t += red[wv][threadIdx.x];
            a.sums[(size_t)blockIdx.x * 10 + threadIdx.x] = t;
        }
    }
}

int slide_grid(const SlideArgs& a, int num_sms) {
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const long long total = (long long)((G + SLIDE_CHUNK_GROUPS - 1) / SLIDE_CHUNK_GROUPS) * a.B;
    const long long cap = (long long)num_sms * 2;
    return (int)(total < cap ? total : cap);
}

template <int PASS>
static int launch_slide_pass_t(const SlideArgs& a, int grid, cudaStream_t stream) {
    static DeviceOnce once;      // per instantiation
    const size_t smem = OD_REP_BYTES + 2 * L1_BINS * sizeof(unsigned);
    {
        cudaError_t e = ensure_dyn_smem(once, slide_pass_kernel<PASS>, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    slide_pass_kernel<PASS><<<grid, SLIDE_NT, smem, stream>>>(a);
    return (int)cudaGetLastError();
}

int launch_slide_pass(const SlideArgs& a, int pass, int grid, cudaStream_t stream) {
    switch (pass) {
        case 0: return launch_slide_pass_t<0>(a, grid, stream);
        case 1: return launch_slide_pass_t<1>(a, grid, stream);
        case 2: return launch_slide_pass_t<2>(a, grid, stream);
        case 3: return launch_slide_pass_t<3>(a, grid, stream);
        case 4: return launch_slide_pass_t<4>(a, grid, stream);
        case 5: return launch_slide_pass_t<5>(a, grid, stream);
        default: return launch_slide_pass_t<6>(a, grid, stream);
    }
}

}  // namespace sb
