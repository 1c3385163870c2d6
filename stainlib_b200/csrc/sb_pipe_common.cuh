// sb_pipe_common.cuh -- device building blocks shared by the fused per-tile kernel (sb_pipeline.cu) and the streaming
// per-pass kernels (sb_stream.cu): shared-memory state of a tile, fixed-point block reductions, rank selection in
// histograms and key lists, the sampled-bracket machinery, the rare-pixel compaction queues, table lookups.
// Both kernels run the SAME per-pixel arithmetic (same FMA order) on the SAME keys, so a tile gives the same bits
// whichever of them processes it.
#pragma once
#include "sb_kernels.h"

namespace sb {

constexpr double DL_SAMPLE_TOL = 1e-4;   // residual norm at which the Vahadane sample passes stop (oracle: DL_SAMPLE_TOL)
constexpr double DL_FULL_TOL = 2e-5;     // ... and the full passes: the step taken at the stop leaves <= 1.1e-5 (mean 2e-6) to the fixed point
constexpr unsigned WQ_CAP = 160;   // entries per warp queue: drained to < 32 once per group, a group adds at most 128 kept pushes

struct __align__(16) PipeShared {
    unsigned hist[2 * L1_BINS];      // 32 KB: 2 x 4096 (level 1) or 4 x 2048 (level 2)
    long long red[NWARP][10];       // fixed-point partial sums (integers: associative, so any grouping gives the same bits)
    long long part[2][12];           // this CTA's partial sums (double-buffered; read by cluster peers)
    unsigned long long acc64[10];    // Vahadane: per-pass fixed-point accumulators, fed by one atomic per warp and unit
    double tot[12];
    unsigned wtot[NWARP];
    unsigned q_rank[4], q_bin[4], q_rem[4], q_key[4], q_hist[4], q_tmp[4];
    unsigned d_bin[4], d_src[4];     // distinct level-2 histograms: level-1 bin and key source (0/1)
    int n_distinct;
    unsigned okey[4];                // the four selected 23-bit keys (either selection path writes them)
    // sampled-bracket selection
    unsigned lhist[2][256];
    // 16-byte aligned: the compiler fetches l_len/l_below with one LDS.128, which must not straddle lhist's last word
    // (a block zeroing lhist while a slower thread still loads its ranks is how compute-sanitizer racecheck saw it)
    alignas(16) unsigned l_len[2];
    unsigned l_below[2], l_bin[2], l_rem[2], l_cnt[2], l_min[2], s_cnt;
    double ang[4], cs[4];            // the four selected angles; cos/sin of the two interpolated ones
    unsigned brk_a[2], brk_b[2];
    float fast_lo[2], fast_hi[2];    // conservative float thresholds that let most pixels skip the exact key
    int wq_overflow;
    int s_ok;
    float V[6];
    LassoK lk;
    int flags;
    double D[6];                     // Vahadane dictionary, rows = atoms
    int dl_stop;                     // the current phase has converged (residual below DL_SAMPLE_TOL / DL_FULL_TOL)
    double Msrc[6];
    double maxC[2];
    // ---- the members below are used by the fused kernel only: the per-tile kernels of the streaming path allocate the
    // structure up to here (offsetof(PipeShared, aa)) to fit six CTAs per SM
    AAState aa;                      // Anderson history of the dictionary iteration (thread 0)
    unsigned wq[NWARP][WQ_CAP];      // per-warp compaction queues of the rare pixels that need the exact key
};

__device__ __forceinline__ void tile_sync(int S) {
    if (S > 1) cg::this_cluster().sync(); else __syncthreads();
}

// Per-tile sums are accumulated in FIXED POINT (int64): integer addition is associative, so the totals do not depend on
// how the pixels are grouped into threads, warps, CTAs of a cluster, launches or ranks -- a tile gives the same bits
// whatever cluster size the launcher picked (sharded == unsharded).  The fp32 value that enters a sum is
// itself defined independently of the cluster size: the 16-pixel group sums of pass A, the warp x unit sums of the
// Vahadane passes (see for_each_unit).
constexpr float FIX_MOMENT = 4294967296.f;        // 2^32: group sums <= 16 * ln(255)^2 < 2^9, tiles <= 2^20 groups -> < 2^61
constexpr float FIX_DL = 1073741824.f;            // 2^30: per-pixel terms <= ~2^7, tiles <= 2^24 pixels -> < 2^61
__device__ __forceinline__ long long to_fix(float v, float scale) { return __float2ll_rn(v * scale); }
__device__ __forceinline__ long long warp_sum_ll(long long x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    return x;
}

// Sums 9 fixed-point values + a count over the block into sh->part[buf].
__device__ __forceinline__ void block_reduce10(PipeShared* sh, int buf, long long (&acc)[9], unsigned cnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = warp_sum_ll(acc[i]);
    const long long c = warp_sum_ll((long long)cnt);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) sh->red[warp][i] = acc[i];
        sh->red[warp][9] = c;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        long long s = 0;
        for (int w = 0; w < NWARP; ++w) s += sh->red[w][threadIdx.x];
        sh->part[buf][threadIdx.x] = s;
    }
}

// After tile_sync: every CTA sums the partials of all cluster ranks -> sh->tot (identical everywhere); entries 0..8 are
// scaled back by inv_scale, entry 9 is a plain count.
__device__ __forceinline__ void cluster_total10(PipeShared* sh, int buf, int S, double inv_scale) {
    if (threadIdx.x < 10) {
        long long s = 0;
        if (S > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            for (int r = 0; r < S; ++r) s += cluster.map_shared_rank(&sh->part[buf][0], r)[threadIdx.x];
        } else {
            s = sh->part[buf][threadIdx.x];
        }
        sh->tot[threadIdx.x] = threadIdx.x < 9 ? (double)s * inv_scale : (double)s;
    }
    __syncthreads();
}

// Finds, for nq target ranks, the bin of the (cluster-wide) histogram that contains each rank and the rank inside it.
template <int NB>
__device__ __forceinline__ void select_ranks(PipeShared* sh, const unsigned* hist, int S, const unsigned* ranks, int nq,
                                             unsigned* out_bin, unsigned* out_rem) {
    constexpr int PER = NB / NT;
    static_assert(PER == 4 || PER == 8, "bins per thread");
    unsigned v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = 0;
    for (int r = 0; r < S; ++r) {
        const unsigned* hs = hist;
        if (S > 1) hs = cg::this_cluster().map_shared_rank(hist, r);
        const uint4* p = reinterpret_cast<const uint4*>(hs + threadIdx.x * PER);
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) {
            uint4 x = p[i];
            v[4 * i] += x.x; v[4 * i + 1] += x.y; v[4 * i + 2] += x.z; v[4 * i + 3] += x.w;
        }
    }
    unsigned sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) sum += v[i];
    const unsigned incl = warp_incl_scan(sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 31) sh->wtot[warp] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; ++w) base += sh->wtot[w];
    const unsigned excl = base + incl - sum;
    for (int q = 0; q < nq; ++q) {
        const unsigned r = ranks[q];
        if (r >= excl && r < excl + sum) {
            unsigned c = excl;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                if (r >= c && r < c + v[i]) { out_bin[q] = threadIdx.x * PER + i; out_rem[q] = r - c; }
                c += v[i];
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void zero_hist(PipeShared* sh) {
    uint4* h = reinterpret_cast<uint4*>(sh->hist);
    for (int i = threadIdx.x; i < 2 * L1_BINS / 4; i += NT) h[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
}

// Level-2 bookkeeping (thread 0): queries q (level-1 bin q_bin[q], key source src[q]) -> distinct histograms.
__device__ inline void plan_level2(PipeShared* sh, const unsigned src[4]) {
    int nd = 0;
    for (int q = 0; q < 4; ++q) {
        int found = -1;
        for (int d = 0; d < nd; ++d)
            if (sh->d_bin[d] == sh->q_bin[q] && sh->d_src[d] == src[q]) found = d;
        if (found < 0) { found = nd; sh->d_bin[nd] = sh->q_bin[q]; sh->d_src[nd] = src[q]; ++nd; }
        sh->q_hist[q] = found;
    }
    sh->n_distinct = nd;
}

// Runs f over this CTA's share [gb, ge) of the tile's 16-pixel groups.  Complete groups go through the main loop with
// TAIL = false_type (no validity checks); the single ragged group of a tile whose pixel count is not a multiple of 16
// is handled by one thread with TAIL = true_type.  f(tail, w, nvalid, g) sees the raw 12 words of the group.
template <int LM> struct LassoMode { static constexpr int value = LM; };
struct NoTail { static constexpr bool value = false; };
struct IsTail { static constexpr bool value = true; };

template <bool KEEP, class F>
__device__ __forceinline__ void for_each_group(const uint8_t* __restrict__ tile, int npx, int gb, int ge, bool aligned, F&& f) {
    const int nfull = npx / GROUP_PX;
    const int fe = ge < nfull ? ge : nfull;
    for (int g = gb + threadIdx.x; g < fe; g += NT) {
        uint32_t w[12];
        int nvalid;
        load_group<KEEP>(tile, npx, g, aligned, w, nvalid);
        f(NoTail{}, w, GROUP_PX, g);
    }
    if (ge > nfull && threadIdx.x == 0) {
        uint32_t w[12];
        int nvalid;
        load_group<KEEP>(tile, npx, nfull, false, w, nvalid);
        f(IsTail{}, w, nvalid, nfull);
    }
}

// The lookup table of this kernel holds one {od[v], gamma[v]} PAIR per lane in each 256-byte row: one PRMT builds the
// offset (value << 8 | lane << 3) and ONE conflict-free LDS.64 returns the optical density and the linearised sRGB
// value the tissue test needs (LDS.32 at the same address returns the density alone).
__device__ __forceinline__ void fill_odg_rep(unsigned char* rep, const float* od, const unsigned short* gamma, int nthreads) {
    for (int i = threadIdx.x; i < 256 * 32; i += nthreads)
        *reinterpret_cast<float2*>(rep + (i >> 5) * OD_ROW_BYTES + (i & 31) * 8) = make_float2(od[i >> 5], (float)gamma[i >> 5]);
}
__device__ __forceinline__ float2 odg_lookup(const unsigned char* tab, uint32_t w, uint32_t lane_off, int k) {
    const uint32_t off = __byte_perm(w, lane_off, 0x6504u | (k << 4));
    return *reinterpret_cast<const float2*>(tab + off);
}
// Tissue <=> cv2's L channel below the threshold <=> 871 g[R] + 2929 g[G] + 296 g[B] < ybound (integers < 2^24: the
// fp32 FMAs are exact, so this is the bit-exact mask of stain_utils.py:32-48).
struct YCoef { float r, g, b, bound; };
__device__ __forceinline__ float tissue_y(const YCoef& c, float gr, float gg, float gb) { return fmaf(c.b, gb, fmaf(c.g, gg, c.r * gr)); }

__device__ __forceinline__ uint32_t set_lt(float a, float b) { uint32_t d; asm("set.lt.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t set_gt(float a, float b) { uint32_t d; asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ uint32_t set_le(float a, float b) { uint32_t d; asm("set.le.u32.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// One bit per pixel of the tissue mask, 16 bits per group, for the Vahadane iterations (the 32 KB histogram buffer is
// idle until the concentration passes).  Macenko needs the mask in two passes only and recomputes it: gamma arrives
// with the density in the same LDS.64.
__device__ __forceinline__ unsigned short* mask_slot(unsigned* hist, int gl) { return reinterpret_cast<unsigned short*>(hist) + gl; }
constexpr int MASK_CAP_GROUPS = 2 * L1_BINS * 2;   // 16-bit slots in the histogram buffer = 262,144 pixels per CTA

// Densities AND gammas of the 16 pixels of a group: f(i, {od_r, g_r}, {od_g, g_g}, {od_b, g_b}).
template <class F>
__device__ __forceinline__ void for_each_px_odg(const unsigned char* tab, uint32_t lane_off, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, odg_lookup(tab, a, lane_off, 0), odg_lookup(tab, a, lane_off, 1), odg_lookup(tab, a, lane_off, 2));
        f(4 * q + 1, odg_lookup(tab, a, lane_off, 3), odg_lookup(tab, b, lane_off, 0), odg_lookup(tab, b, lane_off, 1));
        f(4 * q + 2, odg_lookup(tab, b, lane_off, 2), odg_lookup(tab, b, lane_off, 3), odg_lookup(tab, c, lane_off, 0));
        f(4 * q + 3, odg_lookup(tab, c, lane_off, 1), odg_lookup(tab, c, lane_off, 2), odg_lookup(tab, c, lane_off, 3));
    }
}
// 16-bit tissue mask of a group.
template <bool TAIL>
__device__ __forceinline__ uint32_t mask16(const unsigned char* tab, uint32_t lane_off, const uint32_t (&w)[12], const YCoef& yc, int nvalid) {
    uint32_t mbits = 0;
    for_each_px_odg(tab, lane_off, w, [&](int i, float2 r, float2 g, float2 b) {
        uint32_t m = set_lt(tissue_y(yc, r.y, g.y, b.y), yc.bound);
        if (TAIL && i >= nvalid) m = 0u;
        mbits |= m & (1u << i);
    });
    return mbits;
}

// OD of the three channels of the 16 pixels of a group through the replicated table: calls f(i, od_r, od_g, od_b).
template <class F>
__device__ __forceinline__ void for_each_px_od(const unsigned char* tab, uint32_t lane_off, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, od_lookup(tab, a, lane_off, 0), od_lookup(tab, a, lane_off, 1), od_lookup(tab, a, lane_off, 2));
        f(4 * q + 1, od_lookup(tab, a, lane_off, 3), od_lookup(tab, b, lane_off, 0), od_lookup(tab, b, lane_off, 1));
        f(4 * q + 2, od_lookup(tab, b, lane_off, 2), od_lookup(tab, b, lane_off, 3), od_lookup(tab, c, lane_off, 0));
        f(4 * q + 3, od_lookup(tab, c, lane_off, 1), od_lookup(tab, c, lane_off, 2), od_lookup(tab, c, lane_off, 3));
    }
}
// Same, two pixels at a time for the packed f32x2 pipe: f(i0, {od_r(i0), od_r(i0+1)}, {od_g ..}, {od_b ..}).
template <class F>
__device__ __forceinline__ void for_each_pair_od(const unsigned char* tab, uint32_t lane_off, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, f2(od_lookup(tab, a, lane_off, 0), od_lookup(tab, a, lane_off, 3)), f2(od_lookup(tab, a, lane_off, 1), od_lookup(tab, b, lane_off, 0)),
          f2(od_lookup(tab, a, lane_off, 2), od_lookup(tab, b, lane_off, 1)));
        f(4 * q + 2, f2(od_lookup(tab, b, lane_off, 2), od_lookup(tab, c, lane_off, 1)), f2(od_lookup(tab, b, lane_off, 3), od_lookup(tab, c, lane_off, 2)),
          f2(od_lookup(tab, c, lane_off, 0), od_lookup(tab, c, lane_off, 3)));
    }
}

// Pass A inner step: if (y < bound) { n += 1; s += od; S += od x od }  -- ten predicated instructions, no selects.
__device__ __forceinline__ void accum_if_tissue(float y, float bound, float o0, float o1, float o2, float (&f)[9], unsigned& cnt) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.f32 p, %10, %11;\n\t"
        "@p add.f32 %0, %0, %12;\n\t"
        "@p add.f32 %1, %1, %13;\n\t"
        "@p add.f32 %2, %2, %14;\n\t"
        "@p fma.rn.f32 %3, %12, %12, %3;\n\t"
        "@p fma.rn.f32 %4, %12, %13, %4;\n\t"
        "@p fma.rn.f32 %5, %12, %14, %5;\n\t"
        "@p fma.rn.f32 %6, %13, %13, %6;\n\t"
        "@p fma.rn.f32 %7, %13, %14, %7;\n\t"
        "@p fma.rn.f32 %8, %14, %14, %8;\n\t"
        "@p add.u32 %9, %9, 1;\n\t}"
        : "+f"(f[0]), "+f"(f[1]), "+f"(f[2]), "+f"(f[3]), "+f"(f[4]), "+f"(f[5]), "+f"(f[6]), "+f"(f[7]), "+f"(f[8]), "+r"(cnt)
        : "f"(y), "f"(bound), "f"(o0), "f"(o1), "f"(o2));
}

// ------------------------------------------------------------------------------------------ sampled-bracket selection
// Exact order statistics in ONE full pass: a 1-in-16 sample of the groups gives a 4096-bin histogram from which a key
// bracket [ka, kb) around each target rank is chosen (3 sigma of the binomial sampling error plus slack); the full pass
// counts the keys below the bracket and appends the keys inside it to a shared-memory list; the target rank is then
// selected inside the list.  The exact counts prove (or refute) that the rank fell inside the bracket -- on a miss or
// a list overflow the caller falls back to the two-level histogram selection, so the result is always exact.
// Two lists alias the 32 KB histogram buffer, 16 KB each.  A bracket that spans at most 2^16 keys (the dense middle of a
// big tile) stores 16-bit offsets from its start, 8192 entries; a wider one (sparse tails) stores the 4096 full keys.
constexpr unsigned LIST_BYTES = L1_BINS * 4;
constexpr unsigned LIST_SPAN = 65536u;
struct KeyList {
    void* base;
    unsigned start;     // bracket start (offset origin of the narrow form)
    bool wide;
    unsigned bytes = LIST_BYTES;     // size of the buffer at base (16 KB in the fused kernel, 32 KB in the per-tile kernels of the streaming path)
    __device__ __forceinline__ unsigned cap() const { return wide ? bytes / 4 : bytes / 2; }
    __device__ __forceinline__ void put(unsigned idx, unsigned key) const {
        if (wide) static_cast<unsigned*>(base)[idx] = key; else static_cast<unsigned short*>(base)[idx] = (unsigned short)(key - start);
    }
    __device__ __forceinline__ unsigned get(unsigned idx) const {
        return wide ? static_cast<const unsigned*>(base)[idx] : start + static_cast<const unsigned short*>(base)[idx];
    }
};
constexpr int SAMPLE_STRIDE = 16;

__device__ __forceinline__ unsigned warp_sum_u(unsigned x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    return x;
}

// Half-width m (in sample ranks) of the bracket around sample rank pos.  The sample consists of 16-pixel groups, i.e. of
// spatially correlated pixels, so its rank error exceeds the binomial sigma.  Measured on eight pools of 512^2 tiles
// (profiles/r01_bracket_sweep.txt): with 3 sigma + 8 about 2 % of the tiles miss a bracket and pay the two-level fallback
// (1.31-1.70 ms per 1024 tiles depending on the pool), 5 sigma + 16 has no miss on any pool (1.38 ms everywhere), 8 sigma
// overflows the lists.  Small samples (256^2 tiles) do not miss at 3 sigma + 8 and only pay for wider brackets (+6 %),
// and beyond ~80 ranks the brackets of big tiles approach the list capacity: hence the two regimes and the cap.
// sigmas >= 0 (SB_BRACKET_SIGMAS / SB_BRACKET_PAD, sweeps only) overrides the rule with sigmas * sigma + pad.
// cap: largest half-width the key lists can take (80 in the fused kernel, whose lists hold 4096 full keys; the per-tile
// kernels of the streaming path, with 8192, pass stream_bracket_cap(n) -- wide enough for 5 sigma on megapixel tiles too).
__device__ inline void plan_bracket(unsigned n, unsigned n_s, unsigned lo, unsigned& ra, unsigned& rb, float sigmas, float pad, double cap = 80.0,
                                    double wide_sigmas = 5.0) {
    const double q = (double)lo / (double)n;
    const double pos = q * (double)n_s;
    const double sd = sqrt((double)n_s * q * (1.0 - q));
    const double m_narrow = 3.0 * sd + 8.0, m_wide = wide_sigmas * sd + 16.0;
    double m = n_s < 6000u ? m_narrow : fmin(m_wide, fmax(m_narrow, cap));
    if (sigmas >= 0.f) m = (double)sigmas * sd + (double)pad;
    const double a = floor(pos - m), b = ceil(pos + m) + 1.0;
    ra = a < 0.0 ? 0u : (unsigned)a;
    rb = b > (double)(n_s - 1) ? n_s - 1 : (unsigned)b;
}

// The order statistics r_lo[j] and r_hi[j] (= r_lo[j] or r_lo[j] + 1: numpy's two interpolation neighbours) of BOTH
// lists in one sweep: 8-bit radix levels over the two lists side by side (two for offsets, three when a list holds full
// keys; warp j scans list j's 256-bin histogram), then -- only if a rank's upper neighbour is not another copy of the
// same key -- one pass for the smallest key above it.  A dozen block barriers instead of four selections' three dozen.
// key[2j] / key[2j+1] receive the keys of r_lo[j] / r_hi[j].  Whole block calls.
__device__ __forceinline__ void list_select_pairs(PipeShared* sh, const KeyList& l0, const KeyList& l1, const unsigned (&r_lo)[2],
                                                  const unsigned (&r_hi)[2], unsigned* key) {
    const unsigned len[2] = {sh->l_len[0], sh->l_len[1]};
    const unsigned origin[2] = {l0.wide ? 0u : l0.start, l1.wide ? 0u : l1.start};
    unsigned prefix[2] = {0u, 0u}, rank[2] = {r_lo[0], r_lo[1]};
    unsigned mask = 0;
    for (int shift = (l0.wide || l1.wide) ? 16 : 8; shift >= 0; shift -= 8) {
        (&sh->lhist[0][0])[threadIdx.x] = 0;                     // NT == 512 == 2 x 256
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const KeyList& l = j ? l1 : l0;
            for (unsigned i = threadIdx.x; i < len[j]; i += NT) {
                const unsigned k = l.get(i) - origin[j];
                if ((k & mask) == prefix[j]) atomicAdd(&sh->lhist[j][(k >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
            unsigned v[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = sh->lhist[j][lane * 8 + i]; sum += v[i]; }
            const unsigned incl = warp_incl_scan(sum);
            unsigned c = incl - sum;
            if (rank[j] >= c && rank[j] < incl) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (rank[j] >= c && rank[j] < c + v[i]) { sh->l_bin[j] = lane * 8 + i; sh->l_rem[j] = rank[j] - c; sh->l_cnt[j] = v[i]; }
                    c += v[i];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) { prefix[j] |= sh->l_bin[j] << shift; rank[j] = sh->l_rem[j]; }
        mask |= 255u << shift;
    }
    // rank[j] = position of r_lo[j] among the l_cnt[j] copies of its key
    bool next[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) next[j] = r_hi[j] > r_lo[j] && rank[j] + 1u >= sh->l_cnt[j];
    if (next[0] || next[1]) {
        if (threadIdx.x < 2) sh->l_min[threadIdx.x] = 0xFFFFFFFFu;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!next[j]) continue;
            const KeyList& l = j ? l1 : l0;
            unsigned m = 0xFFFFFFFFu;
            for (unsigned i = threadIdx.x; i < len[j]; i += NT) {
                const unsigned k = l.get(i) - origin[j];
                if (k > prefix[j] && k < m) m = k;
            }
            if (m != 0xFFFFFFFFu) atomicMin(&sh->l_min[j], m);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        key[2 * j] = origin[j] + prefix[j];
        key[2 * j + 1] = next[j] ? origin[j] + sh->l_min[j] : origin[j] + prefix[j];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------- rare-pixel compaction queues
// In the bracket passes ~98 % of the pixels are classified by a cheap float test; the rest need the exact 23-bit key
// and possibly a list append.  Handling them in place would make almost every warp step diverge (some lane out of 32
// is nearly always "rare"), so each warp pushes its rare pixels (packed RGB + flags) into a 64-entry shared-memory
// queue with one ballot, and drains the queue 32 entries at a time with all lanes busy.
struct WarpQueue {
    unsigned* q;
    unsigned len;      // warp-uniform
};
// Pushes val for every lane with pred set.  The queue is bounded: a push that does not fit is dropped and reported,
// which makes the caller's validation fail and the tile take the two-level histogram path instead.
__device__ __forceinline__ void wq_push(WarpQueue& wq, bool pred, unsigned val, int* overflow) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m) {
        const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
        const unsigned idx = wq.len + __popc(m & lt);
        if (pred) { if (idx < WQ_CAP) wq.q[idx] = val; else *overflow = 1; }
        wq.len = min(wq.len + __popc(m), WQ_CAP);
    }
}
template <class P>
__device__ __forceinline__ void wq_drain(WarpQueue& wq, bool final, P&& proc) {
    while (wq.len >= 32u || (final && wq.len > 0u)) {
        const unsigned n = wq.len < 32u ? wq.len : 32u;
        const unsigned start = wq.len - n;
        __syncwarp();
        const bool has = (threadIdx.x & 31u) < n;
        const unsigned val = has ? wq.q[start + (threadIdx.x & 31u)] : 0u;
        wq.len = start;
        proc(has, val);
        __syncwarp();
    }
}
// Group epilogue of the bracket passes: every lane pushes the POSITIONS (pixel index in the tile) of the pixels flagged
// in `bits` (16-bit mask over its group) into the warp queue, one per round; the drain loads the three bytes of each
// queued pixel with all 32 lanes busy (they are in L1/L2: the group was just read).  Warp-uniform: all lanes call it.
__device__ __forceinline__ void wq_push_flagged(WarpQueue& wq, unsigned bits, unsigned first_px, int* overflow) {
    while (__ballot_sync(0xffffffffu, bits != 0u)) {
        const bool has = bits != 0u;
        const unsigned val = first_px + (unsigned)(__ffs(bits) - 1);
        bits &= bits - 1u;
        wq_push(wq, has, val, overflow);
    }
}
__device__ __forceinline__ uint32_t load_px(const uint8_t* __restrict__ tile, unsigned px) {
    const uint8_t* p = tile + (size_t)px * 3;
    return (unsigned)__ldg(p) | ((unsigned)__ldg(p + 1) << 8) | ((unsigned)__ldg(p + 2) << 16);
}

// The RGB bytes of pixel i (0..15) of a group as one word (R in byte 0), i a compile-time constant after unrolling.
__device__ __forceinline__ uint32_t pixel_word(const uint32_t (&w)[12], int i) {
    const int q = i >> 2, p = i & 3;
    const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
    switch (p) {
        case 0: return __byte_perm(a, 0u, 0x4210u);
        case 1: return __byte_perm(a, b, 0x0543u) & 0x00FFFFFFu;
        case 2: return __byte_perm(b, c, 0x0432u) & 0x00FFFFFFu;
        default: return __byte_perm(c, 0u, 0x4321u);
    }
}
// Warp-uniform iteration over the complete groups of [gb, ge): every lane of a warp runs the same number of iterations
// (lanes past the end get active = false and an all-white group) so warp votes are legal inside f(w, active, g).
template <class F>
__device__ __forceinline__ void for_each_group_uniform(const uint8_t* __restrict__ tile, int npx, int gb, int ge, bool aligned, F&& f) {
    const int nfull = npx / GROUP_PX;
    const int fe = ge < nfull ? ge : nfull;
    const int lane = threadIdx.x & 31;
    for (int g0 = gb + (int)(threadIdx.x & ~31u); g0 < fe; g0 += NT) {
        const int g = g0 + lane;
        const bool active = g < fe;
        uint32_t w[12];
        if (active) {
            int nvalid;
            load_group<true>(tile, npx, g, aligned, w, nvalid);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) w[i] = 0xFFFFFFFFu;
        }
        f(w, active, g);
    }
}

// Like for_each_group but visits ONE complete group out of every SAMPLE_STRIDE consecutive groups, at a hashed offset
// inside the block (a fixed offset would alias with the row length and sample vertical stripes of the image).
// The sample is defined on the TILE's group index, so it does not depend on how a cluster splits the tile.
__device__ __forceinline__ int sample_group_of_block(int j) { return j * SAMPLE_STRIDE + (int)(((uint32_t)j * 2654435761u) >> 28); }
__device__ __forceinline__ bool is_sample_group(int g, int nfull) {
    const int j = g / SAMPLE_STRIDE;
    return j < nfull / SAMPLE_STRIDE && g == sample_group_of_block(j);
}
template <class F>
__device__ __forceinline__ void for_each_sample_group(const uint8_t* __restrict__ tile, int npx, int gb, int ge, bool aligned, F&& f) {
    const int nfull = npx / GROUP_PX;
    const int jb = gb / SAMPLE_STRIDE;
    int je = (ge + SAMPLE_STRIDE - 1) / SAMPLE_STRIDE;
    if (je > nfull / SAMPLE_STRIDE) je = nfull / SAMPLE_STRIDE;
    for (int j = jb + (int)threadIdx.x; j < je; j += NT) {
        const int g = sample_group_of_block(j);
        if (g < gb || g >= ge) continue;
        uint32_t w[12];
        int nvalid;
        load_group<true>(tile, npx, g, aligned, w, nvalid);
        f(NoTail{}, w, GROUP_PX, g);
    }
}

// ---------------------------------------------------------------------------------------- cluster-size independent sums
// The Vahadane passes keep fp32 partial sums in registers (fp64 / int64 accumulators do not fit the register budget).  To
// make them independent of the cluster size, the tile's groups are cut into U units of K*NT consecutive groups (K depends
// on the tile size only); a cluster of S CTAs splits the tile at unit boundaries; inside a unit thread t always visits
// the groups (u*K + i)*NT + t, i < K.  The fp32 sum of a (warp, unit) pair is therefore the same number for every S; it is
// reduced over the warp by a fixed shuffle tree and enters the fixed-point accumulator with one atomic per warp.
__host__ __device__ __forceinline__ int unit_groups(int G) {
    const int k = G / (8 * NT);
    return k < 1 ? 1 : (k > 16 ? 16 : k);
}
template <class F, class FL>
__device__ __forceinline__ void for_each_unit(const uint8_t* __restrict__ tile, int npx, int K, int ub, int ue, int U, bool aligned, F&& f, FL&& flush) {
    const int nfull = npx / GROUP_PX;
    for (int u = ub; u < ue; ++u) {
        for (int i = 0; i < K; ++i) {
            const int g = (u * K + i) * NT + (int)threadIdx.x;
            if (g < nfull) {
                uint32_t w[12];
                int nvalid;
                load_group<true>(tile, npx, g, aligned, w, nvalid);
                f(NoTail{}, w, GROUP_PX, g);
            }
        }
        if (u == U - 1 && (npx % GROUP_PX) != 0 && threadIdx.x == 0) {      // the ragged last group rides with the last unit
            uint32_t w[12];
            int nvalid;
            load_group<true>(tile, npx, nfull, false, w, nvalid);
            f(IsTail{}, w, nvalid, nfull);
        }
        flush();
    }
}
// The sample groups of [gb, ge) with one flush per warp step (gb is a multiple of 32 sample blocks, so a warp step always
// covers the same 32 blocks whatever the cluster size).
template <class F, class FL>
__device__ __forceinline__ void for_each_sample_group_flush(const uint8_t* __restrict__ tile, int npx, int gb, int ge, bool aligned, F&& f, FL&& flush) {
    const int nfull = npx / GROUP_PX;
    const int jb = gb / SAMPLE_STRIDE;
    int je = (ge + SAMPLE_STRIDE - 1) / SAMPLE_STRIDE;
    if (je > nfull / SAMPLE_STRIDE) je = nfull / SAMPLE_STRIDE;
    for (int j0 = jb + (int)(threadIdx.x & ~31u); j0 < je; j0 += NT) {
        const int j = j0 + (int)(threadIdx.x & 31u);
        if (j < je) {
            const int g = sample_group_of_block(j);
            if (g >= gb && g < ge) {
                uint32_t w[12];
                int nvalid;
                load_group<true>(tile, npx, g, aligned, w, nvalid);
                f(NoTail{}, w, GROUP_PX, g);
            }
        }
        flush();
    }
}

}  // namespace sb
