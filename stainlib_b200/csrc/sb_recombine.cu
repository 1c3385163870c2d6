// sb_recombine.cu -- K4, the fused OD + recombine kernel (normalizer.py:46,48-50), tuned for instruction issue.
//
// Per pixel: 3 OD lookups -> closed-form 2-stain non-negative LASSO -> per-tile scale folded into a 2x3 matrix ->
// 3 x ex2 -> uint8 wrap -> packed stores.  The kernel is issue-bound, so everything is about instruction count:
//   * OD table replicated per lane with a 256-byte row stride: ONE PRMT builds the shared-memory offset
//     (value << 8 | lane << 2) straight from the packed pixel word, and the lookup is bank-conflict free;
//   * all multiply-adds run two pixels at a time on the packed f32x2 pipe (FFMA2 / FADD2.RD, sm_100 only);
//   * the "value >= 2^23" / NaN guard of the unclipped uint8 wrap is hoisted to a block-uniform template flag: it is
//     only compiled in when the target matrix has a negative entry (otherwise 255*exp(.) <= 255 always);
//   * with row-normalised stain matrices (unit Gram diagonal) the whole LASSO case analysis collapses to
//     c_j = max(0, min(a_j, u_j)) (or one 3-input max when the stain vectors have a negative dot product): no compares;
//   * in the TMA kernel the table sits at a 64 KB-aligned shared address, so the PRMT output IS the LDS address.
#include <cstdlib>
#include "sb_kernels.h"
#include "sb_ring.cuh"   // mbarrier / cp.async.bulk helpers shared with the generic ring kernel

namespace sb {

constexpr int RT = 256;                       // threads per CTA
template <bool CHECK, int LM>
__device__ __forceinline__ void recombine_loop(const PointArgs& a, const K4Consts& k, const unsigned char* tab, int tile) {
    const uint8_t* __restrict__ tin = a.in + (size_t)tile * a.npx * 3;
    uint8_t* __restrict__ tout = a.out + (size_t)tile * a.npx * 3;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const uint32_t lane_off = (threadIdx.x & 31) << 2;
    const bool aligned = a.aligned != 0;
    for (int g = blockIdx.y * RT + threadIdx.x; g < G; g += gridDim.y * RT) {
        uint32_t w[12], o[12];
        int nvalid;
        load_group<false>(tin, a.npx, g, aligned, w, nvalid);
        recombine_words<CHECK, LM>(k, tab, w, o, lane_off);
        store_group(tout, a.npx, g, aligned, o);
    }
}

__global__ void __launch_bounds__(RT, 3) recombine_v2_kernel(PointArgs a, const K4Consts* __restrict__ consts) {
    extern __shared__ __align__(256) unsigned char od_rep[];   // [256 values][64 words]; words 0..31 = lane copies
    const int tile = blockIdx.x;
    fill_od_rep(od_rep, a.tab.od, RT);
    __syncthreads();
    const K4Consts k = consts[tile];
    if (k.mode != 0) {
        const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
        uint8_t* tout = a.out + (size_t)tile * a.npx * 3;
        const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
        for (int g = blockIdx.y * RT + threadIdx.x; g < G; g += gridDim.y * RT) {
            uint32_t w[12];
            int nvalid;
            load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
            if (k.mode == 1) {
#pragma unroll
                for (int i = 0; i < 12; ++i) w[i] = 0;
            }
            store_group(tout, a.npx, g, a.aligned != 0, w);
        }
        return;
    }
    const unsigned char* tab = od_rep;
    switch ((k.need_check ? 3 : 0) + k.lasso_mode) {
        case 0: recombine_loop<false, LASSO_GENERAL>(a, k, tab, tile); break;
        case 1: recombine_loop<false, LASSO_UNIT_POS>(a, k, tab, tile); break;
        case 2: recombine_loop<false, LASSO_UNIT_NEG>(a, k, tab, tile); break;
        case 3: recombine_loop<true, LASSO_GENERAL>(a, k, tab, tile); break;
        case 4: recombine_loop<true, LASSO_UNIT_POS>(a, k, tab, tile); break;
        default: recombine_loop<true, LASSO_UNIT_NEG>(a, k, tab, tile); break;
    }
}

// ------------------------------------------------------------------------------------------- K4 v3: TMA-staged ring
// Persistent CTAs (one per SM).  A tile is cut into chunks of TT*48 bytes; thread 0 streams chunks HBM -> shared memory
// with cp.async.bulk (TMA, 1-D) completing on an mbarrier per stage, every thread recombines its own 48-byte group IN
// PLACE in shared memory, and the chunk leaves with one bulk store.  Loads run NSTAGE-1 chunks ahead of the math, both
// directions are fully coalesced by the copy engine, and no thread ever waits on a global load.
// Template parameters: TT compute threads (+ one producer warp), NSTAGE ring slots of TT*48 bytes.

// Per-tile constants for sb_recombine: source matrices + scales given by the caller.
__global__ void k4_prepare_kernel(PointArgs a, K4Consts* out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.B) return;
    double M[6], Mt[6], sc[2];
    for (int j = 0; j < 6; ++j) { M[j] = a.M[(size_t)tile * 6 + j]; Mt[j] = a.Mt[j]; }
    sc[0] = a.scale[(size_t)tile * 2]; sc[1] = a.scale[(size_t)tile * 2 + 1];
    K4Consts k;
    make_k4_consts(M, a.lasso_lambda, sc, Mt, k);
    out[tile] = k;
}

// Per-tile constants for sb_normalize: source statistics from the fused pipeline kernel, target statistics from fit().
// Flagged tiles: a zero / non-finite source percentile writes zeros like the reference's division by zero
// (normalizer.py:48-50); tiles without a stain matrix (empty mask, < 2 tissue pixels, degenerate) are copied through.
__global__ void k4_prepare_normalize_kernel(int B, const double* __restrict__ M_src, const double* __restrict__ maxC_src,
                                            const double* __restrict__ Mt_dev, const double* __restrict__ maxCt, double lam,
                                            int32_t* __restrict__ status, K4Consts* out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= B) return;
    K4Consts k;
    const int st = status[tile];
    if (st != 0) {
        for (int j = 0; j < 6; ++j) { k.m[j] = 0.f; k.A[j] = 0.f; }
        k.nlam = k.i00 = k.i01 = k.i11 = k.rg00 = k.rg11 = k.g01 = 0.f;
        k.lasso_mode = k.need_check = 0;
        k.mode = 2;
    } else {
        double M[6], Mt[6], sc[2];
        for (int j = 0; j < 6; ++j) { M[j] = M_src[(size_t)tile * 6 + j]; Mt[j] = Mt_dev[j]; }
        sc[0] = maxCt[0] / maxC_src[(size_t)tile * 2]; sc[1] = maxCt[1] / maxC_src[(size_t)tile * 2 + 1];
        make_k4_consts(M, lam, sc, Mt, k);
        if (k.mode == 1) status[tile] = st | SB_STATUS_ZERO_MAXC;
    }
    out[tile] = k;
}

template <bool CHECK, int LM>
__device__ __forceinline__ void recombine_group_smem(const K4Consts& k, const OdAbs tab, uint4* grp) {
    const uint4 va = grp[0], vb = grp[1], vc = grp[2];
    const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
    uint32_t o[12];
    recombine_words<CHECK, LM>(k, tab, w, o, 0u);
    grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
    grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
    grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
}

// Shared-memory plan of the TMA kernel.  The CTA asks for the whole 227 KB; the 64 KB lane-replicated OD table is put
// at the 64 KB-aligned address inside the window (so lookups need no address add), the mbarriers at the window start,
// and the ring slots fill the space in front of and behind the table.
constexpr int K4_SMEM_BYTES = 227 * 1024;
constexpr int K4_BAR_BYTES = 256;

// Template parameters: NG consumer groups of GT threads (+ one producer warp), NSTAGE ring slots of GT*48 bytes.  Group
// g recombines the CTA's chunks g, g+NG, ...: two groups of 15 warps keep 30 warps resident (the 16-pixel body needs
// 56 registers when the compiler is held to that occupancy) and work on different slots at different phases.
template <int GT, int NG, int NSTAGE>
__global__ void __launch_bounds__(GT * NG + 32, 1) recombine_tma_kernel(PointArgs a, const K4Consts* __restrict__ consts, int chunks_per_tile, long long total_chunks) {
    constexpr int TT = GT * NG;
    constexpr int TT_ALL = TT + 32;
    constexpr int CHUNK_BYTES = GT * 48;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t tab_addr = (base + 0xFFFFu) & ~0xFFFFu;
    unsigned char* od_rep = smem + (tab_addr - base);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                            // TMA load landed
    uint64_t* done = full + NSTAGE;                                                // all compute warps wrote the stage back
    const int n_front = tab_addr - base >= (uint32_t)K4_BAR_BYTES ? (int)((tab_addr - base - K4_BAR_BYTES) / CHUNK_BYTES) : 0;
    const int n_back = ((int)K4_SMEM_BYTES - (int)(tab_addr - base) - OD_REP_BYTES) / CHUNK_BYTES;
    if (n_front + n_back < NSTAGE || tab_addr - base < (uint32_t)K4_BAR_BYTES) __trap();   // launch_tma_variant sized the window for this
    auto stage_ptr = [&](int s) -> unsigned char* {
        return s < n_front ? smem + K4_BAR_BYTES + (size_t)s * CHUNK_BYTES : od_rep + OD_REP_BYTES + (size_t)(s - n_front) * CHUNK_BYTES;
    };
    const size_t tile_bytes = (size_t)a.npx * 3;
    const long long c_begin = total_chunks * blockIdx.x / gridDim.x, c_end = total_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);

    auto chunk_geom = [&](long long c, int& tile, size_t& off, uint32_t& bytes) {
        tile = (int)(c / chunks_per_tile);
        off = (size_t)(c % chunks_per_tile) * CHUNK_BYTES;
        const size_t rem = tile_bytes - off;
        bytes = (uint32_t)(rem < (size_t)CHUNK_BYTES ? rem : (size_t)CHUNK_BYTES);
    };

    if (threadIdx.x == TT) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], GT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256 * 32; i += TT_ALL)
        *reinterpret_cast<float*>(od_rep + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = a.tab.od[i >> 5];
    __syncthreads();

    if (threadIdx.x >= TT) {
        // ------------------------------------------------------------------ producer warp (one elected lane)
        if (threadIdx.x == TT) {
            for (int i = 0; i < NSTAGE && i < n_local; ++i) {
                int tile; size_t off; uint32_t bytes;
                chunk_geom(c_begin + i, tile, off, bytes);
                mbar_expect_tx(&full[i], bytes);
                bulk_load(stage_ptr(i), a.in + (size_t)tile * tile_bytes + off, bytes, &full[i]);
            }
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                int tile; size_t off; uint32_t bytes;
                chunk_geom(c_begin + i, tile, off, bytes);
                mbar_wait(&done[s], (uint32_t)((i / NSTAGE) & 1));          // stage s holds the finished output of chunk i
                bulk_store(a.out + (size_t)tile * tile_bytes + off, stage_ptr(s), bytes);
                // refill the stage of chunk i-1 once its store has finished reading shared memory
                if (i >= 1 && i - 1 + NSTAGE < n_local) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int ps = (i - 1) % NSTAGE;
                    int t2; size_t o2; uint32_t b2;
                    chunk_geom(c_begin + i - 1 + NSTAGE, t2, o2, b2);
                    mbar_expect_tx(&full[ps], b2);
                    bulk_load(stage_ptr(ps), a.in + (size_t)t2 * tile_bytes + o2, b2, &full[ps]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warps
    const OdAbs tab{((threadIdx.x & 31u) << 2) | ((tab_addr >> 16) << 8)};
    const bool copy_only = a.debug_copy != 0;
    const int gidx = threadIdx.x / GT, tig = threadIdx.x - gidx * GT;
    int i = gidx;
    while (i < n_local) {
      // run of this group's chunks that belong to one tile: constants are loaded once per run
      const int i0 = i;
      const int tile = (int)((c_begin + i0) / chunks_per_tile);
      const int first_in_tile = (int)((c_begin + i0) - (long long)tile * chunks_per_tile);
      int i_last = i0 + (chunks_per_tile - 1 - first_in_tile);
      if (i_last > n_local - 1) i_last = n_local - 1;
      const K4Consts k = consts[tile];
      const int variant = (copy_only || k.mode == 2) ? 7 : (k.mode == 1 ? 6 : (k.need_check ? 3 : 0) + k.lasso_mode);
      for (; i <= i_last; i += NG) {
        const size_t off = (size_t)(first_in_tile + (i - i0)) * CHUNK_BYTES;
        const int s = i % NSTAGE;
        const uint32_t parity = (uint32_t)((i / NSTAGE) & 1);
        const size_t rem = tile_bytes - off;
        const uint32_t bytes = (uint32_t)(rem < (size_t)CHUNK_BYTES ? rem : (size_t)CHUNK_BYTES);
        unsigned char* buf = stage_ptr(s);
        mbar_wait(&full[s], parity);
        if (tig * 48u < bytes) {
            uint4* grp = reinterpret_cast<uint4*>(buf + tig * 48u);
            switch (variant) {
                case 0: recombine_group_smem<false, LASSO_GENERAL>(k, tab, grp); break;
                case 1: recombine_group_smem<false, LASSO_UNIT_POS>(k, tab, grp); break;
                case 2: recombine_group_smem<false, LASSO_UNIT_NEG>(k, tab, grp); break;
                case 3: recombine_group_smem<true, LASSO_GENERAL>(k, tab, grp); break;
                case 4: recombine_group_smem<true, LASSO_UNIT_POS>(k, tab, grp); break;
                case 5: recombine_group_smem<true, LASSO_UNIT_NEG>(k, tab, grp); break;
                case 6: grp[0] = grp[1] = grp[2] = make_uint4(0, 0, 0, 0); break;
                default: break;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk store
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
      }
    }
}

template <int GT, int NG, int NSTAGE>
static int launch_tma_variant(const PointArgs& a, int num_sms, cudaStream_t stream, const K4Consts* consts) {
    constexpr int CHUNK_BYTES = GT * 48;
    static_assert(GT % 32 == 0 && GT * NG + 32 <= 1024, "block size");
    // the table must sit on a 64 KB boundary of the shared window wherever the window starts: take the whole 227 KB
    static_assert(OD_REP_BYTES + NSTAGE * CHUNK_BYTES + 1024 + K4_BAR_BYTES <= K4_SMEM_BYTES, "ring does not fit");
    static_assert(2 * NSTAGE * 8 <= K4_BAR_BYTES, "barrier area");
    const int smem_bytes = K4_SMEM_BYTES;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(recombine_tma_kernel<GT, NG, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const size_t tile_bytes = (size_t)a.npx * 3;
    const int cpt = (int)((tile_bytes + CHUNK_BYTES - 1) / CHUNK_BYTES);
    const long long total = (long long)cpt * a.B;
    int grid = num_sms;
    if ((long long)grid > total) grid = (int)total;
    recombine_tma_kernel<GT, NG, NSTAGE><<<grid, GT * NG + 32, smem_bytes, stream>>>(a, consts, cpt, total);
    return (int)cudaGetLastError();
}

// Runs K4 over the batch with the given per-tile constants: TMA ring when every tile is a whole number of 16-byte
// vectors at a 16-byte aligned address, register-staged kernel otherwise.
static int launch_k4(const PointArgs& a, int num_sms, cudaStream_t stream, const K4Consts* consts, bool use_tma) {
    if (use_tma) {
        // ring geometry: 512 compute threads x 6 slots of 24 KB (chunks divide 256^2 and 512^2 tiles evenly); two slots sit
        // in front of the 64 KB-aligned table, four behind it
        // (measured alternatives, 1024 x 512^2 tiles: 2 groups x 480 threads 0.420 ms, 3 x 320 0.358 ms, 3 x 256 0.404 ms,
        //  1 x 512 0.357 ms -- more resident warps do not help, the kernel is issue-bound, not latency-bound)
        return launch_tma_variant<512, 1, 6>(a, num_sms, stream, consts);
    }
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(recombine_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OD_REP_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    // each CTA fills a 64 KB table, so give it at least ~64k pixels; aim for >= 2 waves of (SMs x 3) CTAs
    int spans = (G + RT * 16 - 1) / (RT * 16);
    int want = (num_sms * 3 * 2 + a.B - 1) / a.B;
    if (want < 1) want = 1;
    if (spans > want) spans = want;
    if (spans < 1) spans = 1;
    recombine_v2_kernel<<<dim3(a.B, spans), RT, OD_REP_BYTES, stream>>>(a, consts);
    return (int)cudaGetLastError();
}

int launch_recombine(const PointArgs& a, int num_sms, cudaStream_t stream, bool use_tma) {
    K4Consts* consts = nullptr;
    cudaError_t e = cudaMallocAsync(&consts, (size_t)a.B * sizeof(K4Consts), stream);
    if (e != cudaSuccess) return (int)e;
    k4_prepare_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a, consts);
    const int rc = launch_k4(a, num_sms, stream, consts, use_tma);
    cudaFreeAsync(consts, stream);
    return rc;
}

int launch_recombine_normalize(const PointArgs& a, int num_sms, cudaStream_t stream, bool use_tma, const double* M_src,
                               const double* maxC_src, const double* Mt, const double* maxCt, int32_t* status) {
    K4Consts* consts = nullptr;
    cudaError_t e = cudaMallocAsync(&consts, (size_t)a.B * sizeof(K4Consts), stream);
    if (e != cudaSuccess) return (int)e;
    k4_prepare_normalize_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a.B, M_src, maxC_src, Mt, maxCt, a.lasso_lambda, status, consts);
    const int rc = launch_k4(a, num_sms, stream, consts, use_tma);
    cudaFreeAsync(consts, stream);
    return rc;
}

}  // namespace sb
