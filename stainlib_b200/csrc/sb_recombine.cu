// sb_recombine.cu -- K4, the fused OD + recombine kernel (normalizer.py:46,48-50), tuned for instruction issue.
//
// Per pixel: 3 OD lookups -> closed-form 2-stain non-negative LASSO -> per-tile scale folded into a 2x3 matrix ->
// 3 x ex2 -> uint8 wrap -> packed stores.  The kernel is issue-bound, so everything is about instruction count:
//   * OD table replicated per lane with a 256-byte row stride: ONE PRMT builds the shared-memory offset
//     (value << 8 | lane << 2) straight from the packed pixel word, and the lookup is bank-conflict free;
//   * all multiply-adds run two pixels at a time on the packed f32x2 pipe (FFMA2 / FADD2.RD, sm_100 only);
//   * the "value >= 2^23" / NaN guard of the unclipped uint8 wrap is hoisted to a block-uniform template flag: it is
//     only compiled in when the target matrix has a negative entry (otherwise 255*exp(.) <= 255 always);
//   * with row-normalised stain matrices (unit Gram diagonal) the single-active-stain case of the LASSO reduces to
//     "keep the larger of max(u0,0), max(u1,0)".
#include <cstdlib>
#include "sb_kernels.h"

namespace sb {

constexpr int RT = 256;                       // threads per CTA
template <bool CHECK, bool UNIT>
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
        recombine_words<CHECK, UNIT>(k, tab, w, o, lane_off);
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
    if (k.need_check) {
        if (k.unit_diag) recombine_loop<true, true>(a, k, od_rep, tile); else recombine_loop<true, false>(a, k, od_rep, tile);
    } else {
        if (k.unit_diag) recombine_loop<false, true>(a, k, od_rep, tile); else recombine_loop<false, false>(a, k, od_rep, tile);
    }
}

// ------------------------------------------------------------------------------------------- K4 v3: TMA-staged ring
// Persistent CTAs (one per SM).  A tile is cut into chunks of TT*48 bytes; thread 0 streams chunks HBM -> shared memory
// with cp.async.bulk (TMA, 1-D) completing on an mbarrier per stage, every thread recombines its own 48-byte group IN
// PLACE in shared memory, and the chunk leaves with one bulk store.  Loads run NSTAGE-1 chunks ahead of the math, both
// directions are fully coalesced by the copy engine, and no thread ever waits on a global load.
// Template parameters: TT compute threads (+ one producer warp), NSTAGE ring slots of TT*48 bytes.

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

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
        k.unit_diag = k.need_check = 0;
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

template <bool CHECK, bool UNIT>
__device__ __forceinline__ void recombine_group_smem(const K4Consts& k, const unsigned char* tab, uint4* grp, uint32_t lane_off) {
    const uint4 va = grp[0], vb = grp[1], vc = grp[2];
    const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
    uint32_t o[12];
    recombine_words<CHECK, UNIT>(k, tab, w, o, lane_off);
    grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
    grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
    grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
}

template <int TT, int NSTAGE>
__global__ void __launch_bounds__(TT + 32, 1) recombine_tma_kernel(PointArgs a, const K4Consts* __restrict__ consts, int chunks_per_tile, long long total_chunks) {
    constexpr int TT_ALL = TT + 32;
    constexpr int CHUNK_BYTES = TT * 48;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* od_rep = smem;                                   // 64 KB lane-replicated OD table
    unsigned char* stage0 = smem + OD_REP_BYTES;                    // NSTAGE x CHUNK_BYTES
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + NSTAGE * CHUNK_BYTES);   // TMA load landed
    uint64_t* done = full + NSTAGE;                                                // all compute threads wrote the stage back
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
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], TT / 32); }
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
                bulk_load(stage0 + (size_t)i * CHUNK_BYTES, a.in + (size_t)tile * tile_bytes + off, bytes, &full[i]);
            }
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                int tile; size_t off; uint32_t bytes;
                chunk_geom(c_begin + i, tile, off, bytes);
                mbar_wait(&done[s], (uint32_t)((i / NSTAGE) & 1));          // stage s holds the finished output of chunk i
                bulk_store(a.out + (size_t)tile * tile_bytes + off, stage0 + (size_t)s * CHUNK_BYTES, bytes);
                // refill the stage of chunk i-1 once its store has finished reading shared memory
                if (i >= 1 && i - 1 + NSTAGE < n_local) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int ps = (i - 1) % NSTAGE;
                    int t2; size_t o2; uint32_t b2;
                    chunk_geom(c_begin + i - 1 + NSTAGE, t2, o2, b2);
                    mbar_expect_tx(&full[ps], b2);
                    bulk_load(stage0 + (size_t)ps * CHUNK_BYTES, a.in + (size_t)t2 * tile_bytes + o2, b2, &full[ps]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warps
    const uint32_t lane_off = (threadIdx.x & 31) << 2;
    const bool copy_only = a.debug_copy != 0;
    int i = 0;
    while (i < n_local) {
        // run of chunks that belong to one tile: constants are loaded once per run
        const int tile = (int)((c_begin + i) / chunks_per_tile);
        const int first_in_tile = (int)((c_begin + i) % chunks_per_tile);
        int run = chunks_per_tile - first_in_tile;
        if (run > n_local - i) run = n_local - i;
        const K4Consts k = consts[tile];
        const int variant = (copy_only || k.mode == 2) ? 5 : (k.mode == 1 ? 4 : (k.need_check ? 2 : 0) + (k.unit_diag ? 1 : 0));
        for (int j = 0; j < run; ++j, ++i) {
            const int s = i % NSTAGE;
            const uint32_t parity = (uint32_t)((i / NSTAGE) & 1);
            const size_t off = (size_t)(first_in_tile + j) * CHUNK_BYTES;
            const size_t rem = tile_bytes - off;
            const uint32_t bytes = (uint32_t)(rem < (size_t)CHUNK_BYTES ? rem : (size_t)CHUNK_BYTES);
            unsigned char* buf = stage0 + (size_t)s * CHUNK_BYTES;
            mbar_wait(&full[s], parity);
            if (threadIdx.x * 48u < bytes) {
                uint4* grp = reinterpret_cast<uint4*>(buf + threadIdx.x * 48u);
                switch (variant) {
                    case 0: recombine_group_smem<false, false>(k, od_rep, grp, lane_off); break;
                    case 1: recombine_group_smem<false, true>(k, od_rep, grp, lane_off); break;
                    case 2: recombine_group_smem<true, false>(k, od_rep, grp, lane_off); break;
                    case 3: recombine_group_smem<true, true>(k, od_rep, grp, lane_off); break;
                    case 4: grp[0] = grp[1] = grp[2] = make_uint4(0, 0, 0, 0); break;
                    default: break;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk store
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
        }
    }
}

template <int TT, int NSTAGE>
static int launch_tma_variant(const PointArgs& a, int num_sms, cudaStream_t stream, const K4Consts* consts) {
    constexpr int CHUNK_BYTES = TT * 48;
    const int smem_bytes = OD_REP_BYTES + NSTAGE * CHUNK_BYTES + 2 * NSTAGE * 8;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(recombine_tma_kernel<TT, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const size_t tile_bytes = (size_t)a.npx * 3;
    const int cpt = (int)((tile_bytes + CHUNK_BYTES - 1) / CHUNK_BYTES);
    const long long total = (long long)cpt * a.B;
    int grid = num_sms;
    if ((long long)grid > total) grid = (int)total;
    recombine_tma_kernel<TT, NSTAGE><<<grid, TT + 32, smem_bytes, stream>>>(a, consts, cpt, total);
    return (int)cudaGetLastError();
}

// Runs K4 over the batch with the given per-tile constants: TMA ring when every tile is a whole number of 16-byte
// vectors at a 16-byte aligned address, register-staged kernel otherwise.
static int launch_k4(const PointArgs& a, int num_sms, cudaStream_t stream, const K4Consts* consts, bool use_tma) {
    if (use_tma) {
        // ring geometry: 512 compute threads x 6 slots of 24 KB (chunks divide 256^2 and 512^2 tiles evenly); the sweep in
        // profiles/r01_k4_ring_sweep.txt shows 512x6, 640x5 and 768x4 within 2 % of each other
        static int variant = -1;
        if (variant < 0) { const char* v = getenv("SB_K4_TT"); variant = v ? atoi(v) : 512; }
        if (variant == 640) return launch_tma_variant<640, 5>(a, num_sms, stream, consts);
        if (variant == 768) return launch_tma_variant<768, 4>(a, num_sms, stream, consts);
        return launch_tma_variant<512, 6>(a, num_sms, stream, consts);
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
