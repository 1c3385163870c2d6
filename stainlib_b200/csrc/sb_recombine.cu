// sb_recombine.cu -- K4, the fused OD + recombine kernel (normalizer.py:46,48-50), tuned for instruction issue.
//
// Per pixel: 3 OD lookups -> closed-form 2-stain non-negative LASSO -> per-tile scale folded into a 2x3 matrix ->
// 3 x ex2 -> uint8 wrap -> packed stores.  The kernel is issue-bound, so everything is about instruction count:
//   * OD table replicated per lane with a 256-byte row stride: ONE PRMT builds the shared-memory offset
//     (value << 8 | lane << 2) straight from the packed pixel word, and the lookup is bank-conflict free;
//   * all multiply-adds run two pixels at a time on the packed f32x2 pipe (FFMA2 / FADD2.RD, sm_100 only);
//   * the "value >= 2^23" / NaN guard of the unclipped uint8 wrap is hoisted to a per-tile code variant: it only runs
//     when the per-tile constants cannot PROVE 255*2^e < 2^23 for every uint8 pixel (make_k4_consts);
//   * with row-normalised stain matrices (unit Gram diagonal) the whole LASSO case analysis collapses to
//     c_j = max(0, min(a_j, u_j)) (or one 3-input max when the stain vectors have a negative dot product): no compares;
//   * on the TMA ring (sb_ring.cuh, the path for 16-byte aligned tiles) the table sits at a 64 KB-aligned shared
//     address, so the PRMT output IS the LDS address; unaligned tiles take the register-staged recombine_v2_kernel.
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

// ------------------------------------------------------------------------------------------- K4 on the TMA-staged ring
// The transport (persistent CTAs, cp.async.bulk loads into a 6-slot shared-memory ring, in-place transform, bulk store,
// 64 KB-aligned lookup table) lives in sb_ring.cuh; K4 is its K4Op.  Measured alternatives with more resident warps
// (2 x 480, 3 x 320, 3 x 256 consumer threads): 0.36-0.42 ms against 0.33 ms -- the kernel is issue-bound, not latency-bound.

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

// K4 as an Op of the generic ring (sb_ring.cuh).
struct K4RingParams { const K4Consts* consts; const float* od; int copy_only; };
struct K4Op {
    using Consts = K4Consts;
    using Params = K4RingParams;
    struct Acc {};
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = p.od[i >> 5];
    }
    __device__ static void acc_init(Acc&) {}
    using Run = int;                 // code variant of the tile: wrap check x LASSO mode, zero fill, copy through
    __device__ static Run begin_run(const Consts& k, const Params& p) {
        return (p.copy_only || k.mode == 2) ? 7 : (k.mode == 1 ? 6 : (k.need_check ? 3 : 0) + k.lasso_mode);
    }
    __device__ static void process(const Consts& k, const Params&, const Run& variant, const OdAbs tab, uint4* grp, Acc&) {
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
    }
    __device__ static void finish_run(const Params&, int, Acc&) {}
};

// Runs K4 over the batch with the given per-tile constants: TMA ring when every tile is a whole number of 16-byte
// vectors at a 16-byte aligned address, register-staged kernel otherwise.
static int launch_k4(const PointArgs& a, int num_sms, cudaStream_t stream, const K4Consts* consts, bool use_tma) {
    if (use_tma)   // 512 compute threads x 6 slots of 24 KB (chunks divide 256^2 and 512^2 tiles evenly) on the generic ring
        return launch_ring<K4Op>(RingGeom{a.in, a.out, a.B, a.npx}, K4RingParams{consts, a.tab.od, a.debug_copy}, num_sms, stream);
    static DeviceOnce once;
    {
        cudaError_t e = ensure_dyn_smem(once, recombine_v2_kernel, OD_REP_BYTES);
        if (e != cudaSuccess) return (int)e;
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

int launch_recombine(const PointArgs& a, Scratch& scratch, bool use_tma) {
    K4Consts* consts = nullptr;
    cudaError_t e = scratch.get(&consts, (size_t)a.B * sizeof(K4Consts));
    if (e != cudaSuccess) return (int)e;
    k4_prepare_kernel<<<(a.B + 127) / 128, 128, 0, scratch.st>>>(a, consts);
    return launch_k4(a, scratch.h->num_sms, scratch.st, consts, use_tma);
}

int launch_recombine_normalize(const PointArgs& a, Scratch& scratch, bool use_tma, const double* M_src,
                               const double* maxC_src, const double* Mt, const double* maxCt, int32_t* status) {
    K4Consts* consts = nullptr;
    cudaError_t e = scratch.get(&consts, (size_t)a.B * sizeof(K4Consts));
    if (e != cudaSuccess) return (int)e;
    k4_prepare_normalize_kernel<<<(a.B + 127) / 128, 128, 0, scratch.st>>>(a.B, M_src, maxC_src, Mt, maxCt, a.lasso_lambda, status, consts);
    return launch_k4(a, scratch.h->num_sms, scratch.st, consts, use_tma);
}

}  // namespace sb
