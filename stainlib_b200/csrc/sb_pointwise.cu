// sb_pointwise.cu -- per-pixel kernels that need no per-tile reduction (sm_100a).
//
//   mask_kernel            LuminosityThresholdTissueLocator.get_tissue_mask   stain_utils.py:32-48
//   (K4, the fused OD+recombine kernel, lives in sb_recombine.cu)
//   stain_augment_kernel   StainAugmentor.pop                                  augmenter.py:428-449
//   concentrations_kernel  get_concentrations                                  stain_utils.py:69-78
//
// Work unit = one 16-pixel group (48 B, three 16-byte vectors); a tile is a whole number of CTA-sized spans so the
// per-tile constants are block-uniform.  grid = (spans_per_tile, B).
#include "sb_kernels.h"
#include "sb_ring.cuh"

namespace sb {

constexpr int PT = 256;   // threads per CTA for the pointwise kernels

struct __align__(16) PointShared {
    float od[256];
    float gy[768];
    LassoK lk;
    float A[6];
    float alpha[2], beta[2];
    int zero_out;
};

__device__ __forceinline__ void load_tables(PointShared* sh, const Tables& t, bool with_gy) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh->od[i] = t.od[i];
    if (with_gy)
        for (int i = threadIdx.x; i < 768; i += blockDim.x) sh->gy[i] = t.gy[i];
}

__global__ void __launch_bounds__(PT) mask_kernel(PointArgs a) {
    __shared__ PointShared sh;
    load_tables(&sh, a.tab, true);
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int tile = blockIdx.x;
    const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
    uint8_t* mout = a.mask_out + (size_t)tile * a.npx;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const float* gyR = sh.gy, *gyG = sh.gy + 256, *gyB = sh.gy + 512;
    int found = 0;
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < G; g += gridDim.y * blockDim.x) {
        uint32_t w[12];
        int nvalid;
        load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
        uint32_t m[16];
        for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
            m[i] = (gyR[r] + gyG[gg] + gyB[b] < a.ybound) ? 1u : 0u;
            found |= (i < nvalid) ? m[i] : 0u;
        });
        uint8_t* dst = mout + (size_t)g * GROUP_PX;
        if (nvalid == GROUP_PX && (a.npx % 16) == 0) {
            uint4 v;
            v.x = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
            v.y = m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24);
            v.z = m[8] | (m[9] << 8) | (m[10] << 16) | (m[11] << 24);
            v.w = m[12] | (m[13] << 8) | (m[14] << 16) | (m[15] << 24);
            *reinterpret_cast<uint4*>(dst) = v;
        } else {
            for (int i = 0; i < nvalid; ++i) dst[i] = (uint8_t)m[i];
        }
    }
    if (found) any = 1;
    __syncthreads();
    // status was preset to EMPTY_MASK by the host; any CTA that saw tissue clears it
    if (threadIdx.x == 0 && any && a.status) atomicAnd(&a.status[tile], ~SB_STATUS_EMPTY_MASK);
}

__global__ void __launch_bounds__(PT, 4) stain_augment_kernel(PointArgs a) {
    __shared__ PointShared sh;
    load_tables(&sh, a.tab, true);
    const int tile = blockIdx.x;
    if (threadIdx.x == 0) {
        double M[6];
        for (int k = 0; k < 6; ++k) M[k] = a.M[(size_t)tile * 6 + k];
        make_lasso_consts(M, a.lasso_lambda, sh.lk);
        const double LOG2E = 1.4426950408889634;
        for (int k = 0; k < 6; ++k) sh.A[k] = (float)(-M[k] * LOG2E);
        for (int j = 0; j < 2; ++j) { sh.alpha[j] = (float)a.scale[(size_t)tile * 2 + j]; sh.beta[j] = (float)a.beta[(size_t)tile * 2 + j]; }
    }
    __syncthreads();
    const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
    uint8_t* tout = a.out + (size_t)tile * a.npx * 3;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const LassoK lk = sh.lk;
    const float a00 = sh.A[0], a01 = sh.A[1], a02 = sh.A[2], a10 = sh.A[3], a11 = sh.A[4], a12 = sh.A[5];
    const float al0 = sh.alpha[0], al1 = sh.alpha[1], be0 = sh.beta[0], be1 = sh.beta[1];
    const float L255 = LOG2_255_UP;
    const bool all_px = a.augment_background != 0;
    const float ybound = a.ybound;
    const float* od = sh.od;
    const float* gyR = sh.gy, *gyG = sh.gy + 256, *gyB = sh.gy + 512;
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < G; g += gridDim.y * blockDim.x) {
        uint32_t w[12], o[12];
        int nvalid;
        load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
            const uint32_t rr[4] = {byte_of(wa, 0), byte_of(wa, 3), byte_of(wb, 2), byte_of(wc, 1)};
            const uint32_t gg[4] = {byte_of(wa, 1), byte_of(wb, 0), byte_of(wb, 3), byte_of(wc, 2)};
            const uint32_t bb[4] = {byte_of(wa, 2), byte_of(wb, 1), byte_of(wc, 0), byte_of(wc, 3)};
            uint32_t bits[12];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float c0, c1;
                lasso2(lk, od[rr[p]], od[gg[p]], od[bb[p]], c0, c1);
                const bool m = all_px | (gyR[rr[p]] + gyG[gg[p]] + gyB[bb[p]] < ybound);
                c0 = m ? fmaf(c0, al0, be0) : c0;
                c1 = m ? fmaf(c1, al1, be1) : c1;
                bits[3 * p] = clip_u8_bits(ex2_approx(fmaf(c1, a10, fmaf(c0, a00, L255))));
                bits[3 * p + 1] = clip_u8_bits(ex2_approx(fmaf(c1, a11, fmaf(c0, a01, L255))));
                bits[3 * p + 2] = clip_u8_bits(ex2_approx(fmaf(c1, a12, fmaf(c0, a02, L255))));
            }
            o[3 * q] = pack4(bits[0], bits[1], bits[2], bits[3]);
            o[3 * q + 1] = pack4(bits[4], bits[5], bits[6], bits[7]);
            o[3 * q + 2] = pack4(bits[8], bits[9], bits[10], bits[11]);
        }
        store_group(tout, a.npx, g, a.aligned != 0, o);
    }
}

__global__ void __launch_bounds__(PT) concentrations_kernel(PointArgs a) {
    __shared__ PointShared sh;
    load_tables(&sh, a.tab, false);
    const int tile = blockIdx.x;
    if (threadIdx.x == 0) {
        double M[6];
        for (int k = 0; k < 6; ++k) M[k] = a.M[(size_t)tile * 6 + k];
        make_lasso_consts(M, a.lasso_lambda, sh.lk);
    }
    __syncthreads();
    const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
    float2* cout = reinterpret_cast<float2*>(a.conc_out) + (size_t)tile * a.npx;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const LassoK lk = sh.lk;
    const float* od = sh.od;
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < G; g += gridDim.y * blockDim.x) {
        uint32_t w[12];
        int nvalid;
        load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
        for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
            float c0, c1;
            lasso2(lk, od[r], od[gg], od[b], c0, c1);
            if (i < nvalid) cout[(size_t)g * GROUP_PX + i] = make_float2(c0, c1);
        });
    }
}

static dim3 point_grid(const PointArgs& a, int num_sms, int ctas_per_sm) {
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    int spans = (G + PT - 1) / PT;
    // keep the grid near (SMs x resident CTAs) x a few waves when B is large; never more spans than work
    int want = (num_sms * ctas_per_sm * 4 + a.B - 1) / a.B;
    if (want < 1) want = 1;
    if (spans > want) spans = want;
    return dim3(a.B, spans);
}

__global__ void fill_i32_kernel(int32_t* p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int launch_mask(const PointArgs& a, int num_sms, cudaStream_t stream) {
    // preset every tile to EMPTY_MASK; a CTA that sees tissue clears the bit (stream-ordered: no host synchronisation)
    if (a.status) fill_i32_kernel<<<(a.B + 255) / 256, 256, 0, stream>>>(a.status, a.B, SB_STATUS_EMPTY_MASK);
    // 16-byte aligned tiles of whole 16-pixel groups: streaming pass on the TMA ring; register-staged kernel otherwise
    if (a.aligned && (a.npx % GROUP_PX) == 0 && ((size_t)a.npx * 3) % 16 == 0 && a.ycoef[0] != 0.f) return launch_mask_stream(a, num_sms, stream);
    mask_kernel<<<point_grid(a, num_sms, 8), PT, 0, stream>>>(a);
    return (int)cudaGetLastError();
}
// ---- StainAugmentor.pop on the TMA ring (sb_ring.cuh): the path for 16-byte aligned tiles
struct AugConsts {
    LassoK lk;
    float A[6];              // -M_jk log2(e): recombination with the tile's own stain matrix
    float alpha[2], beta[2];
};
struct AugRingParams {
    const AugConsts* consts;
    const float* od;
    const unsigned short* gamma;
    float ycoef[3], ybound;
    int all_px;
};
__global__ void aug_prepare_kernel(PointArgs a, AugConsts* out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.B) return;
    double M[6];
    for (int k = 0; k < 6; ++k) M[k] = a.M[(size_t)tile * 6 + k];
    AugConsts c;
    make_lasso_consts(M, a.lasso_lambda, c.lk);
    const double LOG2E = 1.4426950408889634;
    for (int k = 0; k < 6; ++k) c.A[k] = (float)(-M[k] * LOG2E);
    for (int j = 0; j < 2; ++j) { c.alpha[j] = (float)a.scale[(size_t)tile * 2 + j]; c.beta[j] = (float)a.beta[(size_t)tile * 2 + j]; }
    out[tile] = c;
}
struct AugOp {
    using Consts = AugConsts;
    using Params = AugRingParams;
    struct Acc {};
    static constexpr int kLaneShift = 3;      // {od, gamma} pairs
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float2*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 8) = make_float2(p.od[i >> 5], (float)p.gamma[i >> 5]);
    }
    __device__ static void acc_init(Acc&) {}
    __device__ static void pixel(const Consts& k, const Params& p, float2 r, float2 g, float2 b, uint32_t* bits) {
        float c0, c1;
        lasso2(k.lk, r.x, g.x, b.x, c0, c1);
        const bool m = p.all_px | (fmaf(p.ycoef[2], b.y, fmaf(p.ycoef[1], g.y, p.ycoef[0] * r.y)) < p.ybound);
        c0 = m ? fmaf(c0, k.alpha[0], k.beta[0]) : c0;
        c1 = m ? fmaf(c1, k.alpha[1], k.beta[1]) : c1;
        bits[0] = clip_u8_bits(ex2_approx(fmaf(c1, k.A[3], fmaf(c0, k.A[0], LOG2_255_UP))));
        bits[1] = clip_u8_bits(ex2_approx(fmaf(c1, k.A[4], fmaf(c0, k.A[1], LOG2_255_UP))));
        bits[2] = clip_u8_bits(ex2_approx(fmaf(c1, k.A[5], fmaf(c0, k.A[2], LOG2_255_UP))));
    }
    // Two pixels at a time on the packed f32x2 pipe, for unit-norm stain vectors (the compare-free LASSO of K4).
    // px = {{od, gamma} of r, g, b} of pixel a (index 0..2) and pixel b (3..5), already split into od[] / gm[].
    template <int LM>
    __device__ static void pair(const Consts& k, const Params& p, const float (&od)[6], const float (&gm)[6], uint32_t* bits) {
        float2 c0, c1;
        lasso2_unit_pair<LM>(k.lk, f2(od[0], od[3]), f2(od[1], od[4]), f2(od[2], od[5]), c0, c1);
        const float2 y = __ffma2_rn(dup(p.ycoef[2]), f2(gm[2], gm[5]), __ffma2_rn(dup(p.ycoef[1]), f2(gm[1], gm[4]), __fmul2_rn(dup(p.ycoef[0]), f2(gm[0], gm[3]))));
        const bool ma = p.all_px | (y.x < p.ybound), mb = p.all_px | (y.y < p.ybound);
        const float2 t0 = __ffma2_rn(c0, dup(k.alpha[0]), dup(k.beta[0])), t1 = __ffma2_rn(c1, dup(k.alpha[1]), dup(k.beta[1]));
        c0 = f2(ma ? t0.x : c0.x, mb ? t0.y : c0.y);
        c1 = f2(ma ? t1.x : c1.x, mb ? t1.y : c1.y);
        const float2 L = dup(LOG2_255_UP);
        const float2 e0 = __ffma2_rn(c1, dup(k.A[3]), __ffma2_rn(c0, dup(k.A[0]), L));
        const float2 e1 = __ffma2_rn(c1, dup(k.A[4]), __ffma2_rn(c0, dup(k.A[1]), L));
        const float2 e2 = __ffma2_rn(c1, dup(k.A[5]), __ffma2_rn(c0, dup(k.A[2]), L));
        bits[0] = clip_u8_bits(ex2_approx(e0.x)); bits[1] = clip_u8_bits(ex2_approx(e1.x)); bits[2] = clip_u8_bits(ex2_approx(e2.x));
        bits[3] = clip_u8_bits(ex2_approx(e0.y)); bits[4] = clip_u8_bits(ex2_approx(e1.y)); bits[5] = clip_u8_bits(ex2_approx(e2.y));
    }
    template <int LM>
    __device__ static void group(const Consts& k, const Params& p, const OdAbs tab, const uint32_t (&w)[12], uint32_t (&o)[12]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
            uint32_t bits[12];
            if (LM == LASSO_GENERAL) {
                pixel(k, p, odg_lookup_abs(tab, wa, 0), odg_lookup_abs(tab, wa, 1), odg_lookup_abs(tab, wa, 2), bits);
                pixel(k, p, odg_lookup_abs(tab, wa, 3), odg_lookup_abs(tab, wb, 0), odg_lookup_abs(tab, wb, 1), bits + 3);
                pixel(k, p, odg_lookup_abs(tab, wb, 2), odg_lookup_abs(tab, wb, 3), odg_lookup_abs(tab, wc, 0), bits + 6);
                pixel(k, p, odg_lookup_abs(tab, wc, 1), odg_lookup_abs(tab, wc, 2), odg_lookup_abs(tab, wc, 3), bits + 9);
            } else {
                // pixels: p0=(a0,a1,a2) p1=(a3,b0,b1) p2=(b2,b3,c0) p3=(c1,c2,c3)
                float od[6], gm[6];
                odg_lookup_split(tab, wa, 0, od[0], gm[0]); odg_lookup_split(tab, wa, 1, od[1], gm[1]); odg_lookup_split(tab, wa, 2, od[2], gm[2]);
                odg_lookup_split(tab, wa, 3, od[3], gm[3]); odg_lookup_split(tab, wb, 0, od[4], gm[4]); odg_lookup_split(tab, wb, 1, od[5], gm[5]);
                pair<LM>(k, p, od, gm, bits);
                odg_lookup_split(tab, wb, 2, od[0], gm[0]); odg_lookup_split(tab, wb, 3, od[1], gm[1]); odg_lookup_split(tab, wc, 0, od[2], gm[2]);
                odg_lookup_split(tab, wc, 1, od[3], gm[3]); odg_lookup_split(tab, wc, 2, od[4], gm[4]); odg_lookup_split(tab, wc, 3, od[5], gm[5]);
                pair<LM>(k, p, od, gm, bits + 6);
            }
            o[3 * q] = pack4(bits[0], bits[1], bits[2], bits[3]);
            o[3 * q + 1] = pack4(bits[4], bits[5], bits[6], bits[7]);
            o[3 * q + 2] = pack4(bits[8], bits[9], bits[10], bits[11]);
        }
    }
    using Run = int;                                                         // the tile's LASSO mode
    __device__ static Run begin_run(const Consts& k, const Params&) { return lasso_mode_of(k.lk.rg00, k.lk.rg11, k.lk.g01); }
    __device__ static void process(const Consts& k, const Params& p, const Run& lm, const OdAbs tab, uint4* grp, Acc&) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        uint32_t o[12];
        if (lm == LASSO_UNIT_POS) group<LASSO_UNIT_POS>(k, p, tab, w, o);
        else if (lm == LASSO_UNIT_NEG) group<LASSO_UNIT_NEG>(k, p, tab, w, o);
        else group<LASSO_GENERAL>(k, p, tab, w, o);
        grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
        grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
    __device__ static void finish_run(const Params&, int, Acc&) {}
};

int launch_stain_augment(const PointArgs& a, Scratch& scratch) {
    const int num_sms = scratch.h->num_sms;
    cudaStream_t stream = scratch.st;
    if (a.aligned) {
        AugConsts* consts = nullptr;
        cudaError_t e = scratch.get(&consts, (size_t)a.B * sizeof(AugConsts));
        if (e != cudaSuccess) return (int)e;
        aug_prepare_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a, consts);
        AugRingParams p{};
        p.consts = consts; p.od = a.tab.od; p.gamma = a.tab.gamma;
        p.ycoef[0] = a.ycoef[0]; p.ycoef[1] = a.ycoef[1]; p.ycoef[2] = a.ycoef[2]; p.ybound = a.ybound;
        p.all_px = a.augment_background != 0;
        return launch_ring<AugOp>(RingGeom{a.in, a.out, a.B, a.npx}, p, num_sms, stream);
    }
    stain_augment_kernel<<<point_grid(a, num_sms, 4), PT, 0, stream>>>(a);
    return (int)cudaGetLastError();
}
// ------------------------------------------------------------------------------------------- RGB <-> optical density
// convert_RGB_to_OD (stain_utils.py:101-112): OD = max(-ln(max(v, 1) / 255), 1e-6) is a function of one uint8, so the
// kernel is a 256-entry float64 table lookup: 1 B read, 8 B (or 4 B) written per value.
// One thread converts FOUR consecutive bytes (one 32-bit word in, 16 / 32 bytes out): the lanes of a warp read 128 contiguous
// bytes and write 512 / 1024 contiguous bytes per instruction.
template <typename T>
__global__ void __launch_bounds__(256) rgb_to_od_kernel(const uint8_t* __restrict__ in, T* __restrict__ out, size_t n, const double* __restrict__ od64) {
    __shared__ T tab[256];
    tab[threadIdx.x] = (T)od64[threadIdx.x];
    __syncthreads();
    const size_t nvec = n / 4;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % (4 * sizeof(T)) == 0);
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        uint32_t w;
        if (vec_ok) w = __ldg(reinterpret_cast<const uint32_t*>(in) + v);
        else w = (uint32_t)in[4 * v] | ((uint32_t)in[4 * v + 1] << 8) | ((uint32_t)in[4 * v + 2] << 16) | ((uint32_t)in[4 * v + 3] << 24);
        const T a = tab[w & 255u], b = tab[(w >> 8) & 255u], c = tab[(w >> 16) & 255u], d = tab[w >> 24];
        if (vec_ok) {
            if (sizeof(T) == 4) reinterpret_cast<float4*>(out)[v] = make_float4((float)a, (float)b, (float)c, (float)d);
            else { reinterpret_cast<double2*>(out)[2 * v] = make_double2((double)a, (double)b); reinterpret_cast<double2*>(out)[2 * v + 1] = make_double2((double)c, (double)d); }
        } else {
            out[4 * v] = a; out[4 * v + 1] = b; out[4 * v + 2] = c; out[4 * v + 3] = d;
        }
    }
    // ragged tail (< 4 values)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const size_t i = nvec * 4 + threadIdx.x; out[i] = tab[in[i]]; }
}
// convert_OD_to_RGB (stain_utils.py:114-124): uint8(255 * exp(-max(OD, 1e-6))), truncation toward zero; min_out[0] is
// lowered below zero when any OD is negative (the reference asserts OD.min() >= 0).  Four values per thread: one word out.
template <typename T>
__global__ void __launch_bounds__(256) od_to_rgb_kernel(const T* __restrict__ od, uint8_t* __restrict__ out, size_t n, int* __restrict__ negative) {
    int neg = 0;
    const size_t nvec = n / 4;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) % 4 == 0) && (reinterpret_cast<uintptr_t>(od) % (4 * sizeof(T)) == 0);
    auto conv = [&](T xv) -> uint32_t {
        const double x = (double)xv;
        neg |= (x < 0.0);
        return (uint32_t)(uint8_t)(int)(255.0 * exp(-fmax(x, 1e-6)));
    };
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        T x[4];
        if (vec_ok && sizeof(T) == 4) { const float4 q = reinterpret_cast<const float4*>(od)[v]; x[0] = (T)q.x; x[1] = (T)q.y; x[2] = (T)q.z; x[3] = (T)q.w; }
        else if (vec_ok) { const double2 q0 = reinterpret_cast<const double2*>(od)[2 * v], q1 = reinterpret_cast<const double2*>(od)[2 * v + 1]; x[0] = (T)q0.x; x[1] = (T)q0.y; x[2] = (T)q1.x; x[3] = (T)q1.y; }
        else { x[0] = od[4 * v]; x[1] = od[4 * v + 1]; x[2] = od[4 * v + 2]; x[3] = od[4 * v + 3]; }
        const uint32_t w = conv(x[0]) | (conv(x[1]) << 8) | (conv(x[2]) << 16) | (conv(x[3]) << 24);
        if (vec_ok) reinterpret_cast<uint32_t*>(out)[v] = w;
        else { out[4 * v] = (uint8_t)w; out[4 * v + 1] = (uint8_t)(w >> 8); out[4 * v + 2] = (uint8_t)(w >> 16); out[4 * v + 3] = (uint8_t)(w >> 24); }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const size_t i = nvec * 4 + threadIdx.x; out[i] = (uint8_t)conv(od[i]); }
    if (neg && negative) atomicOr(negative, 1);
}
int launch_rgb_to_od(const uint8_t* in, void* out, size_t n, int f32, const double* od64, int num_sms, cudaStream_t stream) {
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
    if (f32) rgb_to_od_kernel<float><<<(int)blocks, 256, 0, stream>>>(in, static_cast<float*>(out), n, od64);
    else rgb_to_od_kernel<double><<<(int)blocks, 256, 0, stream>>>(in, static_cast<double*>(out), n, od64);
    return (int)cudaGetLastError();
}
int launch_od_to_rgb(const void* od, uint8_t* out, size_t n, int f32, int* negative, int num_sms, cudaStream_t stream) {
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
    if (f32) od_to_rgb_kernel<float><<<(int)blocks, 256, 0, stream>>>(static_cast<const float*>(od), out, n, negative);
    else od_to_rgb_kernel<double><<<(int)blocks, 256, 0, stream>>>(static_cast<const double*>(od), out, n, negative);
    return (int)cudaGetLastError();
}

int launch_concentrations(const PointArgs& a, int num_sms, cudaStream_t stream) {
    concentrations_kernel<<<point_grid(a, num_sms, 8), PT, 0, stream>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace sb
