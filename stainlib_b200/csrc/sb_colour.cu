// sb_colour.cu -- LAB-space and HED-space kernels (sm_100a).
//
//   lab_tile_kernel   one CTA per tile, persistent over the batch; three uses:
//       REINHARD_STATS      ReinhardStainNormalizer.fit         normalizer.py:64-68
//       REINHARD_TRANSFORM  ReinhardStainNormalizer.transform   normalizer.py:70-94
//       LUM_STANDARDIZE     LuminosityStandardizer.standardize  stain_utils.py:53-67
//     All arithmetic on pixels is the integer 8-bit sRGB<->CIELAB path of OpenCV (oracle/cv_lab.py, verified on all
//     2^24 colours); every per-channel floating-point step of the reference is a function of one uint8, so it is
//     evaluated ONCE per tile in fp64 into 256-entry tables (brightness standardisation, the three Reinhard affine
//     maps, the luminosity rescale) -- the per-pixel work is table lookups and integer multiply-adds, bit-exact.
//     Percentiles of uint8-valued data come from exact 256-bin histograms.
//   hed_kernel        HedColorAugmenter.transform  augmenter.py:276-331  (single pass: speculative transform + patch-mean gate
//                     by the last CTA of each tile; skimage 0.17 colour deconvolution collapsed
//                     to  rgb' = b^(log_b(rgb/255+2) . A - c) - 2  with a per-tile 3x3 A and 3-vector c)
//   gray_kernel       GrayscaleAugmentor.pop       augmenter.py:390-401
#include "sb_kernels.h"
#include "sb_ring.cuh"
#include "sb_lab.cuh"
#include "sb_tables.inc"

namespace sb {

enum LabMode { REINHARD_STATS = 0, REINHARD_TRANSFORM = 1, LUM_STANDARDIZE = 2, BRIGHTNESS_STANDARDIZE = 3 };

struct LabArgs {
    const uint8_t* in;
    uint8_t* out;
    int B, npx, aligned, mode;
    Tables tab;
    const double* tmeans;   // [3] device (transform)
    const double* tstds;    // [3]
    double* means_out;      // [B,3] (stats)
    double* stds_out;       // [B,3]
    int mask_background;
    int lmax;               // tissue <=> L <= lmax (L of the standardised image)
    double percentile;      // LUM_STANDARDIZE
    int32_t* status;
    int skip_brightness;    // REINHARD_STATS without standardize_brightness first: get_mean_std (stain_utils.py:174-186)
};

struct __align__(16) LabShared {
    unsigned hist[3][256];
    unsigned short gamma[256];
    unsigned short gam2[256];       // gamma[smap[v]]: brightness standardisation folded into the linearisation table (per tile)
    unsigned short cbrt[3072];
    int yf[512];
    unsigned char invg[4096];
    unsigned char smap[256];        // brightness standardisation  v -> trunc(clip(v*255/p))
    unsigned char cmap[3][256];     // per-channel LAB maps
    double stat[6];                 // means (3), stds (3)
    double p;
    unsigned warp_any[NWARP];
    int any_tissue;
};

__device__ __forceinline__ void lab_forward(const LabShared* sh, int r, int g, int b, int& L, int& A, int& Bc) {
    // r, g, b are the RAW bytes: gam2 applies the tile's brightness standardisation and the sRGB linearisation at once
    const int R = sh->gam2[r], G = sh->gam2[g], Bl = sh->gam2[b];
    const int fX = sh->cbrt[(R * 1777 + G * 1541 + Bl * 778 + 2048) >> 12];
    const int fY = sh->cbrt[(R * 871 + G * 2929 + Bl * 296 + 2048) >> 12];
    const int fZ = sh->cbrt[(R * 73 + G * 448 + Bl * 3575 + 2048) >> 12];
    L = (SB_LAB_LSCALE * fY + SB_LAB_LSHIFT + 16384) >> 15;
    A = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    Bc = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    // no saturation needed: over all 2^24 colours L is in [0,255], a in [42,226], b in [20,223] (tests/test_oracle_lab.py)
}

__device__ __forceinline__ void lab_inverse(const LabShared* sh, int L, int A, int Bc, int& r, int& g, int& b) {
    const int y = sh->yf[2 * L], ify = sh->yf[2 * L + 1];
    const int adiv = ((5 * A * 53687 + 128) >> 13) - 128 * 16384 / 500;
    const int bdiv = ((Bc * 41943 + 16) >> 9) - 128 * 16384 / 200 + 1;
    const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
    int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
    int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
    int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
    r = sh->invg[min(max(ro, 0), 4095)];
    g = sh->invg[min(max(go, 0), 4095)];
    b = sh->invg[min(max(bo, 0), 4095)];
}

__global__ void __launch_bounds__(NT, 2) lab_tile_kernel(LabArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LabShared* sh = reinterpret_cast<LabShared*>(smem_raw);
    for (int i = threadIdx.x; i < 256; i += NT) sh->gamma[i] = a.tab.gamma[i];
    for (int i = threadIdx.x; i < 3072; i += NT) sh->cbrt[i] = a.tab.cbrt[i];
    for (int i = threadIdx.x; i < 512; i += NT) sh->yf[i] = a.tab.lab2yf[i];
    for (int i = threadIdx.x; i < 4096; i += NT) sh->invg[i] = a.tab.invgamma[i];
    __syncthreads();
    const int npx = a.npx;
    const int G = (npx + GROUP_PX - 1) / GROUP_PX;
    const bool aligned = a.aligned != 0;
    const bool reinhard = a.mode != LUM_STANDARDIZE;

    for (int tile = blockIdx.x; tile < a.B; tile += gridDim.x) {
        const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
        for (int i = threadIdx.x; i < 768; i += NT) (&sh->hist[0][0])[i] = 0;
        if (threadIdx.x == 0) sh->any_tissue = 0;
        __syncthreads();
        if (reinhard && !a.skip_brightness) {
            // ---- pass 1: histogram of all 3N channel bytes -> 90th percentile -> brightness table (stain_utils.py:188-194)
            for (int g = threadIdx.x; g < G; g += NT) {
                uint32_t w[12]; int nvalid;
                load_group<true>(tin, npx, g, aligned, w, nvalid);
                for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
                    if (i < nvalid) { atomicAdd(&sh->hist[0][r], 1u); atomicAdd(&sh->hist[0][gg], 1u); atomicAdd(&sh->hist[0][b], 1u); }
                });
            }
            __syncthreads();
            if (threadIdx.x == 0) sh->p = hist_percentile(sh->hist[0], 3ull * (unsigned long long)npx, 90.0);
            __syncthreads();
            if (threadIdx.x < 256) {
                const unsigned char m = trunc_clip_u8((double)threadIdx.x * 255.0 / sh->p);
                sh->smap[threadIdx.x] = m;
                sh->gam2[threadIdx.x] = sh->gamma[m];
                sh->hist[0][threadIdx.x] = 0;
            }
            __syncthreads();
            if (a.mode == BRIGHTNESS_STANDARDIZE) {
                // standardize_brightness alone (stain_utils.py:188-194): every byte through the table, nothing else
                uint8_t* __restrict__ tb = a.out + (size_t)tile * npx * 3;
                for (int g = threadIdx.x; g < G; g += NT) {
                    uint32_t w[12], o[12]; int nvalid;
                    load_group<true>(tin, npx, g, aligned, w, nvalid);
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        o[k] = (uint32_t)sh->smap[w[k] & 255u] | ((uint32_t)sh->smap[(w[k] >> 8) & 255u] << 8) |
                               ((uint32_t)sh->smap[(w[k] >> 16) & 255u] << 16) | ((uint32_t)sh->smap[w[k] >> 24] << 24);
                    store_group(tb, npx, g, aligned, o);
                }
                __syncthreads();
                continue;
            }
        } else if (tile == (int)blockIdx.x) {
            if (threadIdx.x < 256) sh->gam2[threadIdx.x] = sh->gamma[threadIdx.x];
            __syncthreads();
        }
        // ---- pass 2: LAB histograms of the (standardised) tile.  When a result tile will be written, its memory serves as
        // scratch for the LAB bytes, so that pass 3 does not repeat the forward conversion (the costliest step per pixel)
        uint8_t* __restrict__ tout = a.out ? a.out + (size_t)tile * npx * 3 : nullptr;
        const bool keep_lab = tout != nullptr && a.mode != REINHARD_STATS;
        for (int g = threadIdx.x; g < G; g += NT) {
            uint32_t w[12]; int nvalid;
            load_group<true>(tin, npx, g, aligned, w, nvalid);
            uint32_t lab[48];
            for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
                int L, A, Bc;
                lab_forward(sh, r, gg, b, L, A, Bc);
                lab[3 * i] = L; lab[3 * i + 1] = A; lab[3 * i + 2] = Bc;
                if (i < nvalid) {
                    atomicAdd(&sh->hist[0][L], 1u);
                    if (reinhard) { atomicAdd(&sh->hist[1][A], 1u); atomicAdd(&sh->hist[2][Bc], 1u); }
                }
            });
            if (keep_lab) {
                uint32_t o[12];
#pragma unroll
                for (int k = 0; k < 12; ++k) o[k] = lab[4 * k] | (lab[4 * k + 1] << 8) | (lab[4 * k + 2] << 16) | (lab[4 * k + 3] << 24);
                store_group_keep(tout, npx, g, aligned, o);
            }
        }
        __syncthreads();
        if (reinhard) {
            // mean and population std per channel exactly as cv.meanStdDev sees them (stain_utils.py:146-186):
            // I1 = float32(L)/float32(2.55), I2 = a-128, I3 = b-128, double accumulators.
            if (threadIdx.x < 3) {
                const int c = threadIdx.x;
                double s = 0.0, sq = 0.0;
                for (int v = 0; v < 256; ++v) {
                    const double q = c == 0 ? (double)((float)v / 2.55f) : (double)(v - 128);
                    const double h = (double)sh->hist[c][v];
                    s += h * q; sq += h * q * q;
                }
                const double mean = s / (double)npx;
                double var = sq / (double)npx - mean * mean;
                if (var < 0.0) var = 0.0;
                sh->stat[c] = mean; sh->stat[3 + c] = sqrt(var);
            }
            __syncthreads();
            if (a.mode == REINHARD_STATS) {
                if (threadIdx.x < 3) {
                    a.means_out[(size_t)tile * 3 + threadIdx.x] = sh->stat[threadIdx.x];
                    a.stds_out[(size_t)tile * 3 + threadIdx.x] = sh->stat[3 + threadIdx.x];
                }
                __syncthreads();
                continue;
            }
            // per-channel affine maps (normalizer.py:81-83) followed by merge_back's scale/offset, clip and truncation
            for (int i = threadIdx.x; i < 768; i += NT) {
                const int c = i >> 8, v = i & 255;
                const double q = c == 0 ? (double)((float)v / 2.55f) : (double)(v - 128);
                const double n = (q - sh->stat[c]) * (a.tstds[c] / sh->stat[3 + c]) + a.tmeans[c];
                sh->cmap[c][v] = trunc_clip_u8(c == 0 ? n * 2.55 : n + 128.0);
            }
        } else {
            if (threadIdx.x == 0) sh->p = hist_percentile(sh->hist[0], (unsigned long long)npx, a.percentile);
            __syncthreads();
            if (threadIdx.x < 256) sh->cmap[0][threadIdx.x] = trunc_clip_u8(255.0 * (double)threadIdx.x / sh->p);
        }
        __syncthreads();
        // ---- pass 3: map and convert back; reads the LAB bytes this thread stored in pass 2 and overwrites them with RGB
        const bool use_mask = reinhard && a.mask_background;
        const int lmax = a.lmax;
        int seen = 0;
        for (int g = threadIdx.x; g < G; g += NT) {
            uint32_t w[12], o[12]; int nvalid;
            load_group_rw(tout, npx, g, aligned, w, nvalid);
            uint32_t ob[48];
            for_each_px(w, [&](int i, uint32_t Lu, uint32_t Au, uint32_t Bu) {
                const int L = (int)Lu, A = (int)Au, Bc = (int)Bu;
                int L2, A2, B2;
                if (reinhard) {
                    const bool tissue = !use_mask || L <= lmax;
                    seen |= (tissue && i < nvalid) ? 1 : 0;
                    // background: L = clip((254 + 0) * 2.55) = 255, a = b = 0 + 128 (normalizer.py:85-90)
                    L2 = tissue ? sh->cmap[0][L] : 255; A2 = tissue ? sh->cmap[1][A] : 128; B2 = tissue ? sh->cmap[2][Bc] : 128;
                } else {
                    L2 = sh->cmap[0][L]; A2 = A; B2 = Bc;
                }
                int ro, go, bo;
                lab_inverse(sh, L2, A2, B2, ro, go, bo);
                ob[3 * i] = ro; ob[3 * i + 1] = go; ob[3 * i + 2] = bo;
            });
#pragma unroll
            for (int k = 0; k < 12; ++k) o[k] = ob[4 * k] | (ob[4 * k + 1] << 8) | (ob[4 * k + 2] << 16) | (ob[4 * k + 3] << 24);
            store_group(tout, npx, g, aligned, o);
        }
        if (use_mask) {
            if (seen) sh->any_tissue = 1;
            __syncthreads();
            if (threadIdx.x == 0 && a.status) a.status[tile] = sh->any_tissue ? 0 : SB_STATUS_EMPTY_MASK;
        } else if (threadIdx.x == 0 && a.status) {
            a.status[tile] = 0;
        }
        __syncthreads();
    }
}

// ---- lab_split / merge_back (stain_utils.py:146-172) as stand-alone per-pixel kernels: the exported pieces of the Reinhard
// path.  The integer sRGB <-> CIELAB arithmetic is lab_forward / lab_inverse above with the plain linearisation table.
struct LabPointTables {
    unsigned short gamma[256];
    unsigned short cbrt[3072];
    int yf[512];
    unsigned char invg[4096];
};
__device__ __forceinline__ void lab_point_load(LabPointTables* t, const Tables& tab) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) t->gamma[i] = tab.gamma[i];
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) t->cbrt[i] = tab.cbrt[i];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) t->yf[i] = tab.lab2yf[i];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) t->invg[i] = tab.invgamma[i];
    __syncthreads();
}
// I1 = float32(L) / 2.55, I2 = a - 128, I3 = b - 128 (float32 arithmetic, as numpy does on the float32 planes).
__global__ void __launch_bounds__(256) lab_split_kernel(Tables tab, const uint8_t* __restrict__ rgb, size_t npx, float* __restrict__ I1,
                                                        float* __restrict__ I2, float* __restrict__ I3) {
    __shared__ LabPointTables t;
    lab_point_load(&t, tab);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (size_t)gridDim.x * blockDim.x) {
        const int R = t.gamma[rgb[3 * p]], G = t.gamma[rgb[3 * p + 1]], Bl = t.gamma[rgb[3 * p + 2]];
        const int fX = t.cbrt[(R * 1777 + G * 1541 + Bl * 778 + 2048) >> 12];
        const int fY = t.cbrt[(R * 871 + G * 2929 + Bl * 296 + 2048) >> 12];
        const int fZ = t.cbrt[(R * 73 + G * 448 + Bl * 3575 + 2048) >> 12];
        const int L = (SB_LAB_LSCALE * fY + SB_LAB_LSHIFT + 16384) >> 15;
        const int A = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
        const int Bc = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
        I1[p] = __fdiv_rn((float)L, 2.55f);
        I2[p] = (float)A - 128.0f;
        I3[p] = (float)Bc - 128.0f;
    }
}
// uint8(clip((I1 * 2.55, I2 + 128, I3 + 128), 0, 255)) -> LAB2RGB, in the arithmetic of the planes' own dtype.
template <typename T>
__global__ void __launch_bounds__(256) lab_merge_kernel(Tables tab, const T* __restrict__ I1, const T* __restrict__ I2, const T* __restrict__ I3,
                                                        size_t npx, uint8_t* __restrict__ rgb) {
    __shared__ LabPointTables t;
    lab_point_load(&t, tab);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += (size_t)gridDim.x * blockDim.x) {
        const int L = trunc_clip_u8((double)(I1[p] * (T)2.55)), A = trunc_clip_u8((double)(I2[p] + (T)128.0)), Bc = trunc_clip_u8((double)(I3[p] + (T)128.0));
        const int y = t.yf[2 * L], ify = t.yf[2 * L + 1];
        const int adiv = ((5 * A * 53687 + 128) >> 13) - 128 * 16384 / 500;
        const int bdiv = ((Bc * 41943 + 16) >> 9) - 128 * 16384 / 200 + 1;
        const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
        const int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
        const int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
        const int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
        rgb[3 * p] = t.invg[min(max(ro, 0), 4095)];
        rgb[3 * p + 1] = t.invg[min(max(go, 0), 4095)];
        rgb[3 * p + 2] = t.invg[min(max(bo, 0), 4095)];
    }
}

// ------------------------------------------------------------------------------------------------------------ HED
// Two definitions of skimage.color.rgb2hed / hed2rgb exist (call sites augmenter.py:295,319):
//   variant 17 (scikit-image 0.16-0.17, pinned by the reference's environment.yml:107; base e before 0.16):
//       hed = -log_b(rgb + 2) @ inv(M);  rgb' = clip(b^(-hed' @ M) - 2, -1, 1)                  -> affine in log space:
//       log2(rgb' + 2) = log2(rgb + 2) @ A - c,  A = inv(M) diag(1+sigma) M,  c = (bias @ M) log2(b)
//   variant 18 (scikit-image >= 0.18, what an unpinned `pip install scikit-image` gives today -- setup.py:11-18):
//       hed = max(0, (ln(max(rgb, 1e-6)) / ln(1e-6)) @ inv(M));  rgb' = clip(exp(ln(1e-6) * (hed' @ M)), 0, 1)
//       -> L = ln(max(rgb, 1e-6)) / ln(1e-6) (table), s = max(0, L @ inv(M)), log2(rgb') = s @ Q + r with
//          Q = ln(1e-6) log2(e) diag(1+sigma) M,  r = ln(1e-6) log2(e) (bias @ M).
// M = rgb_from_hed.  In both cases the result is clipped to [0,1], scaled by 255 and truncated (augmenter.py:320-325).
struct HedMats { double M[9], Mi[9]; };
__host__ __device__ inline HedMats hed_mats() {
    HedMats h;
    const double M[9] = {0.65, 0.70, 0.29, 0.07, 0.99, 0.11, 0.27, 0.57, 0.78};
    for (int i = 0; i < 9; ++i) h.M[i] = M[i];
    const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
    h.Mi[0] = (M[4] * M[8] - M[5] * M[7]) / det; h.Mi[1] = (M[2] * M[7] - M[1] * M[8]) / det; h.Mi[2] = (M[1] * M[5] - M[2] * M[4]) / det;
    h.Mi[3] = (M[5] * M[6] - M[3] * M[8]) / det; h.Mi[4] = (M[0] * M[8] - M[2] * M[6]) / det; h.Mi[5] = (M[2] * M[3] - M[0] * M[5]) / det;
    h.Mi[6] = (M[3] * M[7] - M[4] * M[6]) / det; h.Mi[7] = (M[1] * M[6] - M[0] * M[7]) / det; h.Mi[8] = (M[0] * M[4] - M[1] * M[3]) / det;
    return h;
}
constexpr double HED18_LN = -13.815510557964274;          // ln(1e-6)
constexpr double HED_LOG2E = 1.4426950408889634;
// Per-tile constants of either variant: A[9] (row-major, input channel x output channel) and the additive c[3], such that
// the exponent of output channel j is  c[j] + sum_i x_i A[3 i + j]  (x = table values for 17, clamped stains for 18).
struct HedConsts { float A[9]; float c[3]; };
__host__ __device__ inline void hed_consts(int variant, const double* sigma, const double* bias, double log_base, HedConsts& k) {
    const HedMats h = hed_mats();
    if (variant == 18) {
        const double f = HED18_LN * HED_LOG2E;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) k.A[3 * i + j] = (float)(f * (1.0 + sigma[i]) * h.M[3 * i + j]);
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int q = 0; q < 3; ++q) v += bias[q] * h.M[3 * q + j];
            k.c[j] = (float)(f * v);
        }
        return;
    }
    const double l2b = log2(log_base);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double v = 0.0;
            for (int q = 0; q < 3; ++q) v += h.Mi[3 * i + q] * (1.0 + sigma[q]) * h.M[3 * q + j];
            k.A[3 * i + j] = (float)v;
        }
    for (int j = 0; j < 3; ++j) {
        double v = 0.0;
        for (int q = 0; q < 3; ++q) v += bias[q] * h.M[3 * q + j];
        k.c[j] = (float)(-v * l2b);
    }
}
// Table value of a uint8 channel for either variant.
__device__ __forceinline__ float hed_table_value(int variant, int v) {
    if (variant == 18) return (float)(log(fmax((double)v / 255.0, 1e-6)) / HED18_LN);
    return (float)log2((double)v / 255.0 + 2.0);
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float d; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
// Two pixels on the packed f32x2 pipe: table values of the three channels -> six output bytes (low byte of each word).
template <int V>
__device__ __forceinline__ void hed_pair(const HedConsts& k, const float (&mi)[9], float2 lr, float2 lg, float2 lb, uint32_t (&bits)[6]) {
    if (V == 18) {
        // stains = max(0, L @ inv(M))
        float2 s0 = __ffma2_rn(lb, dup(mi[6]), __ffma2_rn(lg, dup(mi[3]), __fmul2_rn(lr, dup(mi[0]))));
        float2 s1 = __ffma2_rn(lb, dup(mi[7]), __ffma2_rn(lg, dup(mi[4]), __fmul2_rn(lr, dup(mi[1]))));
        float2 s2 = __ffma2_rn(lb, dup(mi[8]), __ffma2_rn(lg, dup(mi[5]), __fmul2_rn(lr, dup(mi[2]))));
        s0 = f2(fmaxf(s0.x, 0.f), fmaxf(s0.y, 0.f)); s1 = f2(fmaxf(s1.x, 0.f), fmaxf(s1.y, 0.f)); s2 = f2(fmaxf(s2.x, 0.f), fmaxf(s2.y, 0.f));
        lr = s0; lg = s1; lb = s2;
    }
    const float2 e0 = __ffma2_rn(lb, dup(k.A[6]), __ffma2_rn(lg, dup(k.A[3]), __ffma2_rn(lr, dup(k.A[0]), dup(k.c[0]))));
    const float2 e1 = __ffma2_rn(lb, dup(k.A[7]), __ffma2_rn(lg, dup(k.A[4]), __ffma2_rn(lr, dup(k.A[1]), dup(k.c[1]))));
    const float2 e2 = __ffma2_rn(lb, dup(k.A[8]), __ffma2_rn(lg, dup(k.A[5]), __ffma2_rn(lr, dup(k.A[2]), dup(k.c[2]))));
    const float off = V == 18 ? 0.f : -2.f;      // clip(2^e + off, 0, 1) * 255, truncated: one saturating FMA + one round-down FMA
    bits[0] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e0.x), 1.f, off), 255.f, 8388608.f));
    bits[1] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e1.x), 1.f, off), 255.f, 8388608.f));
    bits[2] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e2.x), 1.f, off), 255.f, 8388608.f));
    bits[3] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e0.y), 1.f, off), 255.f, 8388608.f));
    bits[4] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e1.y), 1.f, off), 255.f, 8388608.f));
    bits[5] = __float_as_uint(__fmaf_rd(fma_sat(ex2_approx(e2.y), 1.f, off), 255.f, 8388608.f));
}
// The 16 pixels of a group (12 packed words) through a lane-replicated table.
template <int V, class TAB>
__device__ __forceinline__ void hed_words(const HedConsts& k, const float (&mi)[9], const TAB tab, uint32_t lane_off, const uint32_t (&w)[12], uint32_t (&o)[12]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
        // pixels: p0=(a0,a1,a2) p1=(a3,b0,b1) p2=(b2,b3,c0) p3=(c1,c2,c3)
        uint32_t b01[6], b23[6];
        hed_pair<V>(k, mi, f2(od_lookup(tab, wa, lane_off, 0), od_lookup(tab, wa, lane_off, 3)), f2(od_lookup(tab, wa, lane_off, 1), od_lookup(tab, wb, lane_off, 0)),
                    f2(od_lookup(tab, wa, lane_off, 2), od_lookup(tab, wb, lane_off, 1)), b01);
        hed_pair<V>(k, mi, f2(od_lookup(tab, wb, lane_off, 2), od_lookup(tab, wc, lane_off, 1)), f2(od_lookup(tab, wb, lane_off, 3), od_lookup(tab, wc, lane_off, 2)),
                    f2(od_lookup(tab, wc, lane_off, 0), od_lookup(tab, wc, lane_off, 3)), b23);
        o[3 * q] = pack4(b01[0], b01[1], b01[2], b01[3]);
        o[3 * q + 1] = pack4(b01[4], b01[5], b23[0], b23[1]);
        o[3 * q + 2] = pack4(b23[2], b23[3], b23[4], b23[5]);
    }
}
__device__ __forceinline__ void hed_load_mi(float (&mi)[9]) {
    const HedMats h = hed_mats();
#pragma unroll
    for (int i = 0; i < 9; ++i) mi[i] = (float)h.Mi[i];
}

struct HedArgs {
    const uint8_t* in;
    uint8_t* out;
    int B, npx, aligned;
    const double* sigma;   // [B,3]
    const double* bias;    // [B,3]
    double cutoff_lo, cutoff_hi, log_base;
    int variant;           // 17: scikit-image 0.16-0.17 (log_base, "+2" offset); 18: scikit-image >= 0.18 (see Hed18)
    int32_t* status;       // 1 = tile outside the cutoff (copied through)
    unsigned long long* sums;   // [B] workspace (zeroed): sum of all channel bytes
    unsigned* done;             // [B] workspace (zeroed): CTAs of the tile that have finished
};

template <int V>
__global__ void __launch_bounds__(256, 3) hed_kernel(HedArgs a) {
    // table value of every uint8, replicated per lane with 256-byte rows like the OD table of K4: one PRMT builds the offset
    // from the packed pixel word, the lookup is bank-conflict free
    extern __shared__ __align__(256) unsigned char l2_rep[];
    __shared__ float l2[256];
    __shared__ HedConsts kc;
    __shared__ int skip;
    __shared__ unsigned long long wsum[8];
    const int tile = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) l2[i] = hed_table_value(V, i);
    __syncthreads();
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x)
        *reinterpret_cast<float*>(l2_rep + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = l2[i >> 5];
    if (threadIdx.x == 0) hed_consts(V, a.sigma + (size_t)tile * 3, a.bias + (size_t)tile * 3, a.log_base, kc);
    __syncthreads();
    const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
    uint8_t* tout = a.out + (size_t)tile * a.npx * 3;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const HedConsts k = kc;
    float mi[9];
    hed_load_mi(mi);
    const uint32_t lane_off = (threadIdx.x & 31) << 2;
    // The patch-mean gate (augmenter.py:288-293) needs the sum of ALL bytes of the tile -- a global dependency that would
    // cost a separate read pass.  Tiles outside the cutoff are rare, so every CTA transforms its share speculatively while
    // it sums its bytes; the last CTA of a tile to finish evaluates the gate and, if the tile is to be left unchanged,
    // copies the input over the speculative output.
    unsigned long long acc = 0;
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < G; g += gridDim.y * blockDim.x) {
        uint32_t w[12], o[12]; int nvalid;
        load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
        {
            unsigned sg = 0;
            if (nvalid == GROUP_PX) {
#pragma unroll
                for (int k = 0; k < 12; ++k) sg += __vsadu4(w[k], 0u);      // sum of the four bytes
            } else {
                for (int k = 0; k < nvalid * 3; ++k) sg += (w[k >> 2] >> (8 * (k & 3))) & 255u;
            }
            acc += sg;
        }
        hed_words<V>(k, mi, (const unsigned char*)l2_rep, lane_off, w, o);
        store_group(tout, a.npx, g, a.aligned != 0, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();                                   // also orders this CTA's output stores before the ticket below
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += wsum[wv];
        atomicAdd(&a.sums[tile], t);
        __threadfence();
        const unsigned ticket = atomicAdd(&a.done[tile], 1u);
        int sk = 0;
        if (ticket == gridDim.y - 1) {
            __threadfence();
            const unsigned long long total = atomicAdd(&a.sums[tile], 0ull);
            // patch mean in [0,1]; exact integer sum instead of numpy's float32 pairwise mean
            const double mean = (double)total / (3.0 * (double)a.npx) / 255.0;
            sk = !(a.cutoff_lo <= mean && mean <= a.cutoff_hi);
            if (a.status) a.status[tile] = sk;
        }
        skip = sk;
    }
    __syncthreads();
    if (skip) {                                        // rare: the whole tile is returned unchanged
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            uint32_t w[12]; int nvalid;
            load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
            store_group(tout, a.npx, g, a.aligned != 0, w);
        }
    }
}

// ---- HED on the TMA ring (sb_ring.cuh): the path for 16-byte aligned tiles; hed_kernel above serves the rest.
struct HedRingParams {
    const HedConsts* consts;       // [B]
    unsigned long long* sums;      // [B] zeroed: sum of all channel bytes of the tile (patch-mean gate)
};

// Per-tile constants (hed_consts) for the ring operator.
__global__ void hed_prepare_kernel(int B, const double* __restrict__ sigma, const double* __restrict__ bias, double log_base, int variant, HedConsts* out) {
    const int tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= B) return;
    HedConsts k;
    hed_consts(variant, sigma + (size_t)tile * 3, bias + (size_t)tile * 3, log_base, k);
    out[tile] = k;
}

template <int V>
struct HedOpT {
    using Consts = HedConsts;
    using Params = HedRingParams;
    using Acc = unsigned;
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params&, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = hed_table_value(V, i >> 5);
    }
    __device__ static void acc_init(Acc& a) { a = 0u; }
    struct Run { float mi[9]; };
    __device__ static Run begin_run(const Consts&, const Params&) { Run r; hed_load_mi(r.mi); return r; }
    __device__ static void process(const Consts& k, const Params&, const Run& run, const OdAbs tab, uint4* grp, Acc& acc) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        uint32_t o[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) acc += __vsadu4(w[i], 0u);           // the speculative transform also sums the input bytes
        hed_words<V>(k, run.mi, tab, 0u, w, o);
        grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
        grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
    __device__ static void finish_run(const Params& p, int tile, Acc& acc) {
        unsigned long long t = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&p.sums[tile], t);
        acc = 0u;
    }
};
using HedOp = HedOpT<17>;
using Hed18Op = HedOpT<18>;

// ---- float patches (augmenter.py:288-291, 323-327): the reference takes float images in [0,1] as they are and returns
// floats.  Not a hot path: one pass for the patch mean (fixed-point sum: deterministic), one pass for the transform with
// accurate log2f / exp2f (the uint8 path's lookup table does not apply).
__global__ void __launch_bounds__(256) hed_float_sum_kernel(const float* __restrict__ in, size_t n_per_tile, long long* __restrict__ sums) {
    const int tile = blockIdx.x;
    const float* t = in + (size_t)tile * n_per_tile;
    long long acc = 0;
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n_per_tile; i += (size_t)gridDim.y * blockDim.x)
        acc += __double2ll_rn((double)t[i] * 1099511627776.0);      // 2^40
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(reinterpret_cast<unsigned long long*>(&sums[tile]), (unsigned long long)acc);
}
template <int V>
__global__ void __launch_bounds__(256) hed_float_kernel(const float* __restrict__ in, float* __restrict__ out, int npx, const double* __restrict__ sigma,
                                                         const double* __restrict__ bias, double lo, double hi, double log_base,
                                                         const long long* __restrict__ sums, int32_t* status) {
    const int tile = blockIdx.x;
    __shared__ HedConsts kc;
    if (threadIdx.x == 0) hed_consts(V, sigma + (size_t)tile * 3, bias + (size_t)tile * 3, log_base, kc);
    __syncthreads();
    const double mean = (double)sums[tile] / 1099511627776.0 / (3.0 * (double)npx);
    const bool skip = !(lo <= mean && mean <= hi);
    if (threadIdx.x == 0 && blockIdx.y == 0 && status) status[tile] = skip ? 1 : 0;
    const float* tin = in + (size_t)tile * npx * 3;
    float* tout = out + (size_t)tile * npx * 3;
    const HedConsts k = kc;
    float mi[9];
    hed_load_mi(mi);
    for (int p = blockIdx.y * blockDim.x + threadIdx.x; p < npx; p += gridDim.y * blockDim.x) {
        const float r = tin[3 * p], g = tin[3 * p + 1], b = tin[3 * p + 2];
        if (skip) { tout[3 * p] = r; tout[3 * p + 1] = g; tout[3 * p + 2] = b; continue; }
        float x0, x1, x2;
        if (V == 18) {
            const float inv = (float)(1.0 / (HED18_LN * HED_LOG2E));        // ln(x)/ln(1e-6) = log2(x) / (ln(1e-6) log2 e)
            const float l0 = log2f(fmaxf(r, 1e-6f)) * inv, l1 = log2f(fmaxf(g, 1e-6f)) * inv, l2 = log2f(fmaxf(b, 1e-6f)) * inv;
            x0 = fmaxf(fmaf(l2, mi[6], fmaf(l1, mi[3], l0 * mi[0])), 0.f);
            x1 = fmaxf(fmaf(l2, mi[7], fmaf(l1, mi[4], l0 * mi[1])), 0.f);
            x2 = fmaxf(fmaf(l2, mi[8], fmaf(l1, mi[5], l0 * mi[2])), 0.f);
        } else {
            x0 = log2f(r + 2.f); x1 = log2f(g + 2.f); x2 = log2f(b + 2.f);
        }
        const float off = V == 18 ? 0.f : -2.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float e = fmaf(x2, k.A[6 + j], fmaf(x1, k.A[3 + j], fmaf(x0, k.A[j], k.c[j])));
            tout[3 * p + j] = fminf(fmaxf(exp2f(e) + off, 0.f), 1.f);
        }
    }
}

// Runs behind the ring kernel: evaluates the patch-mean gate of every tile (augmenter.py:288-293) and copies the input
// back over the (rare) tiles that are to be left unchanged.  One CTA per tile; 16-byte aligned tiles.
__global__ void __launch_bounds__(256) hed_gate_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int npx,
                                                        const unsigned long long* __restrict__ sums, double lo, double hi, int32_t* status) {
    const int tile = blockIdx.x;
    const double mean = (double)sums[tile] / (3.0 * (double)npx) / 255.0;
    const bool skip = !(lo <= mean && mean <= hi);
    if (threadIdx.x == 0 && status) status[tile] = skip ? 1 : 0;
    if (!skip) return;
    const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)tile * npx * 3);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)tile * npx * 3);
    const size_t nvec = (size_t)npx * 3 / 16;
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = src[i];
}

struct GrayArgs {
    const uint8_t* in;
    uint8_t* out;
    int B, npx, aligned;
    const double* alpha;   // [B]
    const double* beta;    // [B]
};

__global__ void __launch_bounds__(256, 4) gray_kernel(GrayArgs a) {
    const int tile = blockIdx.x;
    const float al = (float)a.alpha[tile], be = (float)a.beta[tile];
    const uint8_t* tin = a.in + (size_t)tile * a.npx * 3;
    uint8_t* tout = a.out + (size_t)tile * a.npx * 3;
    const int G = (a.npx + GROUP_PX - 1) / GROUP_PX;
    const float kr = (float)(0.2125 / 255.0), kg = (float)(0.7154 / 255.0), kb = (float)(0.0721 / 255.0);
    for (int g = blockIdx.y * blockDim.x + threadIdx.x; g < G; g += gridDim.y * blockDim.x) {
        uint32_t w[12], o[12]; int nvalid;
        load_group<false>(tin, a.npx, g, a.aligned != 0, w, nvalid);
        uint32_t v[16];
        for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
            const float gray = fmaf((float)b, kb, fmaf((float)gg, kg, (float)r * kr));
            const float x = fminf(fmaxf(fmaf(gray, al, be), 0.f), 1.f) * 255.f;
            v[i] = clip_u8_bits(x) & 255u;
        });
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t p0 = v[4 * q], p1 = v[4 * q + 1], p2 = v[4 * q + 2], p3 = v[4 * q + 3];
            o[3 * q] = p0 | (p0 << 8) | (p0 << 16) | (p1 << 24);
            o[3 * q + 1] = p1 | (p1 << 8) | (p2 << 16) | (p2 << 24);
            o[3 * q + 2] = p2 | (p3 << 8) | (p3 << 16) | (p3 << 24);
        }
        store_group(tout, a.npx, g, a.aligned != 0, o);
    }
}

// ---- GrayscaleAugmentor.pop on the TMA ring
struct GrayConsts { float al, be; };
struct GrayConstsView {                      // consts[tile] builds the pair from the caller's double arrays
    const double* alpha;
    const double* beta;
    __device__ GrayConsts operator[](int tile) const { return GrayConsts{(float)alpha[tile], (float)beta[tile]}; }
};
struct GrayRingParams { GrayConstsView consts; };
struct GrayOp {
    using Consts = GrayConsts;
    using Params = GrayRingParams;
    struct Acc {};
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char*, const Params&, int, int) {}
    __device__ static void acc_init(Acc&) {}
    struct Run {};
    __device__ static Run begin_run(const Consts&, const Params&) { return Run{}; }
    __device__ static void process(const Consts& k, const Params&, const Run&, const OdAbs, uint4* grp, Acc&) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        const float kr = (float)(0.2125 / 255.0), kg = (float)(0.7154 / 255.0), kb = (float)(0.0721 / 255.0);
        uint32_t v[16], o[12];
        for_each_px(w, [&](int i, uint32_t r, uint32_t gg, uint32_t b) {
            const float gray = fmaf((float)b, kb, fmaf((float)gg, kg, (float)r * kr));
            const float x = fminf(fmaxf(fmaf(gray, k.al, k.be), 0.f), 1.f) * 255.f;
            v[i] = clip_u8_bits(x) & 255u;
        });
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t p0 = v[4 * q], p1 = v[4 * q + 1], p2 = v[4 * q + 2], p3 = v[4 * q + 3];
            o[3 * q] = p0 | (p0 << 8) | (p0 << 16) | (p1 << 24);
            o[3 * q + 1] = p1 | (p1 << 8) | (p2 << 16) | (p2 << 24);
            o[3 * q + 2] = p2 | (p3 << 8) | (p3 << 16) | (p3 << 24);
        }
        grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
        grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
    __device__ static void finish_run(const Params&, int, Acc&) {}
};

}  // namespace sb

// ------------------------------------------------------------------------------------------------------ C ABI glue

namespace {
int lab_lmax(double thr) {
    int best = -1;
    for (int L = 0; L < 256; ++L)
        if ((double)L / 255.0 < thr) best = L;
    return best;
}
// Diagnostic switch (A/B timing, path-equality tests): SB_REINHARD_TILE_KERNEL=1 keeps aligned tiles on lab_tile_kernel.
bool ring_disabled() {
    const char* e = getenv("SB_REINHARD_TILE_KERNEL");
    return e && e[0] == '1';
}
int launch_lab(sb_handle* hh, sb::LabArgs& a, cudaStream_t st) {
    sb_handle* h = hh;
    // 16-byte aligned tiles: the streaming passes of sb_reinhard.cu (same bytes; lab_tile_kernel keeps the rest)
    if (a.mode != sb::BRIGHTNESS_STANDARDIZE && a.aligned && sb::reinhard_ring_eligible(a.in, a.out ? a.out : a.in, a.npx) && !ring_disabled())
        return sb::launch_reinhard_ring(h, a.in, a.out, a.B, a.npx, a.mode, a.skip_brightness, a.tmeans, a.tstds, a.means_out, a.stds_out,
                                        a.mask_background, a.lmax, a.percentile, a.status, st);
    a.tab = h->tab;
    static sb::DeviceOnce once;
    if (sb::ensure_dyn_smem(once, sb::lab_tile_kernel, (int)sizeof(sb::LabShared)) != cudaSuccess) return SB_ERR_CUDA;
    int grid = h->num_sms * 2;
    if (grid > a.B) grid = a.B;
    sb::lab_tile_kernel<<<grid, sb::NT, sizeof(sb::LabShared), st>>>(a);
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 1;
    return SB_OK;
}
bool bad_img(const void* h, const void* p, int B, int H, int W) {
    return !h || !p || B <= 0 || H <= 0 || W <= 0 || (long long)H * W > (1LL << 24);
}
int aligned16(const void* a, const void* b, int npx) {
    return (((uintptr_t)a | (uintptr_t)b) % 16 == 0) && (((size_t)npx * 3) % 16 == 0);
}
dim3 tile_grid(int B, int npx, int num_sms) {
    const int G = (npx + sb::GROUP_PX - 1) / sb::GROUP_PX;
    int spans = (G + 255) / 256;
    int want = (num_sms * 16 + B - 1) / B;
    if (want < 1) want = 1;
    if (spans > want) spans = want;
    return dim3(B, spans);
}
}  // namespace

extern "C" {

int sb_reinhard_stats(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double* means, double* stds, void* stream) {
    if (bad_img(h, rgb, B, H, W) || !means || !stds) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_reinhard_stats");
    sb::LabArgs a{};
    a.in = rgb; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb, rgb, a.npx); a.mode = sb::REINHARD_STATS;
    a.means_out = means; a.stds_out = stds;
    return launch_lab(h, a, (cudaStream_t)stream);
}

int sb_reinhard_transform(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* target_means,
                          const double* target_stds, int mask_background, double luminosity_threshold, int32_t* status, void* stream) {
    if (bad_img(h, rgb_in, B, H, W) || !rgb_out || !target_means || !target_stds) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_reinhard_transform");
    sb::LabArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb_in, rgb_out, a.npx); a.mode = sb::REINHARD_TRANSFORM;
    a.tmeans = target_means; a.tstds = target_stds; a.mask_background = mask_background; a.lmax = lab_lmax(luminosity_threshold);
    a.status = status;
    return launch_lab(h, a, (cudaStream_t)stream);
}

int sb_standardize_brightness(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, void* stream) {
    if (bad_img(h, rgb_in, B, H, W) || !rgb_out) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_standardize_brightness");
    sb::LabArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb_in, rgb_out, a.npx); a.mode = sb::BRIGHTNESS_STANDARDIZE;
    return launch_lab(h, a, (cudaStream_t)stream);
}

int sb_lab_mean_std(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double* means, double* stds, void* stream) {
    if (bad_img(h, rgb, B, H, W) || !means || !stds) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_lab_mean_std");
    sb::LabArgs a{};
    a.in = rgb; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb, rgb, a.npx); a.mode = sb::REINHARD_STATS; a.skip_brightness = 1;
    a.means_out = means; a.stds_out = stds;
    return launch_lab(h, a, (cudaStream_t)stream);
}

int sb_lab_split(sb_handle* h, const uint8_t* rgb, size_t n_pixels, float* I1, float* I2, float* I3, void* stream) {
    if (!h || !rgb || !I1 || !I2 || !I3 || n_pixels == 0) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_lab_split");
    size_t blocks = (n_pixels + 255) / 256;
    if (blocks > (size_t)h->num_sms * 8) blocks = (size_t)h->num_sms * 8;
    sb::lab_split_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(h->tab, rgb, n_pixels, I1, I2, I3);
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 1;
    return SB_OK;
}

int sb_lab_merge(sb_handle* h, const void* I1, const void* I2, const void* I3, int is_f64, size_t n_pixels, uint8_t* rgb, void* stream) {
    if (!h || !rgb || !I1 || !I2 || !I3 || n_pixels == 0) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_lab_merge");
    size_t blocks = (n_pixels + 255) / 256;
    if (blocks > (size_t)h->num_sms * 8) blocks = (size_t)h->num_sms * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (is_f64) sb::lab_merge_kernel<double><<<(int)blocks, 256, 0, st>>>(h->tab, (const double*)I1, (const double*)I2, (const double*)I3, n_pixels, rgb);
    else sb::lab_merge_kernel<float><<<(int)blocks, 256, 0, st>>>(h->tab, (const float*)I1, (const float*)I2, (const float*)I3, n_pixels, rgb);
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 1;
    return SB_OK;
}

int sb_luminosity_standardize(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, double percentile, void* stream) {
    if (bad_img(h, rgb_in, B, H, W) || !rgb_out) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_luminosity_standardize");
    sb::LabArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb_in, rgb_out, a.npx); a.mode = sb::LUM_STANDARDIZE;
    a.percentile = percentile;
    return launch_lab(h, a, (cudaStream_t)stream);
}

int sb_hed_augment(sb_handle* hh, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* sigma, const double* bias,
                   double cutoff_lo, double cutoff_hi, double log_base, int skimage_variant, int32_t* status, void* stream) {
    if (bad_img(hh, rgb_in, B, H, W) || !rgb_out || !sigma || !bias) return SB_ERR_ARG;
    if (skimage_variant != 17 && skimage_variant != 18) return SB_ERR_ARG;
    if (skimage_variant == 17 && !(log_base > 1.0)) return SB_ERR_ARG;
    sb_handle* h = hh;
    cudaStream_t st = (cudaStream_t)stream;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_hed_augment");
    sb::Scratch scratch(h, st);
    const int npx = H * W, al = aligned16(rgb_in, rgb_out, npx);
    if (al) {
        // TMA ring: per-tile constants -> ring kernel (speculative transform + byte sums) -> gate kernel
        unsigned char* ws = nullptr;
        const size_t c_bytes = ((size_t)B * sizeof(sb::HedConsts) + 15) & ~(size_t)15;
        if (scratch.get(&ws, c_bytes + (size_t)B * 8) != cudaSuccess) return SB_ERR_CUDA;
        sb::HedConsts* consts = reinterpret_cast<sb::HedConsts*>(ws);
        unsigned long long* sums = reinterpret_cast<unsigned long long*>(ws + c_bytes);
        cudaMemsetAsync(sums, 0, (size_t)B * 8, st);
        sb::hed_prepare_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, sigma, bias, log_base, skimage_variant, consts);
        sb::RingGeom g{rgb_in, rgb_out, B, npx};
        sb::HedRingParams p{consts, sums};
        int rc = skimage_variant == 18 ? sb::launch_ring<sb::Hed18Op>(g, p, h->num_sms, st) : sb::launch_ring<sb::HedOp>(g, p, h->num_sms, st);
        if (rc == 0) {
            sb::hed_gate_kernel<<<B, 256, 0, st>>>(rgb_in, rgb_out, npx, sums, cutoff_lo, cutoff_hi, status);
            rc = (int)cudaGetLastError();
        }
        if (rc != 0) return SB_ERR_CUDA;
        h->launches += 3;
        return SB_OK;
    }
    unsigned long long* sums = nullptr;                 // [B] byte sums followed by [B] finished-CTA counters
    if (scratch.get(&sums, (size_t)B * 12) != cudaSuccess) return SB_ERR_CUDA;
    if (cudaMemsetAsync(sums, 0, (size_t)B * 12, st) != cudaSuccess) return SB_ERR_CUDA;
    const dim3 grid = tile_grid(B, npx, h->num_sms);
    sb::HedArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = npx; a.aligned = al; a.sigma = sigma; a.bias = bias;
    a.cutoff_lo = cutoff_lo; a.cutoff_hi = cutoff_hi; a.log_base = log_base; a.variant = skimage_variant; a.status = status; a.sums = sums;
    a.done = reinterpret_cast<unsigned*>(sums + B);
    static sb::DeviceOnce once17, once18;
    if (skimage_variant == 18) {
        if (sb::ensure_dyn_smem(once18, sb::hed_kernel<18>, sb::OD_REP_BYTES) != cudaSuccess) return SB_ERR_CUDA;
        sb::hed_kernel<18><<<grid, 256, sb::OD_REP_BYTES, st>>>(a);
    } else {
        if (sb::ensure_dyn_smem(once17, sb::hed_kernel<17>, sb::OD_REP_BYTES) != cudaSuccess) return SB_ERR_CUDA;
        sb::hed_kernel<17><<<grid, 256, sb::OD_REP_BYTES, st>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 1;
    return SB_OK;
}

int sb_hed_augment_f32(sb_handle* hh, const float* rgb_in, float* rgb_out, int B, int H, int W, const double* sigma, const double* bias,
                       double cutoff_lo, double cutoff_hi, double log_base, int skimage_variant, int32_t* status, void* stream) {
    if (bad_img(hh, rgb_in, B, H, W) || !rgb_out || !sigma || !bias) return SB_ERR_ARG;
    if (skimage_variant != 17 && skimage_variant != 18) return SB_ERR_ARG;
    if (skimage_variant == 17 && !(log_base > 1.0)) return SB_ERR_ARG;
    sb_handle* h = hh;
    cudaStream_t st = (cudaStream_t)stream;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_hed_augment_f32");
    sb::Scratch scratch(h, st);
    long long* sums = nullptr;
    if (scratch.get(&sums, (size_t)B * 8) != cudaSuccess) return SB_ERR_CUDA;
    if (cudaMemsetAsync(sums, 0, (size_t)B * 8, st) != cudaSuccess) return SB_ERR_CUDA;
    const int npx = H * W;
    int spans = (npx + 256 * 16 - 1) / (256 * 16);
    const int want = (h->num_sms * 8 + B - 1) / B;
    if (spans > want) spans = want;
    if (spans < 1) spans = 1;
    const dim3 grid(B, spans);
    sb::hed_float_sum_kernel<<<grid, 256, 0, st>>>(rgb_in, (size_t)npx * 3, sums);
    if (skimage_variant == 18)
        sb::hed_float_kernel<18><<<grid, 256, 0, st>>>(rgb_in, rgb_out, npx, sigma, bias, cutoff_lo, cutoff_hi, log_base, sums, status);
    else
        sb::hed_float_kernel<17><<<grid, 256, 0, st>>>(rgb_in, rgb_out, npx, sigma, bias, cutoff_lo, cutoff_hi, log_base, sums, status);
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 2;
    return SB_OK;
}

int sb_grayscale_augment(sb_handle* hh, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* alpha,
                         const double* beta, void* stream) {
    if (bad_img(hh, rgb_in, B, H, W) || !rgb_out || !alpha || !beta) return SB_ERR_ARG;
    sb_handle* h = hh;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_grayscale_augment");
    sb::GrayArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = aligned16(rgb_in, rgb_out, a.npx); a.alpha = alpha; a.beta = beta;
    if (a.aligned) {
        // TMA ring; the per-tile constants are the caller's alpha / beta arrays themselves
        sb::GrayRingParams p{sb::GrayConstsView{alpha, beta}};
        if (sb::launch_ring<sb::GrayOp>(sb::RingGeom{rgb_in, rgb_out, B, a.npx}, p, h->num_sms, (cudaStream_t)stream) != 0) return SB_ERR_CUDA;
        h->launches += 1;
        return SB_OK;
    }
    sb::gray_kernel<<<tile_grid(B, a.npx, h->num_sms), 256, 0, (cudaStream_t)stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += 1;
    return SB_OK;
}

}  // extern "C"
