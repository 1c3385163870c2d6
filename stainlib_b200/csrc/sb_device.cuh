// sb_device.cuh -- device-side building blocks shared by the stain kernels (sm_100a).
//
// Per-pixel arithmetic is fp32 on 256-entry LUTs (OD is a function of a uint8: stain_utils.py:101-112); per-tile
// reductions are fp64 in a fixed order (deterministic); order statistics are exact selections on 23-bit monotone keys.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sb {

constexpr int NT = 512;             // threads per CTA for the tile kernels
constexpr int NWARP = NT / 32;
constexpr int GROUP_PX = 16;        // pixels per thread-iteration: 48 B = 3 x 16-byte vectors
constexpr int KEY_BITS = 23;
constexpr int L1_BITS = 12, L1_BINS = 1 << L1_BITS;   // level-1 histogram: top 12 bits of the key
constexpr int L2_BITS = 11, L2_BINS = 1 << L2_BITS;   // level-2 histogram: low 11 bits
// log2(255) rounded UP by one float step (2^x = 255 * (1 + 5.6e-7)): a pixel with zero concentrations must come out
// as exactly 255 like the reference's 255*exp(0), even when ex2.approx errs low by its 2^-22 bound.
constexpr float LOG2_255_UP = 7.994354248046875f;

// ------------------------------------------------------------------------------------------------ tables (device)
struct Tables {
    const float* od;        // [256]   OD LUT, fp32
    const float* gy;        // [3*256] pre-weighted luminance terms: coeffY[c] * gamma[v] (exact integers < 2^24)
    const double* od64;     // [256]
    const unsigned short* gamma;   // [256]
    const unsigned short* cbrt;    // [3072]
    const int* lab2yf;             // [512]
    const unsigned char* invgamma; // [4096]
};

// ------------------------------------------------------------------------------------------------ memory helpers
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_keep(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0u, 0x4440u + i); }

// Loads pixel group g (16 px = 12 words) of a tile. Full + 16B-aligned groups use three 16-byte loads; the ragged
// tail (or an unaligned tile) is assembled bytewise and padded with 255 (white = background).
template <bool KEEP>
__device__ __forceinline__ void load_group(const uint8_t* __restrict__ tile, int npx, int g, bool aligned, uint32_t (&w)[12], int& nvalid) {
    const int p0 = g * GROUP_PX;
    nvalid = min(GROUP_PX, npx - p0);
    const uint8_t* src = tile + (size_t)p0 * 3;
    if (aligned && nvalid == GROUP_PX) {
        const uint4* v = reinterpret_cast<const uint4*>(src);
        uint4 a = KEEP ? ldg_keep(v) : ldg_stream(v), b = KEEP ? ldg_keep(v + 1) : ldg_stream(v + 1), c = KEEP ? ldg_keep(v + 2) : ldg_stream(v + 2);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
        const int nbytes = nvalid * 3;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int k = i * 4 + j;
                uint32_t byte = k < nbytes ? (uint32_t)src[k] : 255u;
                x |= byte << (8 * j);
            }
            w[i] = x;
        }
    }
}

__device__ __forceinline__ void store_group(uint8_t* __restrict__ tile, int npx, int g, bool aligned, const uint32_t (&w)[12]) {
    const int p0 = g * GROUP_PX;
    const int nvalid = min(GROUP_PX, npx - p0);
    uint8_t* dst = tile + (size_t)p0 * 3;
    if (aligned && nvalid == GROUP_PX) {
        uint4* v = reinterpret_cast<uint4*>(dst);
        stg_stream(v, make_uint4(w[0], w[1], w[2], w[3]));
        stg_stream(v + 1, make_uint4(w[4], w[5], w[6], w[7]));
        stg_stream(v + 2, make_uint4(w[8], w[9], w[10], w[11]));
    } else {
        const int nbytes = nvalid * 3;
#pragma unroll
        for (int k = 0; k < 48; ++k)
            if (k < nbytes) dst[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
}

// Loads a group from memory THIS kernel has written (coherent ld.global, never the read-only .nc path).
__device__ __forceinline__ void load_group_rw(const uint8_t* tile, int npx, int g, bool aligned, uint32_t (&w)[12], int& nvalid) {
    const int p0 = g * GROUP_PX;
    nvalid = min(GROUP_PX, npx - p0);
    const uint8_t* src = tile + (size_t)p0 * 3;
    if (aligned && nvalid == GROUP_PX) {
        uint4 a, b, c;
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src));
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(src + 16));
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(src + 32));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
        const int nbytes = nvalid * 3;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = i * 4 + j;
                uint32_t byte = 255u;
                if (k < nbytes) asm volatile("ld.global.u8 %0, [%1];" : "=r"(byte) : "l"(src + k));
                x |= byte << (8 * j);
            }
            w[i] = x;
        }
    }
}

// Same with default caching (the data is read back soon: keep it in L2).
__device__ __forceinline__ void store_group_keep(uint8_t* __restrict__ tile, int npx, int g, bool aligned, const uint32_t (&w)[12]) {
    const int p0 = g * GROUP_PX;
    const int nvalid = min(GROUP_PX, npx - p0);
    uint8_t* dst = tile + (size_t)p0 * 3;
    if (aligned && nvalid == GROUP_PX) {
        uint4* v = reinterpret_cast<uint4*>(dst);
        v[0] = make_uint4(w[0], w[1], w[2], w[3]);
        v[1] = make_uint4(w[4], w[5], w[6], w[7]);
        v[2] = make_uint4(w[8], w[9], w[10], w[11]);
    } else {
        const int nbytes = nvalid * 3;
#pragma unroll
        for (int k = 0; k < 48; ++k)
            if (k < nbytes) dst[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    }
}

// Calls f(pixel_index_in_group, r, g, b) for the 16 pixels of a group.
template <class F>
__device__ __forceinline__ void for_each_px(const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, byte_of(a, 0), byte_of(a, 1), byte_of(a, 2));
        f(4 * q + 1, byte_of(a, 3), byte_of(b, 0), byte_of(b, 1));
        f(4 * q + 2, byte_of(b, 2), byte_of(b, 3), byte_of(c, 0));
        f(4 * q + 3, byte_of(c, 1), byte_of(c, 2), byte_of(c, 3));
    }
}

// Packs four values (each already reduced to its low byte inside a 32-bit word) into one word.
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

// ------------------------------------------------------------------------------------------------ per-pixel math
// 2-atom non-negative LASSO in closed form (oracle: lasso_pos2; reference call site stain_utils.py:78).
struct LassoK {
    float m00, m01, m02, m10, m11, m12;  // dictionary rows (stain vectors)
    float lam, lam1;                     // l1 weight of atom 0 / atom 1 (equal except in the normalised form of make_dict_lasso_consts)
    float i00, i01, i11;                 // inverse Gram matrix
    float rg00, rg11, g01;               // 1/G00, 1/G11, G01
};

__device__ __forceinline__ void lasso2(const LassoK& k, float o0, float o1, float o2, float& c0, float& c1) {
    const float u0 = fmaf(k.m02, o2, fmaf(k.m01, o1, fmaf(k.m00, o0, -k.lam)));
    const float u1 = fmaf(k.m12, o2, fmaf(k.m11, o1, fmaf(k.m10, o0, -k.lam)));
    const float a0 = fmaf(k.i01, u1, k.i00 * u0);
    const float a1 = fmaf(k.i11, u1, k.i01 * u0);
    const bool both = (a0 > 0.f) & (a1 > 0.f);
    const float p0 = fmaxf(u0, 0.f) * k.rg00, p1 = fmaxf(u1, 0.f) * k.rg11;
    const bool only0 = (p0 > 0.f) & (fmaf(-k.g01, p0, u1) <= 0.f);
    const bool only1 = (p1 > 0.f) & (fmaf(-k.g01, p1, u0) <= 0.f);
    c0 = both ? a0 : (only0 ? p0 : 0.f);
    c1 = both ? a1 : ((!only0 & only1) ? p1 : 0.f);
}

// Same problem when the dictionary rows have unit norm (Gram diagonal exactly 1 in fp32), g = G01.  With u = G a:
//   g >= 0:  a1 > 0  <=>  a0 < u0, and the optimum is  c0 = max(0, min(a0, u0))   (a1 <= 0 puts c on axis 0: max(u0, 0));
//   g <  0:  a1 > 0  <=>  a0 > u0, and the optimum is  c0 = max(0, a0, u0)        (one 3-input FMNMX3 on sm_100);
// symmetrically for c1.  No compares, no selects: 4 (or 2) min/max instructions per pixel.
__device__ __forceinline__ float max3f(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
enum { LASSO_GENERAL = 0, LASSO_UNIT_POS = 1, LASSO_UNIT_NEG = 2 };
template <int LM>
__device__ __forceinline__ void lasso2_select(float a0, float a1, float u0, float u1, float& c0, float& c1) {
    if (LM == LASSO_UNIT_POS) { c0 = fmaxf(0.f, fminf(a0, u0)); c1 = fmaxf(0.f, fminf(a1, u1)); }
    else { c0 = max3f(0.f, a0, u0); c1 = max3f(0.f, a1, u1); }
}
template <int LM>
__device__ __forceinline__ void lasso2_unit(const LassoK& k, float o0, float o1, float o2, float& c0, float& c1) {
    const float u0 = fmaf(k.m02, o2, fmaf(k.m01, o1, fmaf(k.m00, o0, -k.lam)));
    const float u1 = fmaf(k.m12, o2, fmaf(k.m11, o1, fmaf(k.m10, o0, -k.lam1)));
    const float a0 = fmaf(k.i01, u1, k.i00 * u0);
    const float a1 = fmaf(k.i11, u1, k.i01 * u0);
    lasso2_select<LM>(a0, a1, u0, u1, c0, c1);
}
__host__ __device__ inline int lasso_mode_of(float rg00, float rg11, float g01) {
    return (rg00 == 1.0f && rg11 == 1.0f) ? (g01 >= 0.f ? LASSO_UNIT_POS : LASSO_UNIT_NEG) : LASSO_GENERAL;
}

__host__ __device__ inline void make_lasso_consts(const double M[6], double lam, LassoK& k) {
    const double g00 = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    const double g11 = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    const double g01 = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    const double det = g00 * g11 - g01 * g01;
    k.m00 = (float)M[0]; k.m01 = (float)M[1]; k.m02 = (float)M[2];
    k.m10 = (float)M[3]; k.m11 = (float)M[4]; k.m12 = (float)M[5];
    k.lam = (float)lam; k.lam1 = (float)lam;
    k.i00 = (float)(g11 / det); k.i01 = (float)(-g01 / det); k.i11 = (float)(g00 / det);
    k.rg00 = (float)(1.0 / g00); k.rg11 = (float)(1.0 / g11); k.g01 = (float)g01;
}

// The same problem for a dictionary whose atoms are NOT on the unit sphere (iterates of the Vahadane dictionary learning:
// the Anderson step leaves the norms slightly below 1), in normalised form: with d^_j = d_j / |d_j| and beta_j = alpha_j |d_j|
// the objective reads  1/2 |x - sum beta_j d^_j|^2 + sum (lambda / |d_j|) beta_j  -- unit Gram diagonal, one l1 weight per
// atom -- so the packed compare-free solver applies to every iterate.  The passes accumulate their sums in beta space;
// dict_scale_sums takes them back: alpha_j = s_j beta_j with s_j = 1 / |d_j|.  An atom that has died (zero vector) gets a
// zero direction and a weight that keeps its code at zero.
__host__ __device__ inline void dict_scales(const double D[6], double s[2]) {
    for (int j = 0; j < 2; ++j) {
        const double n2 = D[3 * j] * D[3 * j] + D[3 * j + 1] * D[3 * j + 1] + D[3 * j + 2] * D[3 * j + 2];
        s[j] = n2 > 1e-60 ? 1.0 / sqrt(n2) : 0.0;
    }
}
__host__ __device__ inline void make_dict_lasso_consts(const double D[6], double lam, LassoK& k) {
    double s[2];
    dict_scales(D, s);
    double dn[6];
    for (int j = 0; j < 2; ++j)
        for (int c = 0; c < 3; ++c) dn[3 * j + c] = D[3 * j + c] * s[j];
    const double g01 = dn[0] * dn[3] + dn[1] * dn[4] + dn[2] * dn[5];
    const double det = 1.0 - g01 * g01;
    k.m00 = (float)dn[0]; k.m01 = (float)dn[1]; k.m02 = (float)dn[2];
    k.m10 = (float)dn[3]; k.m11 = (float)dn[4]; k.m12 = (float)dn[5];
    k.lam = s[0] > 0.0 ? (float)(lam * s[0]) : 1.0f;
    k.lam1 = s[1] > 0.0 ? (float)(lam * s[1]) : 1.0f;
    k.i00 = (float)(1.0 / det); k.i01 = (float)(-g01 / det); k.i11 = (float)(1.0 / det);
    k.rg00 = 1.0f; k.rg11 = 1.0f; k.g01 = (float)g01;
}
// t = (sum b0 b0, sum b0 b1, sum b1 b1, sum x b0 (3), sum x b1 (3)) in beta space -> alpha space.
__host__ __device__ inline void dict_scale_sums(double* t, const double s[2]) {
    t[0] *= s[0] * s[0]; t[1] *= s[0] * s[1]; t[2] *= s[1] * s[1];
    for (int c = 0; c < 3; ++c) { t[3 + c] *= s[0]; t[6 + c] *= s[1]; }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// uint8(x) exactly as numpy's astype(np.uint8) does on x86-64 for x >= 0: truncate via int32, keep the low byte;
// out of int32 range / inf / NaN -> 0 (normalizer.py:50 does not clip).  Returns a word whose LOW BYTE is the value.
__device__ __forceinline__ uint32_t wrap_u8_bits(float x) {
    if (x < 8388608.f) return __float_as_uint(__fadd_rd(x, 8388608.f));  // mantissa = floor(x); low byte = floor(x) & 255
    if (x < 2147483648.f) return (uint32_t)__float2int_rz(x);
    return 0u;
}
// uint8(clip(x, 0, 255)) for x >= 0 or NaN (augmenter.py:447).
__device__ __forceinline__ uint32_t clip_u8_bits(float x) {
    return __float_as_uint(__fadd_rd(fminf(x, 255.f), 8388608.f));
}

// ------------------------------------------------------------------------------------------------ K4 arithmetic
// Lane-replicated OD table (256-byte rows: words 0..31 = one copy per lane, words 32..63 free for the caller), the
// PRMT-built lookup offset, and the packed two-pixel recombine step shared by sb_recombine.cu and sb_pipeline.cu.
constexpr int OD_ROW_BYTES = 256;             // row stride of the lane-replicated OD table
constexpr int OD_REP_BYTES = 256 * OD_ROW_BYTES;

struct __align__(16) K4Consts {
    float m[6];      // source stain matrix rows
    float nlam;      // -lambda
    float i00, i01, i11;
    float rg00, rg11, g01;
    float A[6];      // -scale_j * Mt_jk * log2(e)
    int lasso_mode, need_check;   // LASSO_GENERAL / LASSO_UNIT_POS / LASSO_UNIT_NEG; need_check: 255*2^e may reach 2^23
    int mode;        // 0 = recombine, 1 = write zeros (reference divides by a zero percentile), 2 = copy the input through
};

__device__ __forceinline__ float od_lookup(const unsigned char* tab, uint32_t w, uint32_t lane_off, int k) {
    // offset = (byte k of w) << 8 | lane << 2 : one PRMT
    const uint32_t off = __byte_perm(w, lane_off, 0x6504u | (k << 4));
    return *reinterpret_cast<const float*>(tab + off);
}

// Same lookup with an ABSOLUTE shared-window address: the table sits at a 64 KB-aligned shared address T, and
// lane_base = (lane << 2) | ((T >> 16) << 8), so the one PRMT yields T | value << 8 | lane << 2 -- no address add.
struct OdAbs { uint32_t lane_base; };
__device__ __forceinline__ float od_lookup(const OdAbs& t, uint32_t w, uint32_t /*lane_off*/, int k) {
    const uint32_t addr = __byte_perm(w, t.lane_base, 0x6504u | (k << 4));
    float r;
    asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ float od_lookup(const OdAbs* t, uint32_t w, uint32_t lane_off, int k) { return od_lookup(*t, w, lane_off, k); }
// The two halves of a pair-table entry as separate 4-byte loads (same PRMT address, immediate offset): lets the od
// values and the gamma values of two pixels land in adjacent registers for the packed f32x2 pipe.
__device__ __forceinline__ void odg_lookup_split(const OdAbs& t, uint32_t w, int k, float& od, float& gam) {
    const uint32_t addr = __byte_perm(w, t.lane_base, 0x6504u | (k << 4));
    asm("ld.shared.f32 %0, [%1];" : "=f"(od) : "r"(addr));
    asm("ld.shared.f32 %0, [%1+4];" : "=f"(gam) : "r"(addr));
}
// {od, gamma} pair table (8 bytes per lane, lane_base = lane << 3 | T >> 16 << 8): one LDS.64.
__device__ __forceinline__ float2 odg_lookup_abs(const OdAbs& t, uint32_t w, int k) {
    const uint32_t addr = __byte_perm(w, t.lane_base, 0x6504u | (k << 4));
    float2 r;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 dup(float a) { return make_float2(a, a); }

// Two pixels at a time on the packed f32x2 pipe (each half is bitwise the scalar lasso2_unit of that pixel).
template <int LM>
__device__ __forceinline__ void lasso2_unit_pair(const LassoK& k, const float2 o0, const float2 o1, const float2 o2, float2& c0, float2& c1) {
    const float2 u0 = __ffma2_rn(dup(k.m02), o2, __ffma2_rn(dup(k.m01), o1, __ffma2_rn(dup(k.m00), o0, dup(-k.lam))));
    const float2 u1 = __ffma2_rn(dup(k.m12), o2, __ffma2_rn(dup(k.m11), o1, __ffma2_rn(dup(k.m10), o0, dup(-k.lam1))));
    const float2 a0 = __ffma2_rn(dup(k.i01), u1, __fmul2_rn(dup(k.i00), u0));
    const float2 a1 = __ffma2_rn(dup(k.i11), u1, __fmul2_rn(dup(k.i01), u0));
    lasso2_select<LM>(a0.x, a1.x, u0.x, u1.x, c0.x, c1.x);
    lasso2_select<LM>(a0.y, a1.y, u0.y, u1.y, c0.y, c1.y);
}

template <bool CHECK, int LM>
__device__ __forceinline__ void recombine_pair(const K4Consts& k, const float2 o0, const float2 o1, const float2 o2, uint32_t (&bits)[6]) {
    const float2 u0 = __ffma2_rn(dup(k.m[2]), o2, __ffma2_rn(dup(k.m[1]), o1, __ffma2_rn(dup(k.m[0]), o0, dup(k.nlam))));
    const float2 u1 = __ffma2_rn(dup(k.m[5]), o2, __ffma2_rn(dup(k.m[4]), o1, __ffma2_rn(dup(k.m[3]), o0, dup(k.nlam))));
    const float2 a0 = __ffma2_rn(dup(k.i01), u1, __fmul2_rn(dup(k.i00), u0));
    const float2 a1 = __ffma2_rn(dup(k.i11), u1, __fmul2_rn(dup(k.i01), u0));
    float2 c0, c1;
    if (LM != LASSO_GENERAL) {
        lasso2_select<LM>(a0.x, a1.x, u0.x, u1.x, c0.x, c1.x);
        lasso2_select<LM>(a0.y, a1.y, u0.y, u1.y, c0.y, c1.y);
    } else {
        // general Gram diagonal: KKT form (same as lasso2 in sb_device.cuh)
        const float p0a = fmaxf(u0.x, 0.f) * k.rg00, p1a = fmaxf(u1.x, 0.f) * k.rg11;
        const float p0b = fmaxf(u0.y, 0.f) * k.rg00, p1b = fmaxf(u1.y, 0.f) * k.rg11;
        const bool ba = (a0.x > 0.f) & (a1.x > 0.f), bb = (a0.y > 0.f) & (a1.y > 0.f);
        const bool o0a = (p0a > 0.f) & (fmaf(-k.g01, p0a, u1.x) <= 0.f), o1a = (p1a > 0.f) & (fmaf(-k.g01, p1a, u0.x) <= 0.f);
        const bool o0b = (p0b > 0.f) & (fmaf(-k.g01, p0b, u1.y) <= 0.f), o1b = (p1b > 0.f) & (fmaf(-k.g01, p1b, u0.y) <= 0.f);
        c0.x = ba ? a0.x : (o0a ? p0a : 0.f); c1.x = ba ? a1.x : ((!o0a & o1a) ? p1a : 0.f);
        c0.y = bb ? a0.y : (o0b ? p0b : 0.f); c1.y = bb ? a1.y : ((!o0b & o1b) ? p1b : 0.f);
    }
    const float2 L = dup(LOG2_255_UP);
    const float2 e0 = __ffma2_rn(c1, dup(k.A[3]), __ffma2_rn(c0, dup(k.A[0]), L));
    const float2 e1 = __ffma2_rn(c1, dup(k.A[4]), __ffma2_rn(c0, dup(k.A[1]), L));
    const float2 e2 = __ffma2_rn(c1, dup(k.A[5]), __ffma2_rn(c0, dup(k.A[2]), L));
    const float2 x0 = f2(ex2_approx(e0.x), ex2_approx(e0.y));
    const float2 x1 = f2(ex2_approx(e1.x), ex2_approx(e1.y));
    const float2 x2 = f2(ex2_approx(e2.x), ex2_approx(e2.y));
    if (!CHECK) {
        const float2 MAGIC = dup(8388608.f);
        const float2 r0 = __fadd2_rd(x0, MAGIC), r1 = __fadd2_rd(x1, MAGIC), r2 = __fadd2_rd(x2, MAGIC);
        bits[0] = __float_as_uint(r0.x); bits[1] = __float_as_uint(r1.x); bits[2] = __float_as_uint(r2.x);
        bits[3] = __float_as_uint(r0.y); bits[4] = __float_as_uint(r1.y); bits[5] = __float_as_uint(r2.y);
    } else {
        bits[0] = wrap_u8_bits(x0.x); bits[1] = wrap_u8_bits(x1.x); bits[2] = wrap_u8_bits(x2.x);
        bits[3] = wrap_u8_bits(x0.y); bits[4] = wrap_u8_bits(x1.y); bits[5] = wrap_u8_bits(x2.y);
    }
}


__device__ __forceinline__ void fill_od_rep(unsigned char* od_rep, const float* od, int nthreads) {
    for (int i = threadIdx.x; i < 256 * 32; i += nthreads)
        *reinterpret_cast<float*>(od_rep + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = od[i >> 5];
}

// Fills the recombine constants from a source matrix, the per-stain scale and the target matrix (all fp64).
__device__ inline void make_k4_consts(const double M[6], double lam, const double scale[2], const double Mt[6], K4Consts& k) {
    LassoK lk;
    make_lasso_consts(M, lam, lk);
    k.m[0] = lk.m00; k.m[1] = lk.m01; k.m[2] = lk.m02; k.m[3] = lk.m10; k.m[4] = lk.m11; k.m[5] = lk.m12;
    k.nlam = -lk.lam; k.i00 = lk.i00; k.i01 = lk.i01; k.i11 = lk.i11; k.rg00 = lk.rg00; k.rg11 = lk.rg11; k.g01 = lk.g01;
    k.lasso_mode = lasso_mode_of(lk.rg00, lk.rg11, lk.g01);
    const double LOG2E = 1.4426950408889634;
    bool finite = true;
    for (int j = 0; j < 2; ++j) {
        finite = finite && isfinite(scale[j]);
        for (int c = 0; c < 3; ++c) k.A[3 * j + c] = (float)(-scale[j] * Mt[3 * j + c] * LOG2E);
    }
    // The magic-add uint8 wrap is exact while 255 * 2^e < 2^23.  With a negative target-matrix entry some A is positive
    // and e can exceed log2(255); bound it from the largest concentration any uint8 pixel can have (OD <= ln 255):
    //   g >= 0: G_jj c_j <= u_j;   g < 0: c_0 <= max(u_0 / G_00, i00 u_0 + i01 u_1) with i01 > 0 (and symmetrically).
    const double ODMAX = 5.541263545158426;
    const double g00 = M[0] * M[0] + M[1] * M[1] + M[2] * M[2], g11 = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    const double g01 = M[0] * M[3] + M[1] * M[4] + M[2] * M[5], det = g00 * g11 - g01 * g01;
    double U[2];
    for (int j = 0; j < 2; ++j) {
        double s = -lam;
        for (int c = 0; c < 3; ++c) s += (M[3 * j + c] > 0.0 ? M[3 * j + c] : 0.0) * ODMAX;
        U[j] = s > 0.0 ? s : 0.0;
    }
    double cb[2] = {U[0] / g00, U[1] / g11};
    if (g01 < 0.0) {
        const double b0 = (g11 * U[0] - g01 * U[1]) / det, b1 = (g00 * U[1] - g01 * U[0]) / det;
        cb[0] = b0 > cb[0] ? b0 : cb[0];
        cb[1] = b1 > cb[1] ? b1 : cb[1];
    }
    bool need = false;
    for (int c = 0; c < 3; ++c) {
        const double a0 = k.A[c] > 0.f ? (double)k.A[c] : 0.0, a1 = k.A[3 + c] > 0.f ? (double)k.A[3 + c] : 0.0;
        const double emax = 8.0 + cb[0] * a0 + cb[1] * a1;
        need = need || !(emax < 22.0) || !isfinite(k.A[c]) || !isfinite(k.A[3 + c]);
    }
    k.need_check = need ? 1 : 0;
    k.mode = finite ? 0 : 1;
}

// 16 pixels (12 packed words) -> 16 recombined pixels.
template <bool CHECK, int LM, class TAB>
__device__ __forceinline__ void recombine_words(const K4Consts& k, const TAB tab, const uint32_t (&w)[12], uint32_t (&o)[12], uint32_t lane_off) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
        // pixels: p0=(a0,a1,a2) p1=(a3,b0,b1) p2=(b2,b3,c0) p3=(c1,c2,c3)
        uint32_t b01[6], b23[6];
        recombine_pair<CHECK, LM>(k, f2(od_lookup(tab, wa, lane_off, 0), od_lookup(tab, wa, lane_off, 3)),
                                    f2(od_lookup(tab, wa, lane_off, 1), od_lookup(tab, wb, lane_off, 0)),
                                    f2(od_lookup(tab, wa, lane_off, 2), od_lookup(tab, wb, lane_off, 1)), b01);
        recombine_pair<CHECK, LM>(k, f2(od_lookup(tab, wb, lane_off, 2), od_lookup(tab, wc, lane_off, 1)),
                                    f2(od_lookup(tab, wb, lane_off, 3), od_lookup(tab, wc, lane_off, 2)),
                                    f2(od_lookup(tab, wc, lane_off, 0), od_lookup(tab, wc, lane_off, 3)), b23);
        o[3 * q] = pack4(b01[0], b01[1], b01[2], b01[3]);
        o[3 * q + 1] = pack4(b01[4], b01[5], b23[0], b23[1]);
        o[3 * q + 2] = pack4(b23[2], b23[3], b23[4], b23[5]);
    }
}

// ------------------------------------------------------------------------------------------------ selection keys
// Monotone 23-bit key of the angle atan2(y, x): the "diamond angle" d in [-2,2] mapped to t = d/4 + 1.5 in [1,2];
// key = mantissa bits of t.  Exact order statistics of the key give the order statistics of the angle.
__device__ __forceinline__ uint32_t angle_key(float x, float y) {
    const float s = fabsf(x) + fabsf(y);
    float r = y * rcp_approx(s);
    if (!(s > 0.f)) r = 0.f;
    const float d = (x >= 0.f) ? r : ((y >= 0.f) ? 2.f - r : -2.f - r);
    const float t = fmaf(d, 0.25f, 1.5f);
    const uint32_t k = __float_as_uint(t) - 0x3F800000u;
    return min(k, (1u << KEY_BITS) - 1u);
}
__device__ inline double angle_from_key(uint32_t key) {
    const double t = 1.0 + (double)key * (1.0 / 8388608.0);
    const double d = (t - 1.5) * 4.0;
    if (d > 1.0) return atan2(2.0 - d, -(d - 1.0));
    if (d < -1.0) return atan2(-2.0 - d, -(-1.0 - d));
    return atan2(d, 1.0 - fabs(d));
}
// Monotone 23-bit key of a concentration C >= 0 with unbounded range: t = 2 - K/(C+K) in [1,2), key = mantissa of t.
// Resolution dC = 2^-23 (C+K)^2 / K: 7e-7 at C = 1.5, 6e-6 at C = 8 (K = 2).
constexpr float CONC_KEY_K = 2.0f;
__device__ __forceinline__ uint32_t conc_key(float c) {
    const float t = fmaf(-CONC_KEY_K, rcp_approx(c + CONC_KEY_K), 2.f);
    const uint32_t k = __float_as_uint(t) - 0x3F800000u;
    return (c > 0.f) ? min(k, (1u << KEY_BITS) - 1u) : 0u;
}
__device__ inline double conc_from_key(uint32_t key) {
    if (key == 0) return 0.0;
    const double t = 1.0 + (double)key * (1.0 / 8388608.0);
    return (double)CONC_KEY_K * (t - 1.0) / (2.0 - t);
}

// Diamond coordinate d = y/(|x|+|y|) (|d| <= 1 in the half-plane x >= 0) that maps to a given angle key.
__device__ inline double diamond_from_key(double key) { return ((1.0 + key * (1.0 / 8388608.0)) - 1.5) * 4.0; }
// Largest float strictly below x (x finite, positive or negative).
__device__ inline float float_below(double x) {
    float f = (float)x;
    if ((double)f >= x) f = nextafterf(f, -INFINITY);
    return nextafterf(f, -INFINITY);
}
__device__ inline float float_above(double x) {
    float f = (float)x;
    if ((double)f <= x) f = nextafterf(f, INFINITY);
    return nextafterf(f, INFINITY);
}

// numpy.percentile (linear): virtual index (n-1)*q/100; returns lo index and the interpolation weight.
__device__ inline void percentile_index(unsigned n, double pct, unsigned& lo, unsigned& hi, double& frac) {
    const double vi = (double)(n - 1) * (pct / 100.0);
    double fl = floor(vi);
    if (fl < 0) fl = 0;
    if (fl > (double)(n - 1)) fl = (double)(n - 1);
    lo = (unsigned)fl;
    hi = min(lo + 1u, n - 1u);
    frac = vi - fl;
}
__device__ inline double lerp_np(double a, double b, double t) {
    // numpy's _lerp: a + (b-a)*t, computed from the far end when t >= 0.5
    const double d = b - a;
    return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

// ------------------------------------------------------------------------------------------------ 3x3 symmetric eig
// Cyclic Jacobi in fp64 (single thread).  a = {a00,a01,a02,a11,a12,a22}.  Outputs eigenvalues w[3] (unsorted) and
// eigenvectors as columns of v[3][3] (v[r][c]).
__device__ inline void jacobi_eig3(const double a_in[6], double w[3], double v[3][3]) {
    double a[3][3] = {{a_in[0], a_in[1], a_in[2]}, {a_in[1], a_in[3], a_in[4]}, {a_in[2], a_in[4], a_in[5]}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = a[p][q];
                if (apq == 0.0) continue;
                // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)) with theta = d / (2 apq), written with one square root, one
                // division and one reciprocal square root (this runs on ONE thread per tile: fp64 latency is what it costs)
                const double d = a[q][q] - a[p][p];
                const double t = (d >= 0 ? 2.0 : -2.0) * apq / (fabs(d) + sqrt(d * d + 4.0 * apq * apq));
                const double c = rsqrt(t * t + 1.0), s = t * c;
                const int r = 3 - p - q;
                const double app = a[p][p], aqq = a[q][q], arp = a[r][p], arq = a[r][q];
                a[p][p] = app - t * apq;
                a[q][q] = aqq + t * apq;
                a[p][q] = a[q][p] = 0.0;
                a[r][p] = a[p][r] = c * arp - s * arq;
                a[r][q] = a[q][r] = s * arp + c * arq;
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
}

// ------------------------------------------------------------------------------------------------ Anderson acceleration
// The Vahadane dictionary iteration D <- F(D) (sparse-code, accumulate, one block-coordinate update) converges linearly
// at ~0.83 per pass; Anderson acceleration (type II, memory m <= AA_MAX) of the 6-component fixed-point map reaches
// the same fixed point in a fifth of the passes.  Single thread, fp64.  Mirrored by oracle/stain_oracle.py
// (anderson_step): residual growth drops the history; the extrapolated point is projected back onto
// {D >= 0, ||d_j|| <= 1}; a singular or non-finite solve falls back to the plain iterate.
constexpr int AA_MAX = 4;
// Lives in shared memory; one thread steps it.  The history is kept as difference columns (circular) with their Gram
// matrix maintained incrementally, so a step costs a few hundred fp64 operations with compile-time addressing.
struct AAState {
    double dx[AA_MAX][6], dr[AA_MAX][6];   // x_{i+1} - x_i and r_{i+1} - r_i of the last <= m steps
    double G[AA_MAX][AA_MAX];              // dr^T dr
    double px[6], pr[6];                   // previous iterate and its residual
    double W[AA_MAX][AA_MAX + 1];          // elimination scratch
    double last;
    int nd, head, have_prev;
};
__device__ inline void aa_reset(AAState& s) { s.nd = 0; s.head = 0; s.have_prev = 0; s.last = -1.0; }
// Switch to a related map (sample -> full tile): keep the difference columns, forget the last iterate and residual norm.
__device__ inline void aa_carry(AAState& s) { s.have_prev = 0; s.last = -1.0; }
// D: current iterate (in) / next iterate (out); FD = F(D).
__device__ inline void aa_step(AAState& s, int m, double* D, const double* FD) {
    double r[6], rn = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { r[k] = FD[k] - D[k]; rn += r[k] * r[k]; }
    rn = sqrt(rn);
    bool ok = m > 0;
    if (ok) {
        if (s.last >= 0.0 && rn > s.last) { s.nd = 0; s.head = 0; s.have_prev = 0; }   // the last step made things worse: restart
        s.last = rn;
        if (s.have_prev) {
            const int c = s.head;
#pragma unroll
            for (int k = 0; k < 6; ++k) { s.dx[c][k] = D[k] - s.px[k]; s.dr[c][k] = r[k] - s.pr[k]; }
            if (s.nd < m) s.nd += 1;
            s.head = (c + 1 == m) ? 0 : c + 1;
#pragma unroll
            for (int j = 0; j < AA_MAX; ++j) {
                if (j < s.nd) {
                    double g = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) g += s.dr[c][k] * s.dr[j][k];
                    s.G[c][j] = g; s.G[j][c] = g;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) { s.px[k] = D[k]; s.pr[k] = r[k]; }
        s.have_prev = 1;
        ok = s.nd >= 1;
    }
    double xn[6];
    if (ok) {
        const int nd = s.nd;
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < AA_MAX; ++i) if (i < nd) tr += s.G[i][i];
        ok = tr > 0.0 && isfinite(tr);
        // [G + 1e-10 tr I | dr^T r], padded with identity rows to a fixed 4 x 5 system
#pragma unroll
        for (int i = 0; i < AA_MAX; ++i) {
#pragma unroll
            for (int j = 0; j < AA_MAX; ++j)
                s.W[i][j] = (i < nd && j < nd) ? s.G[i][j] + (i == j ? 1e-10 * tr : 0.0) : (i == j ? 1.0 : 0.0);
            double b = 0.0;
            if (i < nd) {
#pragma unroll
                for (int k = 0; k < 6; ++k) b += s.dr[i][k] * r[k];
            }
            s.W[i][AA_MAX] = b;
        }
        // Gaussian elimination (symmetric positive definite after the regularisation: no pivoting) and back substitution
#pragma unroll
        for (int c = 0; c < AA_MAX; ++c) {
            const double piv = s.W[c][c];
            ok = ok && piv > 0.0;
#pragma unroll
            for (int i = c + 1; i < AA_MAX; ++i) {
                const double f = s.W[i][c] / piv;
#pragma unroll
                for (int j = c + 1; j <= AA_MAX; ++j) s.W[i][j] -= f * s.W[c][j];
            }
        }
        double gam[AA_MAX];
#pragma unroll
        for (int i = AA_MAX - 1; i >= 0; --i) {
            double v = s.W[i][AA_MAX];
#pragma unroll
            for (int j = i + 1; j < AA_MAX; ++j) v -= s.W[i][j] * gam[j];
            gam[i] = v / s.W[i][i];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = D[k] + r[k];
#pragma unroll
            for (int i = 0; i < AA_MAX; ++i) if (i < nd) v -= gam[i] * (s.dx[i][k] + s.dr[i][k]);
            xn[k] = v > 0.0 ? v : 0.0;
            ok = ok && isfinite(v);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double nrm = sqrt(xn[3 * j] * xn[3 * j] + xn[3 * j + 1] * xn[3 * j + 1] + xn[3 * j + 2] * xn[3 * j + 2]);
            const double sc = 1.0 / (nrm > 1.0 ? nrm : 1.0);
#pragma unroll
            for (int k = 0; k < 3; ++k) xn[3 * j + k] *= sc;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) D[k] = ok ? xn[k] : FD[k];
}

// ------------------------------------------------------------------------------------------------ block reductions
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ unsigned warp_incl_scan(unsigned x) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

}  // namespace sb
