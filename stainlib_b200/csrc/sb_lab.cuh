// sb_lab.cuh -- pieces of the 8-bit sRGB <-> CIELAB path (OpenCV's fixed-point arithmetic, oracle/cv_lab.py) shared by the
// per-tile kernel (sb_colour.cu) and the streaming Reinhard passes (sb_reinhard.cu).
#pragma once
#include "sb_device.cuh"

namespace sb {

__device__ __forceinline__ int ab_to_xz(int t) {
    // inverse companding in fixed point, C truncating division.  t <= 20545, so the cubic branch stays below 2^31 and,
    // being positive, divides by shifting; the linear branch (very dark colours, possibly negative t) is rare
    if (t > 3390) return (int)((((unsigned)(t * t) >> 14) * (unsigned)t) >> 14);
    return (t * 108) / 841 - 290;
}

// numpy.percentile (linear) of uint8-valued data given its exact histogram (n values), evaluated by one thread.
__device__ inline double hist_percentile(const unsigned* h, unsigned long long n, double pct) {
    const double vi = (double)(n - 1) * (pct / 100.0);
    unsigned long long lo = (unsigned long long)floor(vi);
    if (lo > n - 1) lo = n - 1;
    const unsigned long long hi = lo + 1 < n ? lo + 1 : n - 1;
    const double frac = vi - (double)lo;
    unsigned long long c = 0;
    int vlo = 255, vhi = 255;
    bool flo = false, fhi = false;
    for (int v = 0; v < 256; ++v) {
        c += h[v];
        if (!flo && c > lo) { vlo = v; flo = true; }
        if (!fhi && c > hi) { vhi = v; fhi = true; }
    }
    return lerp_np((double)vlo, (double)vhi, frac);
}

__device__ inline unsigned char trunc_clip_u8(double x) {
    // np.clip(x, 0, 255).astype(np.uint8); NaN -> 0
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 255.0) return 255;
    return (unsigned char)(int)x;
}

}  // namespace sb
