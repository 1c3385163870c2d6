// sb_api.cu -- the extern "C" boundary declared in include/stainb200.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "sb_kernels.h"
#include "sb_tables.inc"

namespace {

thread_local std::string g_last_cuda_error;

int cuda_fail(cudaError_t e, const char* what) {
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return SB_ERR_CUDA;
}
#define SB_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

}  // namespace


namespace {

// Largest luminance table index whose L satisfies L/255.0 < thr (float64), -1 if none.  L is monotone in the index.
int mask_y_bound(double thr) {
    int best = -1;
    for (int i = 0; i < 3072; ++i) {
        long long L = ((long long)SB_LAB_LSCALE * SB_CBRT_TAB[i] + SB_LAB_LSHIFT + (1 << 14)) >> 15;
        if (L < 0) L = 0;
        if (L > 255) L = 255;
        if ((double)L / 255.0 < thr) best = i;
    }
    return best;
}
float mask_ybound_f(double thr) {
    const int yb = mask_y_bound(thr);
    return (float)((double)(yb + 1) * 4096.0 - 2048.0);
}

int check_image_args(const sb_handle* h, const void* rgb, int B, int H, int W) {
    if (!h || !rgb || B <= 0 || H <= 0 || W <= 0) return SB_ERR_ARG;
    if ((long long)H * W > (1LL << 24)) return SB_ERR_UNSUPPORTED;   // histogram counters / rank arithmetic are 32-bit
    return SB_OK;
}
int is_aligned(const void* a, const void* b, int npx) {
    return (((uintptr_t)a | (uintptr_t)b) % 16 == 0) && (((size_t)npx * 3) % 16 == 0);
}
int pick_cluster(const sb_handle* h, int B, int npx, int requested) {
    if (requested == 1 || requested == 2 || requested == 4 || requested == 8) return requested;
    // auto: one CTA per tile (the fast single-pass selection path) whenever the batch can fill the SMs; clusters of
    // 2/4/8 CTAs over DSMEM only to spread a small batch over the machine
    int S = 1;
    while (S < 8 && (long long)B * S < h->num_sms) S *= 2;
    while (S > 1 && npx / S < 16 * sb::NT) S /= 2;
    return S;
}

int run_pipeline(sb_handle* h, int mode, const uint8_t* in, uint8_t* out, int B, int H, int W, const sb_params* p,
                 const double* Mt, const double* maxCt, double* M, double* maxC, int32_t* status, cudaStream_t stream) {
    if (!p) return SB_ERR_ARG;
    if (p->method != SB_METHOD_MACENKO && p->method != SB_METHOD_VAHADANE) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::Scratch scratch(h, stream);
    sb::PipeArgs a{};
    a.in = in; a.out = out; a.B = B; a.npx = H * W;
    a.aligned = is_aligned(in, out ? out : in, a.npx);
    a.tab = h->tab;
    a.method = p->method;
    a.cluster_size = pick_cluster(h, B, a.npx, p->cluster_size);
    a.ybound = mask_ybound_f(p->luminosity_threshold);
    for (int c = 0; c < 3; ++c) a.ycoef[c] = (float)SB_RGB2LAB_COEFFS[3 + c];
    a.ang_pct = p->angular_percentile; a.lasso_lambda = p->lasso_lambda; a.conc_pct = p->conc_percentile;
    a.dl_lambda = p->dl_lambda; a.dl_iters = p->dl_iters;
    a.dl_sample_iters = p->dl_sample_iters < 0 ? 0 : p->dl_sample_iters;
    a.dl_anderson = p->dl_anderson < 0 ? 0 : (p->dl_anderson > sb::AA_MAX ? sb::AA_MAX : p->dl_anderson);
    a.Mt = Mt; a.maxCt = maxCt;
    // Vahadane re-reads the tissue mask in every dictionary pass; tiles whose per-CTA share exceeds the shared-memory
    // cache (262,144 pixels) keep it in per-call scratch instead of recomputing it
    unsigned short* mask_scratch = nullptr;
    const int groups = (a.npx + sb::GROUP_PX - 1) / sb::GROUP_PX;
    if (p->method == SB_METHOD_VAHADANE && groups / a.cluster_size > 16384)
        SB_CUDA(scratch.get(&mask_scratch, (size_t)B * groups * sizeof(unsigned short)));
    a.mask_scratch = mask_scratch;
    {
        // bracket half-width: the rule in plan_bracket (sb_pipeline.cu) unless a sweep overrides it through the environment
        static float sig = -2.f, pad = 0.f;
        if (sig < -1.5f) {
            const char* e1 = getenv("SB_BRACKET_SIGMAS"); const char* e2 = getenv("SB_BRACKET_PAD");
            sig = e1 ? (float)atof(e1) : -1.f; pad = e2 ? (float)atof(e2) : 0.f;
        }
        a.bracket_sigmas = sig; a.bracket_pad = pad;
    }
    // Macenko statistics: streaming passes over the whole batch (sb_stream.cu) for 16-byte aligned tiles of >= 32,768
    // pixels; the fused per-tile kernel otherwise, when a cluster size is requested explicitly, or with SB_NO_STREAM set.
    // Both give the same bits.
    static const bool no_stream = getenv("SB_NO_STREAM") != nullptr;
    auto run_stats = [&](const sb::PipeArgs& pa) -> cudaError_t {
        if (!no_stream && p->cluster_size == 0 && sb::stream_pipeline_eligible(pa)) return (cudaError_t)sb::launch_stream_pipeline(pa, scratch);
        cudaError_t e = (cudaError_t)sb::launch_tile_pipeline(pa, h->num_sms, stream);
        if (e == cudaSuccess) h->launches += 1;
        return e;
    };
    if (mode != sb::PIPE_NORMALIZE) {
        sb::NvtxRange nvtx(mode == sb::PIPE_FIT ? "sb_fit: statistics" : "sb_extract: statistics");
        a.mode = mode; a.M_out = M; a.maxC_out = maxC; a.status = status;
        cudaError_t e = run_stats(a);
        if (e != cudaSuccess) return cuda_fail(e, "statistics launch");
        return SB_OK;
    }
    // transform = fused per-tile statistics kernel (stain matrix + maxC of every source tile) followed by the
    // TMA-ring recombine kernel on the same stream; the statistics go through per-call scratch unless the caller asked
    // for them.
    double* ws = nullptr;
    const size_t need = (size_t)B * 8 * sizeof(double) + (size_t)B * sizeof(int32_t);
    if (!M || !maxC || !status) SB_CUDA(scratch.get(&ws, need));
    double* Mw = M ? M : ws;
    double* Cw = maxC ? maxC : ws + (size_t)B * 6;
    int32_t* Sw = status ? status : reinterpret_cast<int32_t*>(ws + (size_t)B * 8);
    a.mode = sb::PIPE_FIT; a.M_out = Mw; a.maxC_out = Cw; a.status = Sw;
    cudaError_t e;
    {
        sb::NvtxRange nvtx("sb_normalize: statistics");
        e = run_stats(a);
    }
    if (e != cudaSuccess) return cuda_fail(e, "statistics launch");
    sb::PointArgs k{};
    k.in = in; k.out = out; k.B = B; k.npx = a.npx; k.aligned = a.aligned; k.tab = h->tab; k.lasso_lambda = p->lasso_lambda;
    const bool tma = a.aligned && getenv("SB_K4_NO_TMA") == nullptr;
    {
        sb::NvtxRange nvtx("sb_normalize: K4 recombine");
        e = (cudaError_t)sb::launch_recombine_normalize(k, scratch, tma, Mw, Cw, Mt, maxCt, Sw);
    }
    if (e != cudaSuccess) return cuda_fail(e, "recombine launch");
    h->launches += 2;
    return SB_OK;
}

// Bytes of per-call scratch the entry points take for a batch (upper bound over all of them).
size_t workspace_bytes(int B, int H, int W) {
    const size_t groups = ((size_t)H * W + sb::GROUP_PX - 1) / sb::GROUP_PX;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t n = 0;
    n += up((size_t)B * groups * sizeof(unsigned short));                        // Vahadane mask bits of big tiles
    n += up((size_t)B * 8 * sizeof(double) + (size_t)B * sizeof(int32_t));      // per-tile statistics of sb_normalize
    n += up((size_t)B * 128);                                                    // per-tile constants of the ring operators
    n += up((size_t)B * 16);                                                     // HED byte sums / counters
    n += sb::stream_scratch_bytes(B, H * W);                                           // streaming statistics: per-tile state + key lists
    return n;
}

}  // namespace

extern "C" {

void sb_default_params(sb_params* p) {
    if (!p) return;
    p->method = SB_METHOD_MACENKO;
    p->luminosity_threshold = 0.8;
    p->angular_percentile = 99.0;
    p->lasso_lambda = 0.01;
    p->conc_percentile = 99.0;
    p->dl_lambda = 0.1;
    p->dl_iters = 10;
    p->cluster_size = 0;
    p->dl_sample_iters = 12;
    p->dl_anderson = 4;
}

int sb_version(void) { return 100; }

const char* sb_error_string(int code) {
    switch (code) {
        case SB_OK: return "ok";
        case SB_ERR_ARG: return "invalid argument";
        case SB_ERR_CUDA: return "CUDA error";
        case SB_ERR_UNSUPPORTED: return "unsupported request";
        case SB_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        default: return "unknown error";
    }
}
const char* sb_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
long long sb_launch_count(const sb_handle* h) { return h ? h->launches : 0; }

int sb_create(int device, sb_handle** out) {
    if (!out) return SB_ERR_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SB_ERR_NO_DEVICE;
    sb_handle probe;
    probe.device = device;
    sb::DeviceGuard guard(&probe);          // runs on `device`, restores the caller's current device on return
    if (!guard.ok) return SB_ERR_CUDA;
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return SB_ERR_NO_DEVICE;   // kernels are built for sm_100a only
    // Per-call scratch (per-tile constants, statistics) comes from the device's stream-ordered pool.  By default the pool
    // hands its memory back to the driver at every synchronisation, which makes the first launch after each sync pay a
    // fresh allocation (measured 1-60 ms); keep the memory cached instead.
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    sb_handle* h = new sb_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;

    // ---- constant tables, one allocation
    std::vector<float> od32(256), gy(768);
    for (int i = 0; i < 256; ++i) od32[i] = (float)SB_OD_TAB[i];
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 256; ++i) gy[c * 256 + i] = (float)(SB_RGB2LAB_COEFFS[3 + c] * (int)SB_GAMMA_TAB[i]);
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_od = place(256 * 4), o_gy = place(768 * 4), o_od64 = place(256 * 8), o_gam = place(256 * 2),
                 o_cb = place(3072 * 2), o_yf = place(512 * 4), o_ig = place(4096);
    SB_CUDA(cudaMalloc(&h->table_mem, off));
    char* base = (char*)h->table_mem;
    SB_CUDA(cudaMemcpy(base + o_od, od32.data(), 256 * 4, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_gy, gy.data(), 768 * 4, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_od64, SB_OD_TAB, 256 * 8, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_gam, SB_GAMMA_TAB, 256 * 2, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_cb, SB_CBRT_TAB, 3072 * 2, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_yf, SB_LAB2YF_TAB, 512 * 4, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(base + o_ig, SB_INVGAMMA_TAB, 4096, cudaMemcpyHostToDevice));
    h->tab.od = (const float*)(base + o_od);
    h->tab.gy = (const float*)(base + o_gy);
    h->tab.od64 = (const double*)(base + o_od64);
    h->tab.gamma = (const unsigned short*)(base + o_gam);
    h->tab.cbrt = (const unsigned short*)(base + o_cb);
    h->tab.lab2yf = (const int*)(base + o_yf);
    h->tab.invgamma = (const unsigned char*)(base + o_ig);
    *out = h;
    return SB_OK;
}

int sb_destroy(sb_handle* h) {
    if (!h) return SB_ERR_ARG;
    {
    sb::DeviceGuard guard(h);
    for (int i = 0; i < sb_handle::NSLOT; ++i) {
        if (h->slot_in[i]) cudaFree(h->slot_in[i]);
        if (h->slot_out[i]) cudaFree(h->slot_out[i]);
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->d_target) cudaFree(h->d_target);
    if (h->d_status) cudaFree(h->d_status);
    if (h->table_mem) cudaFree(h->table_mem);
    for (int i = 0; i < sb_handle::MAX_PASS_EVENTS; ++i) if (h->pass_ev[i]) cudaEventDestroy(h->pass_ev[i]);
    }
    delete h;
    return SB_OK;
}

int sb_tissue_mask(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold, uint8_t* mask,
                   int32_t* status, void* stream) {
    int rc = check_image_args(h, rgb, B, H, W);
    if (rc) return rc;
    if (!mask) return SB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_tissue_mask");
    sb::PointArgs a{};
    a.in = rgb; a.B = B; a.npx = H * W; a.aligned = is_aligned(rgb, mask, a.npx) && (a.npx % 16 == 0);
    a.tab = h->tab; a.ybound = mask_ybound_f(luminosity_threshold); a.mask_out = mask; a.status = status;
    for (int c = 0; c < 3; ++c) a.ycoef[c] = (float)SB_RGB2LAB_COEFFS[3 + c];
    cudaError_t e = (cudaError_t)sb::launch_mask(a, h->num_sms, st);     // presets status to EMPTY_MASK on the stream first
    if (e != cudaSuccess) return cuda_fail(e, "mask launch");
    h->launches += status ? 2 : 1;
    return SB_OK;
}

int sb_extract(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const sb_params* p, double* M, int32_t* status,
               void* stream) {
    int rc = check_image_args(h, rgb, B, H, W);
    if (rc) return rc;
    if (!M) return SB_ERR_ARG;
    return run_pipeline(h, sb::PIPE_EXTRACT, rgb, nullptr, B, H, W, p, nullptr, nullptr, M, nullptr, status, (cudaStream_t)stream);
}

int sb_fit(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const sb_params* p, double* M, double* maxC,
           int32_t* status, void* stream) {
    int rc = check_image_args(h, rgb, B, H, W);
    if (rc) return rc;
    if (!M || !maxC) return SB_ERR_ARG;
    return run_pipeline(h, sb::PIPE_FIT, rgb, nullptr, B, H, W, p, nullptr, nullptr, M, maxC, status, (cudaStream_t)stream);
}

int sb_normalize(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const sb_params* p,
                 const double* M_target, const double* maxC_target, double* M_src, double* maxC_src, int32_t* status,
                 void* stream) {
    int rc = check_image_args(h, rgb_in, B, H, W);
    if (rc) return rc;
    if (!rgb_out || !M_target || !maxC_target || rgb_in == rgb_out) return SB_ERR_ARG;
    return run_pipeline(h, sb::PIPE_NORMALIZE, rgb_in, rgb_out, B, H, W, p, M_target, maxC_target, M_src, maxC_src, status,
                        (cudaStream_t)stream);
}

static int slide_args(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double thr, sb::SlideArgs& a) {
    int rc = check_image_args(h, rgb, B, H, W);
    if (rc) return rc;
    a = sb::SlideArgs{};
    a.in = rgb; a.B = B; a.npx = H * W; a.aligned = is_aligned(rgb, rgb, a.npx);
    a.tab = h->tab; a.ybound = mask_ybound_f(thr);
    for (int c = 0; c < 3; ++c) a.ycoef[c] = (float)SB_RGB2LAB_COEFFS[3 + c];
    return SB_OK;
}

int sb_slide_grid(sb_handle* h, int B, int H, int W) {
    if (!h || B <= 0 || H <= 0 || W <= 0) return SB_ERR_ARG;
    sb::SlideArgs a{};
    a.B = B; a.npx = H * W;
    return sb::slide_grid(a, h->num_sms);
}

int sb_slide_moments(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold,
                     long long* partials, void* stream) {
    sb::SlideArgs a;
    int rc = slide_args(h, rgb, B, H, W, luminosity_threshold, a);
    if (rc) return rc;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_slide pass");
    if (!partials) return SB_ERR_ARG;
    a.sums = partials;
    cudaError_t e = (cudaError_t)sb::launch_slide_pass(a, 0, sb::slide_grid(a, h->num_sms), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "slide moments launch");
    h->launches += 1;
    return SB_OK;
}

int sb_slide_angle_hist(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold,
                        const double* V, int level, const unsigned* bins, unsigned long long* hist, void* stream) {
    sb::SlideArgs a;
    int rc = slide_args(h, rgb, B, H, W, luminosity_threshold, a);
    if (rc) return rc;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_slide pass");
    if (!V || !hist || (level != 1 && level != 2) || (level == 2 && !bins)) return SB_ERR_ARG;
    for (int k = 0; k < 6; ++k) a.V[k] = (float)V[k];
    for (int k = 0; k < 4; ++k) a.bins[k] = level == 2 ? bins[k] : 0u;
    a.hist = hist;
    cudaError_t e = (cudaError_t)sb::launch_slide_pass(a, level, sb::slide_grid(a, h->num_sms), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "slide angle histogram launch");
    h->launches += 1;
    return SB_OK;
}

int sb_slide_dl_sums(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold, const double* D,
                     double dl_lambda, int sample, long long* partials, void* stream) {
    sb::SlideArgs a;
    int rc = slide_args(h, rgb, B, H, W, luminosity_threshold, a);
    if (rc) return rc;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_slide pass");
    if (!D || !partials) return SB_ERR_ARG;
    sb::make_lasso_consts(D, dl_lambda, a.lk);
    a.sums = partials;
    cudaError_t e = (cudaError_t)sb::launch_slide_pass(a, sample ? 6 : 5, sb::slide_grid(a, h->num_sms), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "slide dictionary pass launch");
    h->launches += 1;
    return SB_OK;
}

int sb_slide_conc_hist(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const double* M, double lasso_lambda,
                       int level, const unsigned* bins, unsigned long long* hist, void* stream) {
    sb::SlideArgs a;
    int rc = slide_args(h, rgb, B, H, W, 0.8, a);
    if (rc) return rc;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_slide pass");
    if (!M || !hist || (level != 1 && level != 2) || (level == 2 && !bins)) return SB_ERR_ARG;
    sb::make_lasso_consts(M, lasso_lambda, a.lk);
    for (int k = 0; k < 4; ++k) a.bins[k] = level == 2 ? bins[k] : 0u;
    a.hist = hist;
    cudaError_t e = (cudaError_t)sb::launch_slide_pass(a, 2 + level, sb::slide_grid(a, h->num_sms), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "slide concentration histogram launch");
    h->launches += 1;
    return SB_OK;
}

int sb_normalize_host(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const sb_params* p,
                      const double* M_target, const double* maxC_target, int32_t* status, int chunk_tiles) {
    int rc = check_image_args(h, rgb_in, B, H, W);
    if (rc) return rc;
    if (!rgb_out || !M_target || !maxC_target || !p) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_normalize_host");
    const size_t tile_bytes = (size_t)H * W * 3;
    if (chunk_tiles <= 0) {
        // 48 MB chunks: link-rate copies, and enough tiles per launch (64 of 512x512) that the ~10-35 launches of the
        // streaming statistics passes stay far below the chunk's copy time (a 12 MB chunk left Vahadane launch-bound)
        chunk_tiles = (int)(((size_t)48 << 20) / tile_bytes);
        if (chunk_tiles < 1) chunk_tiles = 1;
    }
    if (chunk_tiles > B) chunk_tiles = B;
    const size_t need = (size_t)chunk_tiles * tile_bytes;
    if (!h->s_in) {
        SB_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        SB_CUDA(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        SB_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < sb_handle::NSLOT; ++i) {
            SB_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            SB_CUDA(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
            SB_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
        SB_CUDA(cudaMalloc(&h->d_target, 8 * sizeof(double)));
    }
    if (h->slot_bytes < need) {
        for (int i = 0; i < sb_handle::NSLOT; ++i) {
            if (h->slot_in[i]) cudaFree(h->slot_in[i]);
            if (h->slot_out[i]) cudaFree(h->slot_out[i]);
            SB_CUDA(cudaMalloc(&h->slot_in[i], need));
            SB_CUDA(cudaMalloc(&h->slot_out[i], need));
        }
        h->slot_bytes = need;
    }
    if (h->status_cap < (size_t)B) {
        if (h->d_status) cudaFree(h->d_status);
        SB_CUDA(cudaMalloc(&h->d_status, (size_t)B * 4));
        h->status_cap = (size_t)B;
    }
    double tgt[8];
    std::memcpy(tgt, M_target, 6 * sizeof(double));
    std::memcpy(tgt + 6, maxC_target, 2 * sizeof(double));
    SB_CUDA(cudaMemcpyAsync(h->d_target, tgt, sizeof(tgt), cudaMemcpyHostToDevice, h->s_comp));
    SB_CUDA(cudaStreamSynchronize(h->s_comp));

    const int nchunks = (B + chunk_tiles - 1) / chunk_tiles;
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c % sb_handle::NSLOT;
        const int t0 = c * chunk_tiles, nt = (B - t0 < chunk_tiles) ? (B - t0) : chunk_tiles;
        const size_t bytes = (size_t)nt * tile_bytes;
        // the slot's previous occupant must have been consumed by compute (input) and copied out (output)
        if (c >= sb_handle::NSLOT) {
            SB_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_comp[slot], 0));
            SB_CUDA(cudaStreamWaitEvent(h->s_comp, h->ev_out[slot], 0));
        }
        SB_CUDA(cudaMemcpyAsync(h->slot_in[slot], rgb_in + (size_t)t0 * tile_bytes, bytes, cudaMemcpyHostToDevice, h->s_in));
        SB_CUDA(cudaEventRecord(h->ev_in[slot], h->s_in));
        SB_CUDA(cudaStreamWaitEvent(h->s_comp, h->ev_in[slot], 0));
        rc = run_pipeline(h, sb::PIPE_NORMALIZE, h->slot_in[slot], h->slot_out[slot], nt, H, W, p, h->d_target, h->d_target + 6,
                          nullptr, nullptr, h->d_status + t0, h->s_comp);
        if (rc) return rc;
        SB_CUDA(cudaEventRecord(h->ev_comp[slot], h->s_comp));
        SB_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0));
        SB_CUDA(cudaMemcpyAsync(rgb_out + (size_t)t0 * tile_bytes, h->slot_out[slot], bytes, cudaMemcpyDeviceToHost, h->s_out));
        SB_CUDA(cudaEventRecord(h->ev_out[slot], h->s_out));
    }
    SB_CUDA(cudaStreamSynchronize(h->s_out));
    if (status) {
        SB_CUDA(cudaMemcpyAsync(status, h->d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, h->s_out));
        SB_CUDA(cudaStreamSynchronize(h->s_out));
    }
    SB_CUDA(cudaStreamSynchronize(h->s_comp));
    return SB_OK;
}

int sb_concentrations(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const double* M, double lasso_lambda,
                      float* C, void* stream) {
    int rc = check_image_args(h, rgb, B, H, W);
    if (rc) return rc;
    if (!M || !C) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_concentrations");
    sb::PointArgs a{};
    a.in = rgb; a.B = B; a.npx = H * W; a.aligned = is_aligned(rgb, rgb, a.npx);
    a.tab = h->tab; a.M = M; a.lasso_lambda = lasso_lambda; a.conc_out = C;
    // 16-byte aligned tiles: a pass on the TMA ring with coalesced stores; register-staged kernel otherwise (same bits)
    cudaError_t e = (cudaError_t)(sb::concentrations_stream_eligible(a) ? sb::launch_concentrations_stream(a, h->num_sms, (cudaStream_t)stream)
                                                                       : sb::launch_concentrations(a, h->num_sms, (cudaStream_t)stream));
    if (e != cudaSuccess) return cuda_fail(e, "concentrations launch");
    h->launches += 1;
    return SB_OK;
}

int sb_recombine(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* M_src,
                 const double* scale, const double* M_target, double lasso_lambda, void* stream) {
    int rc = check_image_args(h, rgb_in, B, H, W);
    if (rc) return rc;
    if (!rgb_out || !M_src || !scale || !M_target) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_recombine (K4)");
    sb::Scratch scratch(h, (cudaStream_t)stream);
    sb::PointArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = is_aligned(rgb_in, rgb_out, a.npx);
    a.tab = h->tab; a.M = M_src; a.scale = scale; a.Mt = M_target; a.lasso_lambda = lasso_lambda;
    // TMA-staged ring when every tile is a whole number of 16-byte vectors; register-staged kernel otherwise
    const bool tma = a.aligned && getenv("SB_K4_NO_TMA") == nullptr;
    a.debug_copy = getenv("SB_K4_COPY") != nullptr;
    cudaError_t e = (cudaError_t)sb::launch_recombine(a, scratch, tma);
    if (e != cudaSuccess) return cuda_fail(e, "recombine launch");
    h->launches += 2;
    return SB_OK;
}

int sb_stain_augment(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* M,
                     const double* alpha, const double* beta, int augment_background, double luminosity_threshold,
                     double lasso_lambda, void* stream) {
    int rc = check_image_args(h, rgb_in, B, H, W);
    if (rc) return rc;
    if (!rgb_out || !M || !alpha || !beta) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_stain_augment");
    sb::Scratch scratch(h, (cudaStream_t)stream);
    sb::PointArgs a{};
    a.in = rgb_in; a.out = rgb_out; a.B = B; a.npx = H * W; a.aligned = is_aligned(rgb_in, rgb_out, a.npx);
    a.tab = h->tab; a.M = M; a.scale = alpha; a.beta = beta; a.lasso_lambda = lasso_lambda;
    a.augment_background = augment_background; a.ybound = mask_ybound_f(luminosity_threshold);
    for (int c = 0; c < 3; ++c) a.ycoef[c] = (float)SB_RGB2LAB_COEFFS[3 + c];
    cudaError_t e = (cudaError_t)sb::launch_stain_augment(a, scratch);
    if (e != cudaSuccess) return cuda_fail(e, "stain_augment launch");
    h->launches += 1;
    return SB_OK;
}

int sb_set_pass_timing(sb_handle* h, int enable) {
    if (!h) return SB_ERR_ARG;
    h->pass_timing = enable != 0;
    h->n_pass_ev = 0;
    return SB_OK;
}

int sb_get_pass_timing(sb_handle* h, int max_passes, float* ms, char* names, int names_bytes) {
    if (!h || !ms || max_passes <= 0) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    int n = 0;
    std::string all;
    for (int i = 0; i + 1 < h->n_pass_ev && n < max_passes; ++i) {
        if (cudaEventSynchronize(h->pass_ev[i + 1]) != cudaSuccess) return SB_ERR_CUDA;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->pass_ev[i], h->pass_ev[i + 1]) != cudaSuccess) return SB_ERR_CUDA;
        ms[n++] = t;
        all += h->pass_name[i];
        all += '\n';
    }
    if (names && names_bytes > 0) {
        std::strncpy(names, all.c_str(), (size_t)names_bytes - 1);
        names[names_bytes - 1] = 0;
    }
    return n;
}

int sb_stream_fallbacks(sb_handle* h, unsigned* counters, int reset) {
    if (!h || !counters) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    cudaError_t e = (cudaError_t)sb::stream_fallback_counters(counters, reset != 0);      // synchronises with the device
    if (e != cudaSuccess) return cuda_fail(e, "fallback counters");
    return SB_OK;
}

size_t sb_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return workspace_bytes(B, H, W);
}

int sb_set_workspace(sb_handle* h, void* device_mem, size_t bytes) {
    if (!h || (device_mem == nullptr) != (bytes == 0)) return SB_ERR_ARG;
    if (h->scratch_depth != 0) return SB_ERR_ARG;
    h->user_ws = static_cast<unsigned char*>(device_mem);
    h->user_ws_bytes = bytes;
    h->user_ws_off = 0;
    return SB_OK;
}

int sb_rgb_to_od(sb_handle* h, const uint8_t* rgb, size_t n_values, void* od, int out_f32, void* stream) {
    if (!h || !rgb || !od || n_values == 0) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_rgb_to_od");
    cudaError_t e = (cudaError_t)sb::launch_rgb_to_od(rgb, od, n_values, out_f32, h->tab.od64, h->num_sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "rgb_to_od launch");
    h->launches += 1;
    return SB_OK;
}

int sb_od_to_rgb(sb_handle* h, const void* od, int in_f32, size_t n_values, uint8_t* rgb, int32_t* negative, void* stream) {
    if (!h || !od || !rgb || n_values == 0) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_od_to_rgb");
    cudaError_t e = (cudaError_t)sb::launch_od_to_rgb(od, rgb, n_values, in_f32, negative, h->num_sms, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "od_to_rgb launch");
    h->launches += 1;
    return SB_OK;
}

}  // extern "C"
