// sb_kernels.h -- internal interface between the C-ABI layer (sb_api.cu) and the kernel translation units.
#pragma once
#include <atomic>
#include <nvtx3/nvToolsExt.h>     // header-only NVTX v3: ranges are no-ops unless a profiler is attached
#include "sb_device.cuh"
#include "../../include/stainb200.h"

// The opaque handle of include/stainb200.h: device tables + host-streaming state.
struct sb_handle {
    int device = 0;
    int num_sms = 0;
    long long launches = 0;
    sb::Tables tab{};
    void* table_mem = nullptr;
    // host-streaming state (sb_normalize_host)
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    static constexpr int NSLOT = 3;
    uint8_t* slot_in[NSLOT] = {nullptr, nullptr, nullptr};
    uint8_t* slot_out[NSLOT] = {nullptr, nullptr, nullptr};
    size_t slot_bytes = 0;
    cudaEvent_t ev_in[NSLOT] = {}, ev_comp[NSLOT] = {}, ev_out[NSLOT] = {};
    double* d_target = nullptr;   // [6 + 2]
    int32_t* d_status = nullptr;
    size_t status_cap = 0;
    // caller-lent per-call scratch (sb_set_workspace); null = the device's stream-ordered pool
    unsigned char* user_ws = nullptr;
    size_t user_ws_bytes = 0, user_ws_off = 0;
    int scratch_depth = 0;
    // optional per-pass timing of the streaming statistics (sb_set_pass_timing): CUDA events on the launching stream
    static constexpr int MAX_PASS_EVENTS = 48;
    bool pass_timing = false;
    cudaEvent_t pass_ev[MAX_PASS_EVENTS] = {};
    const char* pass_name[MAX_PASS_EVENTS] = {};
    int n_pass_ev = 0;
};

// ---- host-side helpers shared by the translation units that hold entry points / launchers
namespace sb {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: one bit per device ordinal per kernel.
struct DeviceOnce { std::atomic<unsigned long long> bits{0ull}; };
template <class Kern>
inline cudaError_t ensure_dyn_smem(DeviceOnce& once, Kern kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (dev < 64 && (once.bits.load(std::memory_order_acquire) & bit)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) once.bits.fetch_or(bit, std::memory_order_release);
    return e;
}

// Every entry point runs on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(const sb_handle* h) {
        if (!h) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; ok = false; return; }
        if (prev != h->device) ok = cudaSetDevice(h->device) == cudaSuccess; else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// NVTX range around the launches of one entry point (tracing).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Per-call device scratch (per-tile constants, statistics, mask bits).  Served from the workspace the caller lent with
// sb_set_workspace when it fits (bump allocation, reset when the outermost call returns: calls through one handle are
// stream-ordered on one stream at a time, so the next call may reuse the bytes), else from the device's stream-ordered
// memory pool (cudaMallocAsync / cudaFreeAsync on the call's stream -- no synchronisation either way).
struct Scratch {
    sb_handle* h;
    cudaStream_t st;
    void* pooled[16];
    int n = 0;
    Scratch(sb_handle* hh, cudaStream_t s) : h(hh), st(s) { h->scratch_depth += 1; }
    cudaError_t get(void** p, size_t bytes) {
        const size_t need = (bytes + 255) & ~(size_t)255;
        if (h->user_ws && h->user_ws_off + need <= h->user_ws_bytes) {
            *p = h->user_ws + h->user_ws_off;
            h->user_ws_off += need;
            return cudaSuccess;
        }
        if (n >= 16) return cudaErrorMemoryAllocation;
        cudaError_t e = cudaMallocAsync(p, need, st);
        if (e == cudaSuccess) pooled[n++] = *p;
        return e;
    }
    template <class T> cudaError_t get(T** p, size_t bytes) { void* v = nullptr; cudaError_t e = get(&v, bytes); *p = static_cast<T*>(v); return e; }
    ~Scratch() {
        for (int i = 0; i < n; ++i) cudaFreeAsync(pooled[i], st);
        if (--h->scratch_depth == 0) h->user_ws_off = 0;
    }
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
};

// Optional timing of the passes (sb_set_pass_timing): an event in front of every launch and one behind the last, on the
// launching stream; sb_get_pass_timing reads the differences.  Off by default: nothing is recorded.
struct PassTimer {
    sb_handle* h;
    cudaStream_t st;
    void mark(const char* name) {
        if (!h->pass_timing || h->n_pass_ev >= sb_handle::MAX_PASS_EVENTS) return;
        if (!h->pass_ev[h->n_pass_ev] && cudaEventCreate(&h->pass_ev[h->n_pass_ev]) != cudaSuccess) return;
        cudaEventRecord(h->pass_ev[h->n_pass_ev], st);
        h->pass_name[h->n_pass_ev++] = name;
    }
};

}  // namespace sb

namespace sb {

enum PipeMode { PIPE_EXTRACT = 0, PIPE_FIT = 1, PIPE_NORMALIZE = 2 };

struct PipeArgs {
    const uint8_t* in;
    uint8_t* out;
    int B, npx;
    int aligned;          // tile base addresses and tile size are multiples of 16 bytes
    Tables tab;
    int mode, method, cluster_size;
    float ybound;         // tissue <=> sum_c ycoef[c] * gamma[v_c] < ybound
    float ycoef[3];       // the Y row of cv2's fixed-point RGB->XYZ matrix (871, 2929, 296)
    double ang_pct, lasso_lambda, conc_pct, dl_lambda;
    int dl_iters, dl_sample_iters, dl_anderson;
    const double* Mt;     // [2,3] device
    const double* maxCt;  // [2]   device
    double* M_out;        // [B,2,3] or null
    double* maxC_out;     // [B,2] or null
    int32_t* status;      // [B] or null
    unsigned short* mask_scratch;   // Vahadane, tiles too big for the shared-memory mask cache: [B][groups] or null
    float bracket_sigmas, bracket_pad;   // half-width of the sampled rank brackets: sigmas * binomial sigma + pad (sample ranks)
    // optional indirection (fallback of the streaming path): process tiles tile_list[0 .. *tile_count) instead of 0 .. B
    const int* tile_list;
    const int* tile_count;
};

int launch_tile_pipeline(const PipeArgs& a, int num_sms, cudaStream_t stream);
// Streaming (one launch per pass over the whole batch) Macenko statistics: sb_stream.cu.  Same outputs as the fused kernel.
bool stream_pipeline_eligible(const PipeArgs& a);
size_t stream_scratch_bytes(int B, int npx);
int launch_stream_pipeline(const PipeArgs& a, Scratch& scratch);
int stream_fallback_counters(unsigned out[8], bool reset);

// ---- slide-level fit passes (sb_pipeline.cu): 0 moments, 1/2 angle histograms (level 1/2), 3/4 concentration histograms
struct SlideArgs {
    const uint8_t* in;
    int B, npx, aligned;
    Tables tab;
    float ybound, ycoef[3];
    float V[6];                 // projection plane (passes 1, 2): rows = the two leading eigenvectors
    LassoK lk;                  // passes 3, 4
    unsigned bins[4];           // level-1 bins refined by passes 2 and 4
    long long* sums;            // passes 0, 5, 6: [grid][10] per-CTA partials, fixed point (2^32 moments, 2^30 dictionary sums); count last
    unsigned long long* hist;   // passes 1-4: [8192] counters, accumulated (+=)
};
int slide_grid(const SlideArgs& a, int num_sms);
int launch_slide_pass(const SlideArgs& a, int pass, int grid, cudaStream_t stream);

// ---- Reinhard / luminosity standardiser as streaming passes on the TMA ring (sb_reinhard.cu); mode = LabMode 0..2 of sb_colour.cu
bool reinhard_ring_eligible(const void* in, const void* out, int npx);
int launch_reinhard_ring(sb_handle* h, const uint8_t* in, uint8_t* out, int B, int npx, int mode, int skip_brightness, const double* tmeans,
                         const double* tstds, double* means_out, double* stds_out, int mask_background, int lmax, double percentile,
                         int32_t* status, cudaStream_t st);

// ---- pointwise kernels (sb_pointwise.cu)
struct PointArgs {
    const uint8_t* in;
    uint8_t* out;
    int B, npx, aligned;
    Tables tab;
    float ybound;
    float ycoef[3];        // Y row of cv2's RGB->XYZ matrix (stain_augment's ring path computes the mask from gamma values)
    const double* M;       // [B,2,3] per-tile source matrices
    const double* scale;   // [B,2]   (recombine)  /  alpha (augment)
    const double* beta;    // [B,2]   (augment)
    const double* Mt;      // [2,3]   (recombine)
    double lasso_lambda;
    int augment_background;
    int debug_copy;        // development: K4 ring moves bytes without computing
    float* conc_out;       // [B,N,2] (concentrations)
    uint8_t* mask_out;     // [B,N]   (mask)
    int32_t* status;
};
int launch_mask(const PointArgs& a, int num_sms, cudaStream_t stream);
int launch_mask_stream(const PointArgs& a, int num_sms, cudaStream_t stream);      // sb_stream.cu: the mask as a pass on the TMA ring
int launch_recombine(const PointArgs& a, Scratch& scratch, bool use_tma);
int launch_recombine_normalize(const PointArgs& a, Scratch& scratch, bool use_tma, const double* M_src,
                               const double* maxC_src, const double* Mt, const double* maxCt, int32_t* status);
int launch_stain_augment(const PointArgs& a, Scratch& scratch);
int launch_concentrations(const PointArgs& a, int num_sms, cudaStream_t stream);
bool concentrations_stream_eligible(const PointArgs& a);                             // sb_stream.cu: concentrations as a pass on the TMA ring
int launch_concentrations_stream(const PointArgs& a, int num_sms, cudaStream_t stream);
int launch_rgb_to_od(const uint8_t* in, void* out, size_t n, int f32, const double* od64, int num_sms, cudaStream_t stream);
int launch_od_to_rgb(const void* od, uint8_t* out, size_t n, int f32, int* negative, int num_sms, cudaStream_t stream);

}  // namespace sb
