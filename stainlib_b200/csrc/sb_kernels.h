// sb_kernels.h -- internal interface between the C-ABI layer (sb_api.cu) and the kernel translation units.
#pragma once
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
};

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
};

int launch_tile_pipeline(const PipeArgs& a, int num_sms, cudaStream_t stream);

// ---- slide-level fit passes (sb_pipeline.cu): 0 moments, 1/2 angle histograms (level 1/2), 3/4 concentration histograms
struct SlideArgs {
    const uint8_t* in;
    int B, npx, aligned;
    Tables tab;
    float ybound, ycoef[3];
    float V[6];                 // projection plane (passes 1, 2): rows = the two leading eigenvectors
    LassoK lk;                  // passes 3, 4
    unsigned bins[4];           // level-1 bins refined by passes 2 and 4
    double* sums;               // pass 0: [grid][10] per-CTA partials (n last)
    unsigned long long* hist;   // passes 1-4: [8192] counters, accumulated (+=)
};
int slide_grid(const SlideArgs& a, int num_sms);
int launch_slide_pass(const SlideArgs& a, int pass, int grid, cudaStream_t stream);

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
int launch_recombine(const PointArgs& a, int num_sms, cudaStream_t stream, bool use_tma);
int launch_recombine_normalize(const PointArgs& a, int num_sms, cudaStream_t stream, bool use_tma, const double* M_src,
                               const double* maxC_src, const double* Mt, const double* maxCt, int32_t* status);
int launch_stain_augment(const PointArgs& a, int num_sms, cudaStream_t stream);
int launch_concentrations(const PointArgs& a, int num_sms, cudaStream_t stream);

}  // namespace sb
