/* stainb200.h -- C ABI of libstainb200.so: the B200 (sm_100a) stain-processing hot path.
 *
 * The reference (sebastianffx/stainlib) has no FFI: its boundary is the Python class API of
 * stainlib/extraction, stainlib/normalization and stainlib/augmentation.  Each entry point below replaces the body of
 * one of those reference functions (cited per function, paths relative to the reference checkout); the Python mirror
 * in stainlib_b200/ binds them with ctypes (stainlib_b200/_native.py) and keeps the reference's class/method names.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary; every function returns 0 on success or a negative
 *     sb_error code (sb_error_string() describes it; CUDA failures map to SB_ERR_CUDA, text via sb_last_cuda_error()).
 *   - all image / matrix pointers are DEVICE pointers unless the function name ends in _host.
 *   - images are uint8, C-contiguous [B,H,W,3] RGB; "tile" = one [H,W,3] image; N = H*W.
 *   - stain matrices are double [B,2,3] (rows = stains, H first), maxC / scale / alpha / beta are double [B,2].
 *   - work is stream-ordered on `stream` (a cudaStream_t passed as void*); nothing synchronises unless stated.
 *   - the caller owns every image / matrix / status buffer.  The library allocates device memory in sb_create (constant
 *     tables), in the *_host entry points (staging slots, kept in the handle) and, per call, a few hundred bytes per tile
 *     of scratch (per-tile constants and statistics; 2 bytes per 16 pixels more for Vahadane on tiles > 512x512).  That
 *     scratch comes from the workspace lent with sb_set_workspace() when one is set and large enough
 *     (sb_workspace_bytes() says how much a batch needs), else from the device's stream-ordered memory pool
 *     (cudaMallocAsync / cudaFreeAsync on the call's stream: no synchronisation, no cudaMalloc on the hot path).
 *   - every entry point runs on the device of its handle and restores the caller's current device before returning;
 *     a process may hold one handle per device.
 *   - a tile's outputs do not depend on the batch it is in, on how the batch is chunked or sharded, or on the
 *     cluster_size the launcher picks: per-tile sums are fixed-point integers (SURVEY section 8-e determinism).
 *   - per-tile int32 status word instead of raising mid-batch (bits below); outputs of flagged tiles are defined
 *     as documented per function.
 */
#ifndef STAINB200_H
#define STAINB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb_handle sb_handle;

enum sb_error {
    SB_OK = 0,
    SB_ERR_ARG = -1,        /* null pointer, non-positive shape, unknown method ...            */
    SB_ERR_CUDA = -2,       /* a CUDA runtime call failed; see sb_last_cuda_error()             */
    SB_ERR_UNSUPPORTED = -3,/* valid request the kernels do not implement (e.g. N > 2^24)       */
    SB_ERR_NO_DEVICE = -4   /* no CUDA device / not an sm_100 part                              */
};

/* per-tile status bits */
#define SB_STATUS_EMPTY_MASK   1  /* no tissue pixel: reference raises TissueMaskException (stain_utils.py:46-47)   */
#define SB_STATUS_FEW_TISSUE   2  /* < 2 tissue pixels: np.cov is NaN, reference raises LinAlgError                 */
#define SB_STATUS_ZERO_MAXC    4  /* a 99th-percentile concentration is 0 / non-finite: reference divides by zero   */
#define SB_STATUS_DEGENERATE   8  /* non-finite stain matrix (e.g. zero covariance)                                 */

enum sb_method { SB_METHOD_MACENKO = 0, SB_METHOD_VAHADANE = 1 };

/* Parameters shared by the extract / fit / normalize entry points.  Defaults = the reference's keyword defaults. */
typedef struct sb_params {
    int    method;               /* sb_method                                                                      */
    double luminosity_threshold; /* 0.8   stain_utils.py:32                                                        */
    double angular_percentile;   /* 99    macenko_stain_extractor.py:7                                             */
    double lasso_lambda;         /* 0.01  stain_utils.py:69 (get_concentrations regularizer)                       */
    double conc_percentile;      /* 99    normalizer.py:36,47                                                      */
    double dl_lambda;            /* 0.1   vahadane_stain_extractor.py:19 (trainDL lambda1)                         */
    int    dl_iters;             /* 10    at most this many full-batch dictionary-learning passes (they stop at a     */
                                 /*       residual of 2e-6; the reference runs for 1 s of wall-clock time)           */
    int    cluster_size;         /* CTAs cooperating on one tile: 0 = auto, else 1/2/4/8                            */
    int    dl_sample_iters;      /* 12    warm-start passes over a 1-in-16 sample of the tile before the full      */
                                 /*       passes (skipped, and 4 full passes added, when the sample has < 1024     */
                                 /*       tissue pixels); 0 = none                                                  */
    int    dl_anderson;          /* 4     Anderson-acceleration memory of the dictionary fixed-point iteration;    */
                                 /*       0 = plain alternating minimisation                                        */
} sb_params;

void sb_default_params(sb_params* p);

int  sb_create(int device, sb_handle** out);
int  sb_destroy(sb_handle* h);
const char* sb_error_string(int code);
const char* sb_last_cuda_error(void);
int  sb_version(void);
/* number of kernel launches issued through this handle since creation (bench.py's gpu_launches) */
long long sb_launch_count(const sb_handle* h);

/* Diagnostics of the streaming statistics passes (csrc/sb_stream.cu): counters[0] = tiles that left the streaming passes
 * for the fused per-tile kernel since the last reset, counters[1..7] = by reason (too small / too little tissue, sample too
 * small (angle, concentration), bracket at the edge of the key range, bracket missed or list overflowed (angle,
 * concentration), non-unit stain vectors).  Results are the same bits either way; the counters only say which path
 * produced them.  Synchronises with the device.  counters: HOST unsigned[8]. */
int sb_stream_fallbacks(sb_handle* h, unsigned* counters, int reset);

/* Per-pass timing of the streaming statistics (what bench.py's roofline records are made of): with timing enabled, the
 * next sb_extract / sb_fit / sb_normalize records a CUDA event in front of every launch of its statistics passes on the
 * caller's stream; sb_get_pass_timing waits for them and returns the number of passes of the LAST call, their
 * durations (ms) and their names ('\n'-separated).  Off by default: no events, no synchronisation. */
int sb_set_pass_timing(sb_handle* h, int enable);
int sb_get_pass_timing(sb_handle* h, int max_passes, float* ms, char* names, int names_bytes);

/* Caller-owned scratch (SURVEY section 8-b "ownership").  sb_workspace_bytes: upper bound of the per-call scratch any
 * entry point takes for a [B,H,W,3] batch.  sb_set_workspace lends `bytes` of DEVICE memory to the handle (NULL, 0
 * takes it back); calls whose scratch fits use it instead of the stream-ordered pool.  The lender must keep it alive
 * until the last call that used it has finished on its stream, and must not run calls of one handle on two streams at
 * once while a workspace is set. */
size_t sb_workspace_bytes(int B, int H, int W);
int    sb_set_workspace(sb_handle* h, void* device_mem, size_t bytes);

/* convert_RGB_to_OD -- stainlib/utils/stain_utils.py:101-112: od[i] = max(-ln(max(rgb[i], 1) / 255), 1e-6) for n_values
 * uint8 values (any shape, flattened); od: device double [n_values] (float when out_f32 != 0). */
int sb_rgb_to_od(sb_handle* h, const uint8_t* rgb, size_t n_values, void* od, int out_f32, void* stream);

/* convert_OD_to_RGB -- stainlib/utils/stain_utils.py:114-124: rgb[i] = uint8(255 * exp(-max(od[i], 1e-6))), truncated.
 * od: device double [n_values] (float when in_f32 != 0).  negative (device int32, may be NULL; zero it first) is set to 1
 * when some od[i] < 0, where the reference asserts. */
int sb_od_to_rgb(sb_handle* h, const void* od, int in_f32, size_t n_values, uint8_t* rgb, int32_t* negative, void* stream);

/* LuminosityThresholdTissueLocator.get_tissue_mask -- stainlib/utils/stain_utils.py:32-48.
 * mask: uint8 [B,H,W] (1 = tissue).  status bit SB_STATUS_EMPTY_MASK where the reference would raise. */
int sb_tissue_mask(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold,
                   uint8_t* mask, int32_t* status, void* stream);

/* MacenkoStainExtractor.get_stain_matrix -- stainlib/extraction/macenko_stain_extractor.py:7-44  (method MACENKO)
 * VahadaneStainExtractor.get_stain_matrix -- stainlib/extraction/vahadane_stain_extractor.py:19-43 (method VAHADANE)
 * M: double [B,2,3].  Flagged tiles get NaN matrices. */
int sb_extract(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const sb_params* p,
               double* M, int32_t* status, void* stream);

/* ExtractiveStainNormalizer.fit -- stainlib/normalization/normalizer.py:27-36, batched over tiles:
 * M = get_stain_matrix(tile); maxC = percentile(get_concentrations(tile, M), 99, axis=0). */
int sb_fit(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const sb_params* p,
           double* M, double* maxC, int32_t* status, void* stream);

/* ExtractiveStainNormalizer.transform -- stainlib/normalization/normalizer.py:39-50, one fused pass sequence per
 * tile.  M_target double[2,3], maxC_target double[2] live on the DEVICE (outputs of sb_fit).  M_src / maxC_src are
 * optional outputs (may be NULL).  Output is NOT clipped (wraps modulo 256 like the reference's astype(uint8)).
 * Flagged tiles: EMPTY_MASK / FEW_TISSUE / DEGENERATE -> input copied through; ZERO_MAXC -> zeros (as the reference). */
int sb_normalize(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const sb_params* p,
                 const double* M_target, const double* maxC_target, double* M_src, double* maxC_src,
                 int32_t* status, void* stream);

/* Slide-level fit: ExtractiveStainNormalizer.fit (normalizer.py:27-36) of a SET of tiles treated as ONE image (the
 * target slide), possibly sharded over ranks.  Each pass produces statistics that ADD across launches and ranks; the
 * caller all-reduces them and does the small serial steps (covariance + eigenvectors, rank selection in the
 * histograms, stain matrix) on the host between passes -- stainlib_b200/normalization/slide_fit.py is that host side.
 *   sb_slide_moments     masked OD moments: partials int64 [sb_slide_grid()][10] = (sum od[3], sum od x od[6] (00,01,02,
 *                        11,12,22), n) per CTA, the nine sums in FIXED POINT (scale 2^32; n is a plain count): add the rows,
 *                        all-reduce them as integers, divide by 2^32 -- the totals then do not depend on the sharding.
 *   sb_slide_angle_hist  level 1: hist[4096] += counts of the top 12 bits of the 23-bit angle key of every tissue pixel,
 *                        V (HOST double[6]) = the two leading eigenvectors (rows); level 2: hist[q*2048 + low 11 bits]
 *                        += for keys whose level-1 bin is bins[q] (HOST unsigned[4]).
 *   sb_slide_conc_hist   the same for the concentrations of ALL pixels under M (HOST double[6]): level 1 fills
 *                        hist[0..4095] (stain 0) and hist[4096..8191] (stain 1); level 2 refines bins[0], bins[1] of
 *                        stain 0 and bins[2], bins[3] of stain 1.
 * hist: device unsigned long long [8192], accumulated (zero it before the first launch of a pass).  Keys map back to
 * values as in csrc/sb_device.cuh (angle_from_key, conc_from_key). */
int sb_slide_grid(sb_handle* h, int B, int H, int W);
int sb_slide_moments(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold,
                     long long* partials, void* stream);
int sb_slide_angle_hist(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold,
                        const double* V, int level, const unsigned* bins, unsigned long long* hist, void* stream);
/*   sb_slide_dl_sums     one Vahadane dictionary pass: sparse codes of the tissue pixels (sample != 0: of the 1-in-16 sample
 *                        groups of every tile) under the dictionary D (HOST double[6], rows = atoms); partials int64
 *                        [sb_slide_grid()][10] = (sum a a^T (00,01,11), sum x a_0 [3], sum x a_1 [3], pixel count), the nine
 *                        sums in fixed point, scale 2^30. */
int sb_slide_dl_sums(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double luminosity_threshold, const double* D,
                     double dl_lambda, int sample, long long* partials, void* stream);
int sb_slide_conc_hist(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const double* M, double lasso_lambda,
                       int level, const unsigned* bins, unsigned long long* hist, void* stream);

/* Same as sb_normalize but rgb_in / rgb_out / status are HOST buffers (pinned for full speed): the call streams
 * tile chunks H2D -> kernels -> D2H on internal streams with double buffering and returns after the last byte has
 * landed in rgb_out (synchronous).  M_target / maxC_target are HOST doubles here. */
int sb_normalize_host(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const sb_params* p,
                      const double* M_target, const double* maxC_target, int32_t* status, int chunk_tiles);

/* get_concentrations -- stainlib/utils/stain_utils.py:69-78.  C: float [B,N,2]. */
int sb_concentrations(sb_handle* h, const uint8_t* rgb, int B, int H, int W, const double* M, double lasso_lambda,
                      float* C, void* stream);

/* Lines normalizer.py:46,48-50 alone (the fused OD+recombine kernel): out = uint8(255*exp(-(C*scale) @ M_target)),
 * C = get_concentrations(tile, M_src[b]).  M_src [B,2,3], scale [B,2], M_target [2,3], all device doubles. */
int sb_recombine(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* M_src,
                 const double* scale, const double* M_target, double lasso_lambda, void* stream);

/* StainAugmentor.pop -- stainlib/augmentation/augmenter.py:428-449 for given draws: concentrations under M[b],
 * alpha*C+beta on tissue pixels (all pixels if augment_background), recombined with the tile's own M, clipped. */
int sb_stain_augment(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* M,
                     const double* alpha, const double* beta, int augment_background, double luminosity_threshold,
                     double lasso_lambda, void* stream);

/* ReinhardStainNormalizer.fit -- normalizer.py:64-68 (+ standardize_brightness stain_utils.py:188-194, lab_split
 * :146-158, get_mean_std :174-186).  means/stds: double [B,3] of the brightness-standardised tile in LAB. */
int sb_reinhard_stats(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double* means, double* stds,
                      void* stream);

/* The exported pieces of the Reinhard path, each as its own entry point:
 *   sb_standardize_brightness  standardize_brightness -- stain_utils.py:188-194: uint8(clip(I * 255 / percentile(I, 90), 0, 255)) per tile.
 *   sb_lab_mean_std            get_mean_std -- stain_utils.py:174-186: mean / population std of lab_split's three planes (no
 *                              brightness standardisation); means, stds: double [B,3].
 *   sb_lab_split               lab_split -- stain_utils.py:146-158: 8-bit RGB -> LAB, I1 = float32(L) / 2.55, I2 = a - 128,
 *                              I3 = b - 128; three device float planes of n_pixels each.
 *   sb_lab_merge               merge_back -- stain_utils.py:160-172: uint8(clip((I1 * 2.55, I2 + 128, I3 + 128), 0, 255)) -> LAB2RGB;
 *                              planes are device float (double when is_f64) and are NOT modified (the reference scales its
 *                              arguments in place; the Python mirror reproduces that on the caller's arrays). */
int sb_standardize_brightness(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, void* stream);
int sb_lab_mean_std(sb_handle* h, const uint8_t* rgb, int B, int H, int W, double* means, double* stds, void* stream);
int sb_lab_split(sb_handle* h, const uint8_t* rgb, size_t n_pixels, float* I1, float* I2, float* I3, void* stream);
int sb_lab_merge(sb_handle* h, const void* I1, const void* I2, const void* I3, int is_f64, size_t n_pixels, uint8_t* rgb,
                 void* stream);

/* ReinhardStainNormalizer.transform -- normalizer.py:70-94.  target_means/target_stds: device double[3]. */
int sb_reinhard_transform(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W,
                          const double* target_means, const double* target_stds, int mask_background,
                          double luminosity_threshold, int32_t* status, void* stream);

/* Host feeding (replaces the PIL loading of the reference's callers,
 * stainlib_normalization.ipynb:61-74): batched nvJPEG decode of B baseline JPEGs, all H x W, from HOST memory into the
 * DEVICE batch rgb_out uint8 [B,H,W,3] (interleaved RGB), stream-ordered on `stream`.  jpeg[i] / nbytes[i]: the i-th
 * compressed tile.  SB_ERR_ARG for a corrupt / unsupported stream or a tile that is not H x W.  Backend: nvJPEG's GPU-assisted
 * Huffman decoder (library default as the fallback); the environment variable SB_NVJPEG_BACKEND = default | hybrid | gpu_hybrid |
 * hardware overrides it, SB_ERR_UNSUPPORTED when the requested backend does not exist on the device. */
int sb_decode_jpeg(sb_handle* h, const uint8_t* const* jpeg, const size_t* nbytes, int B, int H, int W, uint8_t* rgb_out,
                   void* stream);

/* LuminosityStandardizer.standardize -- stain_utils.py:53-67. */
int sb_luminosity_standardize(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W,
                              double percentile, void* stream);

/* HedColorAugmenter.transform -- augmenter.py:276-331 for given draws.  sigma/bias: device double [B,3] (H,E,D);
 * tiles whose mean/255 lies outside [cutoff_lo, cutoff_hi] are copied through (status bit 0 set to 1 = skipped).
 * skimage_variant selects the definition of skimage.color.rgb2hed / hed2rgb (augmenter.py:295,319):
 *   17  scikit-image 0.16-0.17 (pinned by the reference's environment.yml:107): -log_b(rgb + 2) @ inv(M), log_base 10
 *       (pass log_base e for scikit-image <= 0.15);
 *   18  scikit-image >= 0.18 (what the unpinned setup.py:11-18 installs today): max(0, ln(max(rgb,1e-6))/ln(1e-6) @ inv(M))
 *       and exp(ln(1e-6) hed @ M) clipped to [0,1]; log_base is ignored. */
int sb_hed_augment(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W, const double* sigma,
                   const double* bias, double cutoff_lo, double cutoff_hi, double log_base, int skimage_variant,
                   int32_t* status, void* stream);
/* The same for float patches (augmenter.py:288-291, 323-327: a float image in [0,1] is taken as it is and a float image
 * is returned): rgb_in / rgb_out device float [B,H,W,3]. */
int sb_hed_augment_f32(sb_handle* h, const float* rgb_in, float* rgb_out, int B, int H, int W, const double* sigma,
                       const double* bias, double cutoff_lo, double cutoff_hi, double log_base, int skimage_variant,
                       int32_t* status, void* stream);

/* GrayscaleAugmentor.pop -- augmenter.py:390-401 for given draws alpha/beta: device double [B]. */
int sb_grayscale_augment(sb_handle* h, const uint8_t* rgb_in, uint8_t* rgb_out, int B, int H, int W,
                         const double* alpha, const double* beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STAINB200_H */
