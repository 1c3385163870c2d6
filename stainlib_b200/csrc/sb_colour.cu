// sb_colour.cu -- Reinhard / luminosity standardiser / HED / grayscale entry points (placeholder until implemented).
#include "sb_kernels.h"
extern "C" {
int sb_reinhard_stats(sb_handle*, const uint8_t*, int, int, int, double*, double*, void*) { return SB_ERR_UNSUPPORTED; }
int sb_reinhard_transform(sb_handle*, const uint8_t*, uint8_t*, int, int, int, const double*, const double*, int, double, int32_t*, void*) { return SB_ERR_UNSUPPORTED; }
int sb_luminosity_standardize(sb_handle*, const uint8_t*, uint8_t*, int, int, int, double, void*) { return SB_ERR_UNSUPPORTED; }
int sb_hed_augment(sb_handle*, const uint8_t*, uint8_t*, int, int, int, const double*, const double*, double, double, double, int32_t*, void*) { return SB_ERR_UNSUPPORTED; }
int sb_grayscale_augment(sb_handle*, const uint8_t*, uint8_t*, int, int, int, const double*, const double*, void*) { return SB_ERR_UNSUPPORTED; }
}
