// sb_io.cu -- host feeding: batched JPEG decode on the GPU (nvJPEG) straight into the
// uint8 [B,H,W,3] device batch the stain kernels consume, so that compressed tiles cross PCIe instead of raw pixels
// (a 512x512 H&E tile is ~60-100 KB as JPEG against 768 KB raw: the 47 GB/s duplex link that bounds sb_normalize_host
// at ~15 Gpx/s stops being the bottleneck on the way in).  The reference's callers load tiles with PIL
// (stainlib_normalization.ipynb:61-74); this is the B200-native replacement of that step, not of any stainlib function.
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include <nvjpeg.h>

#include "sb_kernels.h"

namespace {

struct JpegState {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int batch = 0;
};
std::mutex g_mu;
constexpr int N_BACKEND = 4;   // 0 library default (nvjpegCreateSimple), 1 hybrid (CPU Huffman), 2 gpu_hybrid (GPU Huffman), 3 hardware (NVJPG engines)
JpegState g_jpeg[64][N_BACKEND];          // one per device and backend

// Batched decode defaults to the GPU-assisted Huffman backend: measured on B200 (tools/jpeg_probe.py, 256 tiles of 512^2, q90)
// 16 600 tiles/s = 4.35 Gpx/s against 1 530 tiles/s = 0.40 Gpx/s for the library default (Huffman decode on one CPU thread); the
// NVJPG hardware backend is not offered on this part.  SB_NVJPEG_BACKEND = default | hybrid | gpu_hybrid | hardware overrides
// (read at every call); a backend the device does not offer fails with SB_ERR_UNSUPPORTED when asked for explicitly, while the
// implicit choice falls back to the library default.
int backend_from_env(bool& explicit_choice) {
    const char* e = getenv("SB_NVJPEG_BACKEND");
    explicit_choice = e != nullptr;
    if (!e) return 2;
    const std::string v(e);
    if (v == "hybrid") return 1;
    if (v == "gpu_hybrid") return 2;
    if (v == "hardware") return 3;
    return 0;
}

int jpeg_fail(nvjpegStatus_t s) { return s == NVJPEG_STATUS_SUCCESS ? SB_OK : (s == NVJPEG_STATUS_INVALID_PARAMETER || s == NVJPEG_STATUS_BAD_JPEG ||
                                                                                  s == NVJPEG_STATUS_JPEG_NOT_SUPPORTED ? SB_ERR_ARG : SB_ERR_CUDA); }

}  // namespace

extern "C" {

int sb_decode_jpeg(sb_handle* h, const uint8_t* const* jpeg, const size_t* nbytes, int B, int H, int W, uint8_t* rgb_out,
                   void* stream) {
    if (!h || !jpeg || !nbytes || !rgb_out || B <= 0 || H <= 0 || W <= 0 || h->device < 0 || h->device >= 64) return SB_ERR_ARG;
    sb::DeviceGuard guard(h);
    if (!guard.ok) return SB_ERR_CUDA;
    sb::NvtxRange nvtx("sb_decode_jpeg (nvJPEG)");
    std::lock_guard<std::mutex> lock(g_mu);
    bool explicit_choice = false;
    int be = backend_from_env(explicit_choice);
    static const nvjpegBackend_t kBackend[N_BACKEND] = {NVJPEG_BACKEND_DEFAULT, NVJPEG_BACKEND_HYBRID, NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_HARDWARE};
    for (;;) {
        JpegState& c = g_jpeg[h->device][be];
        if (c.handle) break;
        const nvjpegStatus_t cs = be == 0 ? nvjpegCreateSimple(&c.handle) : nvjpegCreateEx(kBackend[be], nullptr, nullptr, 0, &c.handle);
        if (cs == NVJPEG_STATUS_SUCCESS) {
            if (nvjpegJpegStateCreate(c.handle, &c.state) != NVJPEG_STATUS_SUCCESS) { nvjpegDestroy(c.handle); c.handle = nullptr; return SB_ERR_CUDA; }
            break;
        }
        c.handle = nullptr;
        if (be == 0) return SB_ERR_CUDA;
        if (explicit_choice) return SB_ERR_UNSUPPORTED;
        be = 0;                                          // implicit choice not available: library default
    }
    JpegState& js = g_jpeg[h->device][be];
    // every tile must be H x W (the batch is one dense tensor)
    for (int i = 0; i < B; ++i) {
        int comps = 0, widths[NVJPEG_MAX_COMPONENT] = {0}, heights[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        nvjpegStatus_t s = nvjpegGetImageInfo(js.handle, jpeg[i], nbytes[i], &comps, &sub, widths, heights);
        if (s != NVJPEG_STATUS_SUCCESS) return jpeg_fail(s);
        if (widths[0] != W || heights[0] != H) return SB_ERR_ARG;
    }
    if (js.batch != B) {
        nvjpegStatus_t s = nvjpegDecodeBatchedInitialize(js.handle, js.state, B, 1, NVJPEG_OUTPUT_RGBI);
        if (s != NVJPEG_STATUS_SUCCESS) return jpeg_fail(s);
        js.batch = B;
    }
    std::vector<nvjpegImage_t> dst((size_t)B);
    for (int i = 0; i < B; ++i) {
        for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { dst[i].channel[c] = nullptr; dst[i].pitch[c] = 0; }
        dst[i].channel[0] = rgb_out + (size_t)i * H * W * 3;
        dst[i].pitch[0] = (size_t)W * 3;
    }
    nvjpegStatus_t s = nvjpegDecodeBatched(js.handle, js.state, jpeg, nbytes, dst.data(), (cudaStream_t)stream);
    if (s != NVJPEG_STATUS_SUCCESS) return jpeg_fail(s);
    h->launches += 1;
    return SB_OK;
}

}  // extern "C"
