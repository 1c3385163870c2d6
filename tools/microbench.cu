// Design-input microbenchmarks for the stain path on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu
// Not part of the product; results are recorded in profiles/r01_microbench.txt.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

// ---------------------------------------------------------------- 1. streaming copy, 16 B vectors
__global__ void copy_u4(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a = __ldcs(in + i), b = __ldcs(in + i + stride), c = __ldcs(in + i + 2 * stride), d = __ldcs(in + i + 3 * stride);
        __stcs(out + i, a); __stcs(out + i + stride, b); __stcs(out + i + 2 * stride, c); __stcs(out + i + 3 * stride, d);
    }
    for (; i < n; i += stride) __stcs(out + i, __ldcs(in + i));
}

// ---------------------------------------------------------------- 2. smem histogram atomics
template <int BINS, int MODE>
__global__ void hist_smem(const uint32_t* __restrict__ keys, size_t n, unsigned* __restrict__ out) {
    extern __shared__ unsigned h[];
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t k = keys[i % 4096] + (uint32_t)i;
    for (; i < n; i += stride) {
        k = k * 1664525u + 1013904223u; k ^= k >> 15;
        if (MODE == 0) atomicAdd(&h[k % BINS], 1u);                       // random bins
        else if (MODE == 1) atomicAdd(&h[(k & ~31u) % BINS | (threadIdx.x & 31)], 1u);  // conflict-free banks
        else if (MODE == 2) { if ((k & 31u) == 0 && (k >> 20 & 1)) atomicAdd(&h[k % BINS], 1u); } // 1/32 of lanes active
        else if (MODE == 3) atomicAdd(&h[(k >> 8) % 64], 1u);            // heavy collisions (64 bins)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) if (h[i]) atomicAdd(&out[i], h[i]);
}

// ---------------------------------------------------------------- 3. LUT lookups in smem
// MODE 0: plain 256-entry float LUT (bank conflicts), MODE 1: lane-replicated 256x32 LUT (conflict-free)
template <int MODE>
__global__ void lut_smem(const uint32_t* __restrict__ px, size_t nwords, float* __restrict__ out) {
    extern __shared__ float lut[];
    const int LSZ = MODE == 0 ? 256 : 256 * 32;
    for (int i = threadIdx.x; i < LSZ; i += blockDim.x) lut[i] = 1.0f / (1 + (MODE == 0 ? i : i / 32));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t w = px[i % 4096] + (uint32_t)i;
    for (; i < nwords; i += stride) {
        w = w * 1664525u + 1013904223u; w ^= w >> 15;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t v = (w >> (8 * b)) & 255u;
            acc += MODE == 0 ? lut[v] : lut[v * 32 + lane];
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// ---------------------------------------------------------------- 4. pipe rates (per-SM issue) : ex2, f2i, ffma, ffma2
template <int MODE>
__global__ void pipe_rate(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 0.25f, d = a + 0.125f;
    float2 p = make_float2(a, b), q = make_float2(c, d);
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0) { a = exp2f(a) ; b = exp2f(b); c = exp2f(c); d = exp2f(d); }
            else if (MODE == 1) { acc += __float2int_rz(a); acc += __float2int_rz(b); a += 1.f; b += 1.f; }
            else if (MODE == 2) { a = fmaf(a, 1.0001f, b); b = fmaf(b, 0.9999f, c); c = fmaf(c, 1.0001f, d); d = fmaf(d, 0.9999f, a); }
            else if (MODE == 3) {
                asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; fma.rn.f32x2 z, x, y, x; mov.b64 {%0,%1}, z; }"
                             : "+f"(p.x), "+f"(p.y) : "f"(q.x), "f"(q.y));
                asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0,%1}; mov.b64 y, {%2,%3}; fma.rn.f32x2 z, x, y, x; mov.b64 {%0,%1}, z; }"
                             : "+f"(q.x), "+f"(q.y) : "f"(p.x), "f"(p.y));
            }
            else if (MODE == 4) { a = __fadd_rd(a, 8388608.f); b = __fadd_rd(b, 8388608.f); c = __fadd_rd(c, 1.f); d = __fadd_rd(d, 1.f); }
            else if (MODE == 5) { a = __log2f(a + 1.f); b = __log2f(b + 1.f); c = __log2f(c + 1.f); d = __log2f(d + 1.f); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + p.x + p.y + q.x + q.y + acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, SMs %d, smem/block optin %zu, L2 %d MB, clock %d MHz\n", prop.name, prop.multiProcessorCount,
           prop.sharedMemPerBlockOptin, prop.l2CacheSize >> 20, clk_khz / 1000);
    const int SM = prop.multiProcessorCount;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    // 1. copy
    {
        size_t bytes = (size_t)1 << 30; uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        CK(cudaMemset(a, 1, bytes));
        for (int blocks_per_sm : {4, 8, 16}) for (int threads : {256, 512}) {
            float best = 1e9;
            for (int r = 0; r < 6; ++r) {
                CK(cudaEventRecord(e0)); copy_u4<<<SM * blocks_per_sm, threads>>>(a, b, bytes / 16); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); float ms = time_ms(e0, e1); if (r && ms < best) best = ms;
            }
            printf("copy_u4 grid=%dxSM threads=%d: %.1f GB/s (r+w)\n", blocks_per_sm, threads, 2.0 * bytes / best / 1e6);
        }
        {   float best = 1e9;
            for (int r = 0; r < 6; ++r) { CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); float ms = time_ms(e0, e1); if (r && ms < best) best = ms; }
            printf("cudaMemcpy D2D: %.1f GB/s (r+w)\n", 2.0 * bytes / best / 1e6); }
        // PCIe
        void* h; CK(cudaMallocHost(&h, (size_t)256 << 20));
        for (int dir = 0; dir < 2; ++dir) { float best = 1e9;
            for (int r = 0; r < 4; ++r) { CK(cudaEventRecord(e0));
                if (dir == 0) CK(cudaMemcpyAsync(a, h, (size_t)256 << 20, cudaMemcpyHostToDevice)); else CK(cudaMemcpyAsync(h, a, (size_t)256 << 20, cudaMemcpyDeviceToHost));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms = time_ms(e0, e1); if (r && ms < best) best = ms; }
            printf("pinned %s: %.1f GB/s\n", dir == 0 ? "H2D" : "D2H", ((size_t)256 << 20) / best / 1e6); }
        {   // bidirectional
            cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2)); void* h2; CK(cudaMallocHost(&h2, (size_t)256 << 20));
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
            for (int r = 0; r < 4; ++r) { CK(cudaMemcpyAsync(a, h, (size_t)256 << 20, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(h2, b, (size_t)256 << 20, cudaMemcpyDeviceToHost, s2)); }
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            printf("pinned bidirectional: %.1f GB/s each way\n", 4.0 * ((size_t)256 << 20) / time_ms(e0, e1) / 1e6); }
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    // 2. histograms
    {
        size_t n = (size_t)1 << 28; uint32_t* keys; CK(cudaMalloc(&keys, n * 4)); unsigned* out; CK(cudaMalloc(&out, 65536 * 4));
        std::vector<uint32_t> hk(1 << 24); uint32_t s = 12345; for (auto& k : hk) { s = s * 1664525u + 1013904223u; k = s >> 4; }
        for (size_t off = 0; off < n; off += hk.size()) CK(cudaMemcpy(keys + off, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice));
        auto run = [&](const char* name, auto kern, int bins) {
            float best = 1e9; CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bins * 4));
            for (int r = 0; r < 4; ++r) { CK(cudaEventRecord(e0)); kern<<<SM * 4, 512, bins * 4>>>(keys, n, out); CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); float ms = time_ms(e0, e1); if (r && ms < best) best = ms; }
            printf("hist %-28s: %.1f Gkeys/s  (%.2f keys/clk/SM @%d MHz nominal)\n", name, n / best / 1e6, n / best / 1e6 * 1e9 / SM / (clk_khz * 1e3), clk_khz / 1000);
        };
        run("random 256 bins", hist_smem<256, 0>, 256);
        run("random 2048 bins", hist_smem<2048, 0>, 2048);
        run("random 4096 bins", hist_smem<4096, 0>, 4096);
        run("conflict-free banks 4096", hist_smem<4096, 1>, 4096);
        run("1/32 lanes active 4096", hist_smem<4096, 2>, 4096);
        run("64 hot bins", hist_smem<4096, 3>, 4096);
        CK(cudaFree(keys)); CK(cudaFree(out));
    }
    // 3. LUT
    {
        size_t nwords = (size_t)1 << 28; uint32_t* px; CK(cudaMalloc(&px, nwords * 4)); float* out; CK(cudaMalloc(&out, 4));
        std::vector<uint32_t> hk(1 << 24); uint32_t s = 999; for (auto& k : hk) { s = s * 1664525u + 1013904223u; k = s ^ (s >> 13); }
        for (size_t off = 0; off < nwords; off += hk.size()) CK(cudaMemcpy(px + off, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 2; ++mode) { float best = 1e9; size_t sm = mode == 0 ? 1024 : 32768;
            for (int r = 0; r < 4; ++r) { CK(cudaEventRecord(e0));
                if (mode == 0) lut_smem<0><<<SM * 4, 512, sm>>>(px, nwords, out); else lut_smem<1><<<SM * 4, 512, sm>>>(px, nwords, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); float ms = time_ms(e0, e1); if (r && ms < best) best = ms; }
            printf("LUT %s: %.1f Glookups/s (%.2f lookups/clk/SM nominal), input %.1f GB/s\n", mode == 0 ? "plain 256" : "lane-replicated", 4.0 * nwords / best / 1e6,
                   4.0 * nwords / best / 1e6 * 1e9 / SM / (clk_khz * 1e3), 4.0 * nwords / best / 1e6); }
        CK(cudaFree(px)); CK(cudaFree(out));
    }
    // 4. pipe rates
    {
        float* out; CK(cudaMalloc(&out, SM * 8 * 256 * 4)); int iters = 4096;
        const char* names[] = {"MUFU.EX2 (4/iter-unit)", "F2I.RZ (2) + FADD (2)", "FFMA (4)", "FFMA2 f32x2 (2 = 4 flop-lanes)", "FADD.RM magic (4)", "MUFU.LG2+FADD (4)"};
        double ops[] = {4, 2, 4, 2, 4, 4};
        for (int mode = 0; mode < 6; ++mode) { float best = 1e9;
            for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0));
                switch (mode) { case 0: pipe_rate<0><<<SM * 8, 256>>>(out, iters); break; case 1: pipe_rate<1><<<SM * 8, 256>>>(out, iters); break;
                    case 2: pipe_rate<2><<<SM * 8, 256>>>(out, iters); break; case 3: pipe_rate<3><<<SM * 8, 256>>>(out, iters); break;
                    case 4: pipe_rate<4><<<SM * 8, 256>>>(out, iters); break; case 5: pipe_rate<5><<<SM * 8, 256>>>(out, iters); break; }
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError()); float ms = time_ms(e0, e1); if (r && ms < best) best = ms; }
            double total = (double)SM * 8 * 256 * iters * 8 * ops[mode];
            printf("pipe %-32s: %.1f Gop/s = %.1f ops/clk/SM @%d MHz nominal\n", names[mode], total / best / 1e6, total / best / 1e6 * 1e9 / SM / (clk_khz * 1e3), clk_khz / 1000); }
        CK(cudaFree(out));
    }
    return 0;
}
