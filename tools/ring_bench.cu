// ring_bench.cu -- transport ceiling of the TMA-staged shared-memory ring (csrc/sb_ring.cuh, csrc/sb_stream.cu) without
// any arithmetic: how fast can one persistent CTA per SM move 24 KB chunks HBM -> shared memory (-> HBM), as a function
// of the ring depth, the chunk size, who issues the stores and how many CTAs share an SM.  Build and run on a B200:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/ring_bench tools/ring_bench.cu && /tmp/ring_bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// MODE 0: one producer thread, refill lags the store by one chunk (sb_ring.cuh today)
// MODE 1: loader thread + storer thread (two warps), slot handed back through an `empty` barrier as soon as the store has read it
// MODE 2: read only (sb_stream.cu ring_reduce): consumers hand the slot straight back
// WORK: the consumers' in-place pass over their 48 bytes (0: none, they only arrive; 1: xor every word)
template <int GT, int NSTAGE, int MODE, int WORK, int CHUNK_MULT>
__global__ void __launch_bounds__(GT + 64) ring_copy(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long total_chunks, int hole) {
    constexpr int CHUNK = GT * 48 * CHUNK_MULT;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + NSTAGE;
    uint64_t* empty = done + NSTAGE;
    unsigned char* stages = smem + 1024 + hole;
    const long long c_begin = total_chunks * blockIdx.x / gridDim.x, c_end = total_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);
    if (threadIdx.x == GT) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], GT / 32); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= GT) {
        if (MODE == 0 && threadIdx.x == GT) {
            for (int i = 0; i < NSTAGE && i < n_local; ++i) {
                mbar_expect_tx(&full[i], CHUNK);
                bulk_load(stages + (size_t)i * CHUNK, in + (c_begin + i) * CHUNK, CHUNK, &full[i]);
            }
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                mbar_wait(&done[s], (uint32_t)((i / NSTAGE) & 1));
                bulk_store(out + (c_begin + i) * CHUNK, stages + (size_t)s * CHUNK, CHUNK);
                if (i >= 1 && i - 1 + NSTAGE < n_local) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int ps = (i - 1) % NSTAGE;
                    mbar_expect_tx(&full[ps], CHUNK);
                    bulk_load(stages + (size_t)ps * CHUNK, in + (c_begin + i - 1 + NSTAGE) * CHUNK, CHUNK, &full[ps]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        if (MODE == 1 && threadIdx.x == GT) {                 // loader
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                if (i >= NSTAGE) mbar_wait(&empty[s], (uint32_t)(((i / NSTAGE) - 1) & 1));
                mbar_expect_tx(&full[s], CHUNK);
                bulk_load(stages + (size_t)s * CHUNK, in + (c_begin + i) * CHUNK, CHUNK, &full[s]);
            }
        }
        if (MODE == 1 && threadIdx.x == GT + 32) {            // storer
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                mbar_wait(&done[s], (uint32_t)((i / NSTAGE) & 1));
                bulk_store(out + (c_begin + i) * CHUNK, stages + (size_t)s * CHUNK, CHUNK);
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(&empty[s]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        if (MODE == 2 && threadIdx.x == GT) {
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                if (i >= NSTAGE) mbar_wait(&done[s], (uint32_t)(((i / NSTAGE) - 1) & 1));
                mbar_expect_tx(&full[s], CHUNK);
                bulk_load(stages + (size_t)s * CHUNK, in + (c_begin + i) * CHUNK, CHUNK, &full[s]);
            }
        }
        return;
    }
    unsigned acc = 0;
    for (int i = 0; i < n_local; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&full[s], (uint32_t)((i / NSTAGE) & 1));
        if (WORK) {
#pragma unroll
            for (int m = 0; m < CHUNK_MULT; ++m) {
                uint4* v = reinterpret_cast<uint4*>(stages + (size_t)s * CHUNK + (size_t)m * GT * 48 + threadIdx.x * 48u);
                uint4 a = v[0], b = v[1], c = v[2];
                if (MODE == 2) acc += a.x ^ b.y ^ c.z;
                else {
                    a.x ^= 0x01010101u; b.y ^= 0x01010101u; c.z ^= 0x01010101u;
                    v[0] = a; v[1] = b; v[2] = c;
                }
            }
            if (MODE != 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
    }
    if (MODE == 2 && acc == 0x12345678u) out[threadIdx.x] = 1;
}

__global__ void copy_u4(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
__global__ void read_u4(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const uint4 v = in[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) out[0] = make_uint4(1, 1, 1, 1);
}

static float time_ms(cudaStream_t st, int reps, void (*launch)(void*), void* ctx) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch(ctx);
    CK(cudaStreamSynchronize(st));
    float best = 1e9f, sum = 0.f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a, st));
        launch(ctx);
        CK(cudaEventRecord(b, st));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best; sum += ms;
    }
    (void)sum;
    return best;
}

struct Ctx { const uint8_t* in; uint8_t* out; size_t bytes; int sms; int ctas_per_sm; int hole; };

template <int GT, int NSTAGE, int MODE, int WORK, int CM>
static void run_ring(const char* name, Ctx c) {
    constexpr int CHUNK = GT * 48 * CM;
    const int smem = 1024 + c.hole + NSTAGE * CHUNK;
    if (smem * c.ctas_per_sm > 227 * 1024 + (c.ctas_per_sm - 1) * 1024) { printf("%-72s does not fit\n", name); return; }
    CK(cudaFuncSetAttribute(ring_copy<GT, NSTAGE, MODE, WORK, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long total = (long long)(c.bytes / CHUNK);
    struct L { Ctx c; long long total; int smem; } l{c, total, smem};
    auto fn = [](void* p) {
        L* l = (L*)p;
        ring_copy<GT, NSTAGE, MODE, WORK, CM><<<l->c.sms * l->c.ctas_per_sm, GT + 64, l->smem>>>(l->c.in, l->c.out, l->total, l->c.hole);
    };
    const float ms = time_ms(0, 8, fn, &l);
    CK(cudaGetLastError());
    const double moved = (double)total * CHUNK * (MODE == 2 ? 1.0 : 2.0);
    printf("%-72s %.4f ms  %7.1f GB/s\n", name, ms, moved / ms * 1e-6);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t bytes = (size_t)1024 * 512 * 512 * 3;
    uint8_t *in, *out;
    CK(cudaMalloc(&in, bytes)); CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(in, 7, bytes)); CK(cudaMemset(out, 0, bytes));
    printf("device %s, %d SMs; %zu MB in, same out; best of 8 launches\n", prop.name, sms, bytes >> 20);
    {
        struct L { const uint8_t* in; uint8_t* out; size_t n; int grid; } l{in, out, bytes / 16, sms * 16};
        float ms = time_ms(0, 8, [](void* p) { L* l = (L*)p; copy_u4<<<l->grid, 512>>>((const uint4*)l->in, (uint4*)l->out, l->n); }, &l);
        printf("%-72s %.4f ms  %7.1f GB/s\n", "copy_u4 16 CTAs/SM x 512 thr (r+w)", ms, 2.0 * bytes / ms * 1e-6);
        ms = time_ms(0, 8, [](void* p) { L* l = (L*)p; read_u4<<<l->grid, 512>>>((const uint4*)l->in, (uint4*)l->out, l->n); }, &l);
        printf("%-72s %.4f ms  %7.1f GB/s\n", "read_u4 16 CTAs/SM x 512 thr (r)", ms, 1.0 * bytes / ms * 1e-6);
        ms = time_ms(0, 8, [](void* p) { L* l = (L*)p; cudaMemcpyAsync(l->out, l->in, l->n * 16, cudaMemcpyDeviceToDevice, 0); }, &l);
        printf("%-72s %.4f ms  %7.1f GB/s\n", "cudaMemcpy D2D (r+w)", ms, 2.0 * bytes / ms * 1e-6);
    }
    Ctx c{in, out, bytes, sms, 1, 64 * 1024};
    Ctx c0{in, out, bytes, sms, 1, 0};
    Ctx c2{in, out, bytes, sms, 2, 0};
    printf("-- read + write, 64 KB table hole (K4 layout)\n");
    run_ring<512, 6, 0, 0, 1>("rw  mode0 (today)  6 x 24 KB, consumers idle", c);
    run_ring<512, 6, 0, 1, 1>("rw  mode0 (today)  6 x 24 KB, consumers xor", c);
    run_ring<512, 6, 1, 0, 1>("rw  mode1 (ld+st threads)  6 x 24 KB, idle", c);
    run_ring<512, 6, 1, 1, 1>("rw  mode1 (ld+st threads)  6 x 24 KB, xor", c);
    run_ring<256, 12, 0, 1, 1>("rw  mode0  12 x 12 KB (256 thr), xor", c);
    run_ring<256, 12, 1, 1, 1>("rw  mode1  12 x 12 KB (256 thr), xor", c);
    run_ring<512, 3, 1, 1, 2>("rw  mode1  3 x 48 KB, xor", c);
    run_ring<1024 - 64, 3, 1, 1, 1>("rw  mode1  3 x 45 KB (960 thr), xor", c);
    printf("-- read + write, no hole\n");
    run_ring<512, 9, 0, 1, 1>("rw  mode0  9 x 24 KB, xor", c0);
    run_ring<512, 9, 1, 1, 1>("rw  mode1  9 x 24 KB, xor", c0);
    run_ring<512, 4, 1, 1, 1>("rw  mode1  2 CTAs/SM x 4 x 24 KB, xor", c2);
    run_ring<256, 9, 1, 1, 1>("rw  mode1  2 CTAs/SM x 9 x 12 KB (256 thr), xor", c2);
    printf("-- read only, 64 KB table hole (ring_reduce layout)\n");
    run_ring<512, 6, 2, 0, 1>("r   6 x 24 KB, consumers idle", c);
    run_ring<512, 6, 2, 1, 1>("r   6 x 24 KB, consumers read", c);
    run_ring<256, 12, 2, 1, 1>("r   12 x 12 KB (256 thr), read", c);
    run_ring<512, 3, 2, 1, 2>("r   3 x 48 KB, read", c);
    printf("-- read only, no hole\n");
    run_ring<512, 9, 2, 1, 1>("r   9 x 24 KB, read", c0);
    run_ring<512, 4, 2, 1, 1>("r   2 CTAs/SM x 4 x 24 KB, read", c2);
    return 0;
}
