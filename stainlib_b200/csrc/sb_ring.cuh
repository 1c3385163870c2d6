// sb_ring.cuh -- the TMA-staged shared-memory ring of K4 (sb_recombine.cu) as a reusable building block for the other
// single-pass per-pixel operators (HED, stain augmentation, grayscale): same transport, the arithmetic is an "Op".
//
// Persistent CTAs (one per SM).  The batch is cut into chunks of GT*48 bytes; a producer warp streams chunks
// HBM -> shared memory with cp.async.bulk (TMA, 1-D) completing on an mbarrier per ring slot, GT compute threads
// transform their own 48-byte group IN PLACE in shared memory and hand the slot back through a second mbarrier, and
// the chunk leaves with one bulk store.  No thread ever waits on a global load; both directions are fully coalesced
// by the copy engine.  The CTA takes the whole 227 KB and places a 64 KB lane-replicated lookup table at the 64 KB
// aligned shared address inside it, so that one PRMT on the packed pixel word yields the LDS address of a lookup.
//
// An Op provides:
//   Consts                  per-tile constants, read from Params::consts[tile] once per run of chunks of that tile
//   Params                  by-value kernel argument (pointers to the constants and to whatever the Op accumulates)
//   Acc                     per-thread accumulator carried over a run (e.g. a byte sum)
//   kLaneShift              log2 of the table entry size per lane (2: float, 3: float2)
//   fill_table(tab, p, tid, n)   all n threads fill the 64 KB table
//   Run / begin_run(k, p)        per-run state derived once from the tile's constants (e.g. which code variant to run)
//   process(k, p, run, tab, grp, acc)  transform the 48-byte group at grp (3 x uint4 in shared memory) in place
//   finish_run(p, tile, acc)     called by every compute thread when the CTA leaves a tile (warp collectives allowed)
#pragma once
#include "sb_kernels.h"

namespace sb {

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
// The producer lane is AHEAD of the compute warps for the whole kernel, i.e. it is always waiting for a stage to be handed back:
// a bare try_wait loop keeps issuing (try_wait / yield / branch every few cycles) on the scheduler it shares with four compute
// warps, and the stages are released at the pace of the slowest warp.  The producer therefore sleeps between polls: a stage
// freed 0.1 us late is harmless with several chunks in flight.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(100);
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

constexpr int RING_SMEM_BYTES = 227 * 1024;
constexpr int RING_BAR_BYTES = 256;

struct RingGeom {
    const uint8_t* in;
    uint8_t* out;
    int B, npx;          // every tile is npx*3 bytes, a whole number of 16-byte vectors at a 16-byte aligned address
};

template <class Op, int GT, int NSTAGE>
__global__ void __launch_bounds__(GT + 32, 1) ring_pointwise_kernel(RingGeom g, typename Op::Params p, int chunks_per_tile, long long total_chunks) {
    constexpr int TT_ALL = GT + 32;
    constexpr int CHUNK_BYTES = GT * 48;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t tab_addr = (base + 0xFFFFu) & ~0xFFFFu;
    unsigned char* tab_ptr = smem + (tab_addr - base);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);                            // TMA load landed
    uint64_t* done = full + NSTAGE;                                                // all compute warps wrote the stage back
    const int n_front = tab_addr - base >= (uint32_t)RING_BAR_BYTES ? (int)((tab_addr - base - RING_BAR_BYTES) / CHUNK_BYTES) : 0;
    const int n_back = ((int)RING_SMEM_BYTES - (int)(tab_addr - base) - OD_REP_BYTES) / CHUNK_BYTES;
    if (n_front + n_back < NSTAGE || tab_addr - base < (uint32_t)RING_BAR_BYTES) __trap();
    auto stage_ptr = [&](int s) -> unsigned char* {
        return s < n_front ? smem + RING_BAR_BYTES + (size_t)s * CHUNK_BYTES : tab_ptr + OD_REP_BYTES + (size_t)(s - n_front) * CHUNK_BYTES;
    };
    const size_t tile_bytes = (size_t)g.npx * 3;
    const long long c_begin = total_chunks * blockIdx.x / gridDim.x, c_end = total_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);

    auto chunk_geom = [&](long long c, int& tile, size_t& off, uint32_t& bytes) {
        tile = (int)(c / chunks_per_tile);
        off = (size_t)(c % chunks_per_tile) * CHUNK_BYTES;
        const size_t rem = tile_bytes - off;
        bytes = (uint32_t)(rem < (size_t)CHUNK_BYTES ? rem : (size_t)CHUNK_BYTES);
    };

    if (threadIdx.x == GT) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], GT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    Op::fill_table(tab_ptr, p, (int)threadIdx.x, TT_ALL);
    __syncthreads();

    if (threadIdx.x >= GT) {
        // ------------------------------------------------------------------ producer warp (one elected lane)
        if (threadIdx.x == GT) {
            for (int i = 0; i < NSTAGE && i < n_local; ++i) {
                int tile; size_t off; uint32_t bytes;
                chunk_geom(c_begin + i, tile, off, bytes);
                mbar_expect_tx(&full[i], bytes);
                bulk_load(stage_ptr(i), g.in + (size_t)tile * tile_bytes + off, bytes, &full[i]);
            }
            for (int i = 0; i < n_local; ++i) {
                const int s = i % NSTAGE;
                int tile; size_t off; uint32_t bytes;
                chunk_geom(c_begin + i, tile, off, bytes);
                mbar_wait_relaxed(&done[s], (uint32_t)((i / NSTAGE) & 1));  // stage s holds the finished output of chunk i
                bulk_store(g.out + (size_t)tile * tile_bytes + off, stage_ptr(s), bytes);
                // refill the stage of chunk i-1 once its store has finished reading shared memory
                if (i >= 1 && i - 1 + NSTAGE < n_local) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    const int ps = (i - 1) % NSTAGE;
                    int t2; size_t o2; uint32_t b2;
                    chunk_geom(c_begin + i - 1 + NSTAGE, t2, o2, b2);
                    mbar_expect_tx(&full[ps], b2);
                    bulk_load(stage_ptr(ps), g.in + (size_t)t2 * tile_bytes + o2, b2, &full[ps]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warps
    const OdAbs tab{((threadIdx.x & 31u) << Op::kLaneShift) | ((tab_addr >> 16) << 8)};
    typename Op::Acc acc;
    Op::acc_init(acc);
    int i = 0;
    while (i < n_local) {
        // run of chunks that belong to one tile: constants are loaded once per run
        const int tile = (int)((c_begin + i) / chunks_per_tile);
        const int first_in_tile = (int)((c_begin + i) - (long long)tile * chunks_per_tile);
        int run = chunks_per_tile - first_in_tile;
        if (run > n_local - i) run = n_local - i;
        const typename Op::Consts k = p.consts[tile];
        const typename Op::Run rs = Op::begin_run(k, p);
        for (int j = 0; j < run; ++j, ++i) {
            const int s = i % NSTAGE;
            const uint32_t parity = (uint32_t)((i / NSTAGE) & 1);
            const size_t off = (size_t)(first_in_tile + j) * CHUNK_BYTES;
            const size_t rem = tile_bytes - off;
            const uint32_t bytes = (uint32_t)(rem < (size_t)CHUNK_BYTES ? rem : (size_t)CHUNK_BYTES);
            unsigned char* buf = stage_ptr(s);
            mbar_wait(&full[s], parity);
            if (threadIdx.x * 48u < bytes) {
                Op::process(k, p, rs, tab, reinterpret_cast<uint4*>(buf + threadIdx.x * 48u), acc);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk store
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
        }
        Op::finish_run(p, tile, acc);
    }
}

// Launches the ring over a batch of 16-byte aligned tiles.
template <class Op, int GT = 512, int NSTAGE = 6>
static int launch_ring(const RingGeom& g, const typename Op::Params& p, int num_sms, cudaStream_t stream) {
    constexpr int CHUNK_BYTES = GT * 48;
    static_assert(GT % 32 == 0 && GT + 32 <= 1024, "block size");
    static_assert(OD_REP_BYTES + NSTAGE * CHUNK_BYTES + 1024 + RING_BAR_BYTES <= RING_SMEM_BYTES, "ring does not fit");
    static_assert(2 * NSTAGE * 8 <= RING_BAR_BYTES, "barrier area");
    static DeviceOnce once;      // one per instantiation (function-local static of a function template)
    {
        cudaError_t e = ensure_dyn_smem(once, ring_pointwise_kernel<Op, GT, NSTAGE>, RING_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
    }
    const size_t tile_bytes = (size_t)g.npx * 3;
    const int cpt = (int)((tile_bytes + CHUNK_BYTES - 1) / CHUNK_BYTES);
    const long long total = (long long)cpt * g.B;
    int grid = num_sms;
    if ((long long)grid > total) grid = (int)total;
    ring_pointwise_kernel<Op, GT, NSTAGE><<<grid, GT + 32, RING_SMEM_BYTES, stream>>>(g, p, cpt, total);
    return (int)cudaGetLastError();
}

}  // namespace sb
