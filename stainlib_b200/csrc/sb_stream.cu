// sb_stream.cu -- the Macenko statistics of ExtractiveStainNormalizer.fit / .transform (normalizer.py:27-50,
// macenko_stain_extractor.py:7-44) as STREAMING passes over the whole batch, one launch per pass.
//
// The fused per-tile kernel (sb_pipeline.cu) keeps a tile inside one CTA for its whole dependency chain; the price is
// that 296 resident CTAs each re-read "their" tile three times with per-thread global loads (233 MB in flight: more
// than the L2 holds), stall on those loads, idle 511 threads during the serial fp64 steps, and quantise the batch into
// whole tiles per CTA.  Here every full pass is a persistent, read-only kernel on the TMA ring of sb_ring.cuh:
//
//   1  ring_reduce<MomentOp>     tissue mask + masked OD moments of every tile            (3 B/px from HBM, once)
//   2  plan_angle_kernel         per tile: covariance, fp64 eigenvectors, 1-in-16 sample of the angle keys -> brackets
//   3  ring_reduce<AngleOp>      count the keys below each bracket, list the keys inside  (3 B/px)
//   4  select_angle_kernel       per tile: exact angular percentiles -> stain matrix; sample of the concentrations -> brackets
//   5  ring_reduce<ConcOp>       the same for the two concentrations of ALL pixels        (3 B/px)
//   6  select_conc_kernel        per tile: exact 99th percentiles -> maxC
//
// The ring kernels cut the batch into 24 KB chunks and give every CTA (one per SM) a contiguous run of chunks: work is
// balanced to the chunk, not to the tile; a producer warp streams chunks HBM -> shared memory with cp.async.bulk while
// 16 compute warps work on earlier chunks, so no thread ever waits on a global load; each byte crosses the HBM bus
// exactly once per pass.  Per-tile results are sums -- fixed-point int64 moments, integer counts, key lists -- that CTAs
// add into a per-tile state record with atomics when they leave a tile (integer addition: any order gives the same
// bits).  The per-tile serial steps run in their own small kernels, one CTA per tile, all tiles at once.
//
// Per-pixel arithmetic, keys and selections are the SAME code as in the fused kernel (sb_pipe_common.cuh), so a tile
// gives the same bits on either path.  Tiles the streaming path does not serve -- fewer than 16,384 tissue pixels, a
// sample too small for brackets, a bracket that missed its rank, a key list that overflowed, non-unit stain vectors --
// are appended to a device-side list and handled by the fused kernel behind the passes (no host synchronisation).
#include "sb_kernels.h"
#include "sb_pipe_common.cuh"
#include "sb_ring.cuh"

namespace sb {

constexpr int SLIST_CAP_MAX = 16384;                // entries of a per-tile key list in global memory (full 23-bit keys): 8192 for tiles up
// to a megapixel, 16384 beyond; 4096 for tiles up to 2^17 pixels (their brackets hold < 2000 keys), which lets four select CTAs share an SM
__host__ __device__ inline int slist_cap(int npx) { return npx > (1 << 20) ? SLIST_CAP_MAX : npx > (1 << 17) ? SLIST_CAP_MAX / 2 : SLIST_CAP_MAX / 4; }
constexpr int RQ_CAP = 64;                          // entries per warp queue: drained below 32 after every push round
constexpr int RR_GT = 512;                          // compute threads of a ring-reduce CTA (one 16-pixel group each per chunk)
constexpr int RR_STAGES = 6;
constexpr int RR_CHUNK = RR_GT * 48;
constexpr int RR_WARP_WORDS = 3 * RQ_CAP;           // per warp: rare-pixel queue + two key-list staging buffers, 64 words each
constexpr int RR_HEAD_BYTES = 512 + (RR_GT / 32) * RR_WARP_WORDS * 4;     // barriers + per-warp scratch, in front of the first stage

enum TilePath { PATH_STREAM = 0, PATH_FALLBACK = 1, PATH_FLAGGED = 2 };

struct __align__(16) TileState {
    unsigned long long mom[10];   // pass 1: fixed-point (FIX_MOMENT) sums of od (3) and od x od (6) over the tissue pixels; their count
    unsigned below[2], len[2];    // pass 3: keys below each angle bracket / listed inside it
    unsigned cbelow[2], clen[2];  // pass 5: the same for the concentration brackets
    double Vd[6];                 // the two leading eigenvectors (rows), fp64
    double Msrc[6];
    unsigned brk[4];              // angle brackets  [ka0, kb0), [ka1, kb1)
    unsigned cbrk[4];             // concentration brackets
    unsigned n_tissue;
    int flags;                    // SB_STATUS_* bits
    int path;                     // TilePath
    int pad;
};

struct __align__(16) AngleConsts {
    float v[6];                   // the projection plane (exact keys)
    float n1[3], n2[3];           // the wedge between the brackets as two half-spaces of OD space (fast test)
    float margin;                 // n . od must exceed it (+inf: no usable wedge, every tissue pixel takes the exact path)
    unsigned ka0, kb0, ka1, kb1;
    int mode;                     // 0 = process, 1 = skip this tile
    int pad[2];
};
struct __align__(16) ConcConsts {
    LassoK lk;                    // 13 floats
    float mid0, half0, mid1, half1;   // float window around each key bracket (64 key units of slack): |m - mid| <= half -> exact path
    unsigned ka0, kb0, ka1, kb1;
    int lm;                       // LASSO_UNIT_POS / LASSO_UNIT_NEG
    int mode;
    int pad;
};

// Vahadane (vahadane_stain_extractor.py:19-43; the dictionary iteration of the fused kernel, one launch per pass)
struct __align__(16) DictState {
    double D[6];                  // dictionary, rows = atoms
    unsigned long long sums[10];  // fixed-point (FIX_DL) sums of the running pass: a a^T (3), x a_0 (3), x a_1 (3)
    AAState aa;                   // Anderson history of the fixed-point iteration
    int it, n_it;                 // full passes done / allowed
    int pad[2];
};
struct __align__(16) DictConsts {
    LassoK lk;                    // 13 floats
    int lm;                       // LASSO_GENERAL / LASSO_UNIT_POS / LASSO_UNIT_NEG
    int mode;                     // 0 = the tile takes part in the next pass, 1 = skip (converged, flagged, fallback)
    int pad;
};

// ------------------------------------------------------------------------------------------------ lookups
// {od, gamma} pair table at an absolute 64 KB-aligned shared address (one PRMT = the LDS.64 address).
template <class F>
__device__ __forceinline__ void for_each_px_odg_abs(const OdAbs& t, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, odg_lookup_abs(t, a, 0), odg_lookup_abs(t, a, 1), odg_lookup_abs(t, a, 2));
        f(4 * q + 1, odg_lookup_abs(t, a, 3), odg_lookup_abs(t, b, 0), odg_lookup_abs(t, b, 1));
        f(4 * q + 2, odg_lookup_abs(t, b, 2), odg_lookup_abs(t, b, 3), odg_lookup_abs(t, c, 0));
        f(4 * q + 3, odg_lookup_abs(t, c, 1), odg_lookup_abs(t, c, 2), odg_lookup_abs(t, c, 3));
    }
}
// od-only table, two pixels at a time (same pairing as for_each_pair_od of the fused kernel).
template <class F>
__device__ __forceinline__ void for_each_pair_od_abs(const OdAbs& t, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, f2(od_lookup(t, a, 0u, 0), od_lookup(t, a, 0u, 3)), f2(od_lookup(t, a, 0u, 1), od_lookup(t, b, 0u, 0)),
          f2(od_lookup(t, a, 0u, 2), od_lookup(t, b, 0u, 1)));
        f(4 * q + 2, f2(od_lookup(t, b, 0u, 2), od_lookup(t, c, 0u, 1)), f2(od_lookup(t, b, 0u, 3), od_lookup(t, c, 0u, 2)),
          f2(od_lookup(t, c, 0u, 0), od_lookup(t, c, 0u, 3)));
    }
}
// od-only table, one pixel at a time (general LASSO form of the Vahadane passes).
template <class F>
__device__ __forceinline__ void for_each_px_od_abs(const OdAbs& t, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, od_lookup(t, a, 0u, 0), od_lookup(t, a, 0u, 1), od_lookup(t, a, 0u, 2));
        f(4 * q + 1, od_lookup(t, a, 0u, 3), od_lookup(t, b, 0u, 0), od_lookup(t, b, 0u, 1));
        f(4 * q + 2, od_lookup(t, b, 0u, 2), od_lookup(t, b, 0u, 3), od_lookup(t, c, 0u, 0));
        f(4 * q + 3, od_lookup(t, c, 0u, 1), od_lookup(t, c, 0u, 2), od_lookup(t, c, 0u, 3));
    }
}
// Plain 256-entry {od, gamma} table in shared memory for the per-tile kernels (1/16 of the pixels: bank conflicts are
// cheaper than a 64 KB lane-replicated table per CTA).
template <class F>
__device__ __forceinline__ void for_each_px_odg_small(const float2* tab, const uint32_t (&w)[12], F&& f) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        f(4 * q + 0, tab[a & 255u], tab[(a >> 8) & 255u], tab[(a >> 16) & 255u]);
        f(4 * q + 1, tab[a >> 24], tab[b & 255u], tab[(b >> 8) & 255u]);
        f(4 * q + 2, tab[(b >> 16) & 255u], tab[b >> 24], tab[c & 255u]);
        f(4 * q + 3, tab[(c >> 8) & 255u], tab[(c >> 16) & 255u], tab[c >> 24]);
    }
}

// bits |= bit when x > thr (x < thr): one compare into a predicate and one predicated OR -- set.gt + and + or costs a third
// instruction on the ALU pipe, which bounds these kernels.
__device__ __forceinline__ void or_if_gt(unsigned& bits, float x, float thr, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bits) : "f"(x), "f"(thr), "r"(bit));
}
__device__ __forceinline__ void or_if_lt(unsigned& bits, float x, float thr, unsigned bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bits) : "f"(x), "f"(thr), "r"(bit));
}

struct WarpScratch {
    unsigned* w;                  // [0, 64) rare-pixel queue, [64, 128) staged keys of list 0, [128, 192) of list 1
    unsigned qlen, n0, n1;        // warp-uniform fill counts
};

// ------------------------------------------------------------------------------------------------ read-only ring
// The transport of sb_ring.cuh without the store: chunks stream HBM -> shared memory, compute warps reduce them and
// hand the slot straight back.  An Op provides
//   Consts / load_consts(p, tile)      per-tile constants, loaded once per run of chunks of one tile
//   Acc / acc_init                      per-thread accumulators carried over a run
//   kLaneShift, fill_table              the 64 KB lane-replicated lookup table
//   process(k, p, tab, buf, active, px0, wq, acc, tile)   called by ALL lanes of a compute warp for the thread's group
//                                       at buf (active = the group exists); px0 = tile pixel index of the chunk's start
//   finish_run(k, p, tab, wq, acc, tile)   all compute threads, when the CTA leaves a tile
// Per-warp scratch (WarpScratch): a queue of rare pixels (their packed RGB, so entries outlive the chunk they came from
// and are processed 32 at a time with all lanes busy) and two staging buffers for the keys that go to the tile's global
// key lists (flushed 32 at a time: one atomic and one coalesced 128-byte store per 32 keys).
template <class Op>
__global__ void __launch_bounds__(RR_GT + 32, 1) ring_reduce_kernel(RingGeom g, typename Op::Params p, int chunks_per_tile, int unit_chunks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t tab_addr = (base + RR_HEAD_BYTES + 0xFFFFu) & ~0xFFFFu;
    unsigned char* tab_ptr = smem + (tab_addr - base);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + RR_STAGES;
    unsigned* queues = reinterpret_cast<unsigned*>(smem + 512);
    const int n_front = (int)((tab_addr - base - RR_HEAD_BYTES) / RR_CHUNK);
    const int n_back = ((int)RING_SMEM_BYTES - (int)(tab_addr - base) - OD_REP_BYTES) / RR_CHUNK;
    if (n_front + n_back < RR_STAGES) __trap();
    auto stage_ptr = [&](int s) -> unsigned char* {
        return s < n_front ? smem + RR_HEAD_BYTES + (size_t)s * RR_CHUNK : tab_ptr + OD_REP_BYTES + (size_t)(s - n_front) * RR_CHUNK;
    };
    const size_t tile_bytes = (size_t)g.npx * 3;
    // The CTA's share: a contiguous range of UNITS (unit_chunks consecutive chunks of one tile; 1 for the Macenko passes, the
    // unit of for_each_unit for the Vahadane passes, whose fp32 partial sums are defined per unit), as a range of chunks.
    const int upt = (chunks_per_tile + unit_chunks - 1) / unit_chunks;
    const long long total_units = (long long)upt * g.B;
    auto unit_chunk = [&](long long u) -> long long { return u >= total_units ? (long long)chunks_per_tile * g.B : (u / upt) * chunks_per_tile + (u % upt) * unit_chunks; };
    const long long c_begin = unit_chunk(total_units * blockIdx.x / gridDim.x), c_end = unit_chunk(total_units * (blockIdx.x + 1) / gridDim.x);
    const int n_local = (int)(c_end - c_begin);
    {
        // nothing to do in this CTA's range (e.g. a Vahadane pass after every tile has converged): leave before the table fill
        int any = 0;
        if (n_local > 0) {
            const int t_first = (int)(c_begin / chunks_per_tile), t_last = (int)((c_end - 1) / chunks_per_tile);
            for (int t = t_first + (int)threadIdx.x; t <= t_last; t += RR_GT + 32) any |= Op::tile_active(p, t) ? 1 : 0;
        }
        if (!__syncthreads_or(any)) return;
    }
    if (threadIdx.x == RR_GT) {
        for (int s = 0; s < RR_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], RR_GT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    Op::fill_table(tab_ptr, p, (int)threadIdx.x, RR_GT + 32);
    __syncthreads();

    // Producer and consumers walk the same list of runs (chunks of one tile inside the CTA's range) and both skip the tiles
    // the Op declares inactive (flagged / fallback / converged tiles: their bytes never leave HBM); n counts loaded chunks.
    if (threadIdx.x >= RR_GT) {
        // ------------------------------------------------------------------ producer warp (one elected lane)
        if (threadIdx.x == RR_GT) {
            int n = 0;
            for (int i = 0; i < n_local;) {
                const int tile = (int)((c_begin + i) / chunks_per_tile);
                const int first_in_tile = (int)((c_begin + i) - (long long)tile * chunks_per_tile);
                int run = chunks_per_tile - first_in_tile;
                if (run > n_local - i) run = n_local - i;
                i += run;
                if (!Op::tile_active(p, tile)) continue;
                for (int j = 0; j < run; ++j, ++n) {
                    const int s = n % RR_STAGES;
                    if (n >= RR_STAGES) mbar_wait_relaxed(&done[s], (uint32_t)(((n / RR_STAGES) - 1) & 1));     // the slot's previous chunk has been consumed
                    const size_t off = (size_t)(first_in_tile + j) * RR_CHUNK;
                    const size_t rem = tile_bytes - off;
                    const uint32_t bytes = (uint32_t)(rem < (size_t)RR_CHUNK ? rem : (size_t)RR_CHUNK);
                    mbar_expect_tx(&full[s], bytes);
                    bulk_load(stage_ptr(s), g.in + (size_t)tile * tile_bytes + off, bytes, &full[s]);
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warps
    const OdAbs tab{((threadIdx.x & 31u) << Op::kLaneShift) | ((tab_addr >> 16) << 8)};
    WarpScratch wq{queues + (threadIdx.x >> 5) * RR_WARP_WORDS, 0u, 0u, 0u};
    typename Op::Acc acc;
    Op::acc_init(acc);
    int n = 0;
    for (int i = 0; i < n_local;) {
        const int tile = (int)((c_begin + i) / chunks_per_tile);
        const int first_in_tile = (int)((c_begin + i) - (long long)tile * chunks_per_tile);
        int run = chunks_per_tile - first_in_tile;
        if (run > n_local - i) run = n_local - i;
        i += run;
        if (!Op::tile_active(p, tile)) continue;
        const typename Op::Consts k = Op::load_consts(p, tile);
        int in_unit = first_in_tile % unit_chunks;      // (CTA ranges start at unit boundaries: 0 except for unit_chunks == 1)
        for (int j = 0; j < run; ++j, ++n) {
            const int s = n % RR_STAGES;
            const int ci = first_in_tile + j;
            const size_t off = (size_t)ci * RR_CHUNK;
            const size_t rem = tile_bytes - off;
            const uint32_t bytes = (uint32_t)(rem < (size_t)RR_CHUNK ? rem : (size_t)RR_CHUNK);
            unsigned char* buf = stage_ptr(s);
            mbar_wait(&full[s], (uint32_t)((n / RR_STAGES) & 1));
            if (++in_unit == unit_chunks || ci + 1 == chunks_per_tile) in_unit = 0;
            const bool unit_end = in_unit == 0;
            Op::process(k, p, tab, buf, threadIdx.x * 48u < bytes, (unsigned)(off / 3), wq, acc, tile, unit_end);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
        }
        Op::finish_run(k, p, tab, wq, acc, tile);
    }
}

template <class Op>
static int launch_ring_reduce(const RingGeom& g, const typename Op::Params& p, int num_sms, cudaStream_t stream, int unit_chunks = 1) {
    static_assert(OD_REP_BYTES + RR_STAGES * RR_CHUNK + RR_HEAD_BYTES + 1024 <= RING_SMEM_BYTES, "ring does not fit");
    static DeviceOnce once;
    {
        cudaError_t e = ensure_dyn_smem(once, ring_reduce_kernel<Op>, RING_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
    }
    const size_t tile_bytes = (size_t)g.npx * 3;
    const int cpt = (int)((tile_bytes + RR_CHUNK - 1) / RR_CHUNK);
    const long long total = (long long)((cpt + unit_chunks - 1) / unit_chunks) * g.B;
    int grid = num_sms;
    if ((long long)grid > total) grid = (int)total;
    ring_reduce_kernel<Op><<<grid, RR_GT + 32, RING_SMEM_BYTES, stream>>>(g, p, cpt, unit_chunks);
    return (int)cudaGetLastError();
}

// The per-tile kernels read the 1-in-16 sample of a tile two or three times.  Out of the tile that is one 48-byte group every
// 768 bytes: DRAM delivers 64-byte granules, so the sample costs a fifth of the whole tile in traffic, scattered.  The first
// ring pass has every group in shared memory anyway: the thread that holds a sample group copies it to a packed per-tile
// buffer (1/16 of the batch, written once, read back coalesced).
__device__ __forceinline__ void keep_sample_group(uint4* sample, int tile, int groups, int g, const uint32_t (&w)[12]) {
    if (is_sample_group(g, groups)) {
        uint4* dst = sample + ((size_t)tile * (size_t)(groups / SAMPLE_STRIDE) + (size_t)(g / SAMPLE_STRIDE)) * 3;
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
    }
}
// Sample block j of a tile from the packed buffer.
__device__ __forceinline__ void load_sample_block(const uint4* tile_sample, int j, uint32_t (&w)[12]) {
    const uint4 a = tile_sample[3 * (size_t)j], b = tile_sample[3 * (size_t)j + 1], c = tile_sample[3 * (size_t)j + 2];
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
}
// f(w) for every sample block of the tile (all threads of the CTA).
template <class F>
__device__ __forceinline__ void for_each_packed_sample(const uint4* tile_sample, int n_blk, F&& f) {
    for (int j = (int)threadIdx.x; j < n_blk; j += NT) {
        uint32_t w[12];
        load_sample_block(tile_sample, j, w);
        f(w);
    }
}

__device__ __forceinline__ void load_group_smem(const unsigned char* buf, bool active, uint32_t (&w)[12]) {
    if (active) {
        const uint4* v = reinterpret_cast<const uint4*>(buf + threadIdx.x * 48u);
        const uint4 a = v[0], b = v[1], c = v[2];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i] = 0xFFFFFFFFu;        // all white: never tissue
    }
}
// Packed RGB (R in byte 0) of pixel i of the 48-byte group at grp (16-byte aligned shared memory).
__device__ __forceinline__ uint32_t group_px(const unsigned char* grp, int i) {
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(grp);
    const int o = 3 * i, wi = o >> 2;
    const uint32_t lo = gw[wi], hi = gw[wi < 11 ? wi + 1 : 11];
    return __funnelshift_r(lo, hi, (o & 3) * 8) & 0x00FFFFFFu;
}
// Pushes the packed RGB of the pixels flagged in `bits` (16-bit mask over this lane's group, the warp's 32 groups start
// at wgrp) into the warp's queue and processes 32 queued pixels with proc(rgb) whenever that many are ready; all lanes call.
// A warp prefix sum of the per-lane counts gives every lane its slots; each lane drops one small descriptor (lane, pixel)
// per flagged pixel -- a short divergent loop of a few instructions --, then the warp expands the descriptors to RGB words
// with all lanes busy.  Only when the warp's pixels do not fit the queue at once (rare) do they go in round by round.
template <class P>
__device__ __forceinline__ void rq_push_flagged(WarpScratch& ws, unsigned bits, const unsigned char* wgrp, P&& proc) {
    const unsigned lane = threadIdx.x & 31u;
    if (!__any_sync(0xffffffffu, bits != 0u)) return;
    const unsigned c = __popc(bits);
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
    }
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    if (ws.qlen + total <= (unsigned)RQ_CAP) {
        unsigned* q = ws.w + ws.qlen;
        unsigned dst = incl - c;
        const unsigned tag = lane << 4;
        while (bits != 0u) {
            q[dst++] = tag | (unsigned)(__ffs(bits) - 1);
            bits &= bits - 1u;
        }
        __syncwarp();
#pragma unroll
        for (unsigned r = 0; r < 2; ++r) {
            const unsigned e = r * 32u + lane;
            if (e < total) {
                const unsigned d = q[e];
                q[e] = group_px(wgrp + (d >> 4) * 48u, (int)(d & 15u));     // each lane rewrites the slot it read
            }
        }
        ws.qlen += total;
        while (ws.qlen >= 32u) {
            __syncwarp();
            ws.qlen -= 32u;
            proc(true, ws.w[ws.qlen + lane]);
            __syncwarp();
        }
        return;
    }
    unsigned m;
    while ((m = __ballot_sync(0xffffffffu, bits != 0u)) != 0u) {
        if (bits != 0u) {
            ws.w[ws.qlen + __popc(m & ((1u << lane) - 1u))] = group_px(wgrp + lane * 48u, __ffs(bits) - 1);
            bits &= bits - 1u;
        }
        ws.qlen += __popc(m);                             // < 32 on entry: at most 63 entries
        if (ws.qlen >= 32u) {
            __syncwarp();
            ws.qlen -= 32u;
            proc(true, ws.w[ws.qlen + lane]);
            __syncwarp();
        }
    }
}
template <class P>
__device__ __forceinline__ void rq_drain_rest(WarpScratch& ws, P&& proc) {
    if (ws.qlen > 0u) {
        __syncwarp();
        const bool has = (threadIdx.x & 31u) < ws.qlen;
        proc(has, has ? ws.w[threadIdx.x & 31u] : 0x00FFFFFFu);
        ws.qlen = 0u;
        __syncwarp();
    }
}
// Key-list staging buffer j (0 / 1) of the warp: append the keys of the lanes with pred; 32 staged keys go to the tile's
// global list with one atomic reservation.  n is warp-uniform.
__device__ __forceinline__ void stage_flush32(unsigned* buf, unsigned& n, unsigned* glen, unsigned* glist, unsigned cap) {
    __syncwarp();
    unsigned basei = 0;
    if ((threadIdx.x & 31) == 0) basei = atomicAdd(glen, 32u);
    basei = __shfl_sync(0xffffffffu, basei, 0);
    const unsigned idx = basei + (threadIdx.x & 31u);
    if (idx < cap) glist[idx] = buf[threadIdx.x & 31u];
    const unsigned rest = n - 32u;                       // < 32
    const unsigned carry = (threadIdx.x & 31u) < rest ? buf[32u + (threadIdx.x & 31u)] : 0u;
    __syncwarp();
    if ((threadIdx.x & 31u) < rest) buf[threadIdx.x & 31u] = carry;
    n = rest;
    __syncwarp();
}
__device__ __forceinline__ void stage_append(unsigned* buf, unsigned& n, bool pred, unsigned key, unsigned* glen, unsigned* glist, unsigned cap) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m) {
        if (pred) buf[n + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = key;
        n += __popc(m);
        if (n >= 32u) stage_flush32(buf, n, glen, glist, cap);
    }
}
__device__ __forceinline__ void stage_flush_rest(unsigned* buf, unsigned& n, unsigned* glen, unsigned* glist, unsigned cap) {
    if (n > 0u) {
        __syncwarp();
        unsigned basei = 0;
        if ((threadIdx.x & 31) == 0) basei = atomicAdd(glen, n);
        basei = __shfl_sync(0xffffffffu, basei, 0);
        const unsigned idx = basei + (threadIdx.x & 31u);
        if ((threadIdx.x & 31u) < n && idx < cap) glist[idx] = buf[threadIdx.x & 31u];
        n = 0u;
        __syncwarp();
    }
}

// The tissue bits of a group live in global memory (2 bytes per group, written by the first pass).  A load issued when the chunk
// is taken up is needed a few instructions later and every warp of the CTA waits out the L2 latency at once; so a thread asks
// for the bits of ITS group of the NEXT chunk one chunk ahead (same tile, 512 groups on) and finds them in a register when it
// gets there.  First chunk of a run: plain load.
struct MaskPrefetch { unsigned idx, val; };
__device__ __forceinline__ unsigned mask_bits_prefetched(const unsigned short* __restrict__ mask, MaskPrefetch& pf, unsigned tile, unsigned groups, unsigned g) {
    const unsigned idx = tile * groups + g;
    const unsigned bits = pf.idx == idx ? pf.val : (unsigned)mask[idx];
    if (g + (unsigned)RR_GT < groups) { pf.val = mask[idx + RR_GT]; pf.idx = idx + RR_GT; } else pf.idx = 0xFFFFFFFFu;
    return bits;
}

// ------------------------------------------------------------------------------------------------ pass 1: moments
struct StreamParams {
    TileState* state;
    const AngleConsts* aconsts;
    const ConcConsts* cconsts;
    unsigned* lists;              // [B][2][list_cap]
    int list_cap;
    struct DictState* dstate;     // Vahadane: per-tile dictionary, sums of the current pass, Anderson history
    const struct DictConsts* dconsts;
    unsigned short* mask;         // [B][groups]: 16 tissue bits per 16-pixel group, written by pass 1, read by pass 3
    int groups;                   // 16-pixel groups per tile
    uint4* sample;                // [B][groups / SAMPLE_STRIDE][3]: the 1-in-16 sample groups of every tile, packed (written by the first pass)
    const float* od;
    const unsigned short* gamma;
    float ycoef[3], ybound;
};
struct Empty {};

struct MomentOp {
    using Consts = Empty;
    using Params = StreamParams;
    struct Acc { long long s[9]; unsigned cnt; };
    static constexpr int kLaneShift = 3;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float2*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 8) = make_float2(p.od[i >> 5], (float)p.gamma[i >> 5]);
    }
    __device__ static Consts load_consts(const Params&, int) { return Empty{}; }
    __device__ static void acc_init(Acc& a) {
#pragma unroll
        for (int i = 0; i < 9; ++i) a.s[i] = 0;
        a.cnt = 0;
    }
    // accum_if_tissue with the count replaced by the pixel's bit of the group's tissue mask (same instruction count)
    __device__ static __forceinline__ void accum_mask(float y, float bound, float o0, float o1, float o2, float (&f)[9], unsigned& mbits, unsigned bit) {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.lt.f32 p, %10, %11;\n\t"
            "@p add.f32 %0, %0, %12;\n\t"
            "@p add.f32 %1, %1, %13;\n\t"
            "@p add.f32 %2, %2, %14;\n\t"
            "@p fma.rn.f32 %3, %12, %12, %3;\n\t"
            "@p fma.rn.f32 %4, %12, %13, %4;\n\t"
            "@p fma.rn.f32 %5, %12, %14, %5;\n\t"
            "@p fma.rn.f32 %6, %13, %13, %6;\n\t"
            "@p fma.rn.f32 %7, %13, %14, %7;\n\t"
            "@p fma.rn.f32 %8, %14, %14, %8;\n\t"
            "@p or.b32 %9, %9, %15;\n\t}"
            : "+f"(f[0]), "+f"(f[1]), "+f"(f[2]), "+f"(f[3]), "+f"(f[4]), "+f"(f[5]), "+f"(f[6]), "+f"(f[7]), "+f"(f[8]), "+r"(mbits)
            : "f"(y), "f"(bound), "f"(o0), "f"(o1), "f"(o2), "r"(bit));
    }
    __device__ static bool tile_active(const Params&, int) { return true; }
    __device__ static void process(const Consts&, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned px0, WarpScratch&,
                                   Acc& acc, int tile, bool) {
        if (!active) return;
        uint32_t w[12];
        load_group_smem(buf, true, w);
        keep_sample_group(p.sample, tile, p.groups, (int)(px0 >> 4) + (int)threadIdx.x, w);      // (first: the stores drain behind the arithmetic)
        const YCoef yc{p.ycoef[0], p.ycoef[1], p.ycoef[2], p.ybound};
        float f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = 0.f;
        unsigned mbits = 0;
        for_each_px_odg_abs(tab, w, [&](int i, float2 r, float2 g, float2 b) {
            accum_mask(tissue_y(yc, r.y, g.y, b.y), yc.bound, r.x, g.x, b.x, f, mbits, 1u << i);
        });
        acc.cnt += __popc(mbits);
        p.mask[(unsigned)tile * (unsigned)p.groups + (px0 >> 4) + threadIdx.x] = (unsigned short)mbits;      // pass 3 does not recompute the mask
#pragma unroll
        for (int i = 0; i < 9; ++i) acc.s[i] += to_fix(f[i], FIX_MOMENT);
    }
    __device__ static void finish_run(const Consts&, const Params& p, const OdAbs, WarpScratch&, Acc& acc, int tile) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const long long v = warp_sum_ll(acc.s[i]);
            if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&p.state[tile].mom[i], (unsigned long long)v);
        }
        const unsigned c = warp_sum_u(acc.cnt);
        if ((threadIdx.x & 31) == 0 && c != 0) atomicAdd(&p.state[tile].mom[9], (unsigned long long)c);
        acc_init(acc);
    }
};

// ------------------------------------------------------------------------------------------------ pass 3: angle brackets
// Per tissue pixel: does its angle lie safely BETWEEN the two brackets (then it only counts as "below bracket 1")?  The
// wedge between the brackets is the intersection of two half-planes through the origin of the (px, py) plane, i.e. of two
// half-spaces n1 . od > 0, n2 . od > 0 in optical-density space: six FFMA2 for two pixels.  The tissue bits come from
// pass 1 (2 bytes per group), so the table holds densities only and the pixels of a pair land in adjacent registers for
// the packed pipe.  The float tests carry an absolute margin (3e-5 >> the rounding error of the dot products, < 4e-6 for
// any uint8 pixel) on top of the 64 key units of slack in the wedge itself, so a pixel that passes is between the
// brackets in exact-key terms too; all others (~4 %: the two tails and the brackets) get their exact 23-bit key from the
// SAME arithmetic as in the fused kernel.
struct AngleOp {
    using Consts = AngleConsts;
    using Params = StreamParams;
    struct Acc { unsigned below0, below1; };
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = p.od[i >> 5];
    }
    __device__ static Consts load_consts(const Params& p, int tile) { return p.aconsts[tile]; }
    __device__ static void acc_init(Acc& a) { a.below0 = a.below1 = 0; }
    // exact treatment of one (tissue) pixel given as packed RGB: key, below counters, staged list appends; all lanes call
    __device__ static __forceinline__ void exact(const Consts& k, const Params& p, const OdAbs tab, WarpScratch& ws, Acc& acc, int tile, bool has,
                                                 uint32_t rgb) {
        const float r = od_lookup(tab, rgb, 0u, 0), g = od_lookup(tab, rgb, 0u, 1), b = od_lookup(tab, rgb, 0u, 2);
        const float px = fmaf(b, k.v[2], fmaf(g, k.v[1], r * k.v[0]));
        const float py = fmaf(b, k.v[5], fmaf(g, k.v[4], r * k.v[3]));
        const uint32_t key = angle_key(px, py);
        if (has && key < k.ka0) ++acc.below0;
        if (has && key < k.ka1) ++acc.below1;
        unsigned* len = p.state[tile].len;
        unsigned* list0 = p.lists + (size_t)tile * 2 * p.list_cap;
        stage_append(ws.w + RQ_CAP, ws.n0, has && key >= k.ka0 && key < k.kb0, key, &len[0], list0, (unsigned)p.list_cap);
        stage_append(ws.w + 2 * RQ_CAP, ws.n1, has && key >= k.ka1 && key < k.kb1, key, &len[1], list0 + p.list_cap, (unsigned)p.list_cap);
    }
    __device__ static bool tile_active(const Params& p, int tile) { return p.aconsts[tile].mode == 0; }
    __device__ static void process(const Consts& k, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned px0,
                                   WarpScratch& ws, Acc& acc, int tile, bool) {
        if (k.mode != 0) return;                        // warp-uniform (per-tile constant)
        unsigned mbits = 0;
        if (active) mbits = p.mask[(unsigned)tile * (unsigned)p.groups + (px0 >> 4) + threadIdx.x];     // issued first: the group's arithmetic hides it (< 2^26 groups per round)
        uint32_t w[12];
        load_group_smem(buf, active, w);
        const float2 a0 = dup(k.n1[0]), a1 = dup(k.n1[1]), a2 = dup(k.n1[2]);
        const float2 b0 = dup(k.n2[0]), b1 = dup(k.n2[1]), b2 = dup(k.n2[2]);
        const float margin = k.margin;
        uint32_t fastbits = 0;
        for_each_pair_od_abs(tab, w, [&](int i, float2 o0, float2 o1, float2 o2) {
            const float2 h1 = __ffma2_rn(o2, a2, __ffma2_rn(o1, a1, __fmul2_rn(o0, a0)));
            const float2 h2 = __ffma2_rn(o2, b2, __ffma2_rn(o1, b1, __fmul2_rn(o0, b0)));
            or_if_gt(fastbits, fminf(h1.x, h2.x), margin, 1u << i);
            or_if_gt(fastbits, fminf(h1.y, h2.y), margin, 2u << i);
        });
        acc.below1 += __popc(mbits & fastbits);
        rq_push_flagged(ws, mbits & ~fastbits, buf + (threadIdx.x & ~31u) * 48u, [&](bool has, uint32_t rgb) { exact(k, p, tab, ws, acc, tile, has, rgb); });
    }
    __device__ static void finish_run(const Consts& k, const Params& p, const OdAbs tab, WarpScratch& ws, Acc& acc, int tile) {
        if (k.mode == 0) {
            rq_drain_rest(ws, [&](bool has, uint32_t rgb) { exact(k, p, tab, ws, acc, tile, has, rgb); });
            unsigned* list0 = p.lists + (size_t)tile * 2 * p.list_cap;
            stage_flush_rest(ws.w + RQ_CAP, ws.n0, &p.state[tile].len[0], list0, (unsigned)p.list_cap);
            stage_flush_rest(ws.w + 2 * RQ_CAP, ws.n1, &p.state[tile].len[1], list0 + p.list_cap, (unsigned)p.list_cap);
            const unsigned b0 = warp_sum_u(acc.below0), b1 = warp_sum_u(acc.below1);
            if ((threadIdx.x & 31) == 0) {
                if (b0) atomicAdd(&p.state[tile].below[0], b0);
                if (b1) atomicAdd(&p.state[tile].below[1], b1);
            }
        }
        acc_init(acc);
    }
};

// ------------------------------------------------------------------------------------------------ pass 5: concentration brackets
// Per pixel and stain j: m_j = min(a_j, u_j) (max for stain vectors with a negative dot product) is the concentration
// before its clamp at zero; c_j < lo <=> m_j < lo for any lo > 0.  With x_j = m_j - mid_j:
//   |x_j| <= half_j  -> the pixel is "slow" for stain j: its exact key decides (below the bracket / listed / above);
//   else x_j < 0     -> below bracket j.
// The kernel is bound by the ALU pipe (compares, min/max, logic: 16 lanes per clock), so the test is built to spend FMA-pipe
// instructions instead: the main loop counts the NEGATIVE x_j of all pixels with a saturating multiply (x * -2^126 clamps
// to exactly 1.0 or 0.0) and a float add -- no compare, no select --, flags the slow pixels with two compares, and the
// exact path takes a slow pixel's sign back out of the count before its key decides.  Both places run the SAME packed
// arithmetic (the exact path duplicates its pixel into both halves), so they never disagree.
struct ConcOp {
    using Consts = ConcConsts;
    using Params = StreamParams;
    struct Acc { float n0, n1; int corr0, corr1; };
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = p.od[i >> 5];
    }
    __device__ static Consts load_consts(const Params& p, int tile) { return p.cconsts[tile]; }
    __device__ static void acc_init(Acc& a) { a.n0 = a.n1 = 0.f; a.corr0 = a.corr1 = 0; }
    __device__ static __forceinline__ float neg_as_one(float x) {      // 1.0 if x < 0 else 0.0 (denormals flushed), on the FMA pipe
        float s;
        asm("mul.rn.sat.ftz.f32 %0, %1, %2;" : "=f"(s) : "f"(x), "f"(-8.507059173023462e37f));
        return s;
    }
    // {x0, x1} of two pixels: packed LASSO without the clamp at zero, minus the window centres
    template <int LM>
    __device__ static __forceinline__ void window_coords(const Consts& k, const float2 o0, const float2 o1, const float2 o2, float2& m0, float2& m1,
                                                         float2& x0, float2& x1) {
        const LassoK& lk = k.lk;
        const float2 u0 = __ffma2_rn(dup(lk.m02), o2, __ffma2_rn(dup(lk.m01), o1, __ffma2_rn(dup(lk.m00), o0, dup(-lk.lam))));
        const float2 u1 = __ffma2_rn(dup(lk.m12), o2, __ffma2_rn(dup(lk.m11), o1, __ffma2_rn(dup(lk.m10), o0, dup(-lk.lam))));
        const float2 a0 = __ffma2_rn(dup(lk.i01), u1, __fmul2_rn(dup(lk.i00), u0));
        const float2 a1 = __ffma2_rn(dup(lk.i11), u1, __fmul2_rn(dup(lk.i01), u0));
        if (LM == LASSO_UNIT_POS) { m0 = f2(fminf(a0.x, u0.x), fminf(a0.y, u0.y)); m1 = f2(fminf(a1.x, u1.x), fminf(a1.y, u1.y)); }
        else { m0 = f2(fmaxf(a0.x, u0.x), fmaxf(a0.y, u0.y)); m1 = f2(fmaxf(a1.x, u1.x), fmaxf(a1.y, u1.y)); }
        x0 = __fadd2_rn(m0, dup(-k.mid0));
        x1 = __fadd2_rn(m1, dup(-k.mid1));
    }
    template <int LM>
    __device__ static __forceinline__ void exact(const Consts& k, const Params& p, const OdAbs tab, WarpScratch& ws, Acc& acc, int tile, bool has,
                                                 uint32_t v) {
        const float o0 = od_lookup(tab, v, 0u, 0), o1 = od_lookup(tab, v, 0u, 1), o2 = od_lookup(tab, v, 0u, 2);
        float2 m0, m1, x0, x1;
        window_coords<LM>(k, dup(o0), dup(o1), dup(o2), m0, m1, x0, x1);
        const bool s0 = has && fabsf(x0.x) <= k.half0, s1 = has && fabsf(x1.x) <= k.half1;
        const uint32_t k0 = conc_key(fmaxf(m0.x, 0.f)), k1 = conc_key(fmaxf(m1.x, 0.f));
        if (s0) { if (neg_as_one(x0.x) != 0.f) --acc.corr0; if (k0 < k.ka0) ++acc.corr0; }
        if (s1) { if (neg_as_one(x1.x) != 0.f) --acc.corr1; if (k1 < k.ka1) ++acc.corr1; }
        unsigned* len = p.state[tile].clen;
        unsigned* list0 = p.lists + (size_t)tile * 2 * p.list_cap;
        stage_append(ws.w + RQ_CAP, ws.n0, s0 && k0 >= k.ka0 && k0 < k.kb0, k0, &len[0], list0, (unsigned)p.list_cap);
        stage_append(ws.w + 2 * RQ_CAP, ws.n1, s1 && k1 >= k.ka1 && k1 < k.kb1, k1, &len[1], list0 + p.list_cap, (unsigned)p.list_cap);
    }
    template <int LM>
    __device__ static __forceinline__ void body(const Consts& k, const Params& p, const OdAbs tab, const unsigned char* buf, bool active,
                                                WarpScratch& ws, Acc& acc, int tile) {
        uint32_t w[12];
        load_group_smem(buf, active, w);
        const float half0 = k.half0, half1 = k.half1;
        float n0 = 0.f, n1 = 0.f;
        unsigned slow = 0;
        for_each_pair_od_abs(tab, w, [&](int i, float2 o0, float2 o1, float2 o2) {
            float2 m0, m1, x0, x1;
            window_coords<LM>(k, o0, o1, o2, m0, m1, x0, x1);
            n0 += neg_as_one(x0.x); n0 += neg_as_one(x0.y);
            n1 += neg_as_one(x1.x); n1 += neg_as_one(x1.y);
            asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t"
                "abs.f32 t, %1;\n\t setp.le.f32 p, t, %3;\n\t"
                "abs.f32 t, %2;\n\t setp.le.or.f32 p, t, %4, p;\n\t"
                "@p or.b32 %0, %0, %5;\n\t}"
                : "+r"(slow) : "f"(x0.x), "f"(x1.x), "f"(half0), "f"(half1), "r"(1u << i));
            asm("{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t"
                "abs.f32 t, %1;\n\t setp.le.f32 p, t, %3;\n\t"
                "abs.f32 t, %2;\n\t setp.le.or.f32 p, t, %4, p;\n\t"
                "@p or.b32 %0, %0, %5;\n\t}"
                : "+r"(slow) : "f"(x0.y), "f"(x1.y), "f"(half0), "f"(half1), "r"(2u << i));
        });
        if (active) { acc.n0 += n0; acc.n1 += n1; } else slow = 0;
        rq_push_flagged(ws, slow, buf + (threadIdx.x & ~31u) * 48u, [&](bool has, uint32_t rgb) { exact<LM>(k, p, tab, ws, acc, tile, has, rgb); });
    }
    __device__ static bool tile_active(const Params& p, int tile) { return p.cconsts[tile].mode == 0; }
    __device__ static void process(const Consts& k, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned,
                                   WarpScratch& ws, Acc& acc, int tile, bool) {
        if (k.mode != 0) return;
        if (k.lm == LASSO_UNIT_POS) body<LASSO_UNIT_POS>(k, p, tab, buf, active, ws, acc, tile);
        else body<LASSO_UNIT_NEG>(k, p, tab, buf, active, ws, acc, tile);
    }
    __device__ static void finish_run(const Consts& k, const Params& p, const OdAbs tab, WarpScratch& ws, Acc& acc, int tile) {
        if (k.mode == 0) {
            if (k.lm == LASSO_UNIT_POS) rq_drain_rest(ws, [&](bool has, uint32_t rgb) { exact<LASSO_UNIT_POS>(k, p, tab, ws, acc, tile, has, rgb); });
            else rq_drain_rest(ws, [&](bool has, uint32_t rgb) { exact<LASSO_UNIT_NEG>(k, p, tab, ws, acc, tile, has, rgb); });
            unsigned* list0 = p.lists + (size_t)tile * 2 * p.list_cap;
            stage_flush_rest(ws.w + RQ_CAP, ws.n0, &p.state[tile].clen[0], list0, (unsigned)p.list_cap);
            stage_flush_rest(ws.w + 2 * RQ_CAP, ws.n1, &p.state[tile].clen[1], list0 + p.list_cap, (unsigned)p.list_cap);
            // (the float counts are exact integers: at most 16 per chunk and thread, far below 2^24 per run)
            int t0 = (int)acc.n0 + acc.corr0, t1 = (int)acc.n1 + acc.corr1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_down_sync(0xffffffffu, t0, o); t1 += __shfl_down_sync(0xffffffffu, t1, o); }
            if ((threadIdx.x & 31) == 0) {
                if (t0) atomicAdd(&p.state[tile].cbelow[0], (unsigned)t0);
                if (t1) atomicAdd(&p.state[tile].cbelow[1], (unsigned)t1);
            }
        }
        acc_init(acc);
    }
};

// ------------------------------------------------------------------------------------------------ Vahadane passes
// V0: tissue bits of every group (kept for all dictionary passes), tissue count of the tile and of its 1-in-16 sample.
struct MaskOp {
    using Consts = Empty;
    using Params = StreamParams;
    struct Acc { unsigned tissue, sample; };
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = (float)p.gamma[i >> 5];
    }
    __device__ static Consts load_consts(const Params&, int) { return Empty{}; }
    __device__ static bool tile_active(const Params&, int) { return true; }
    __device__ static void acc_init(Acc& a) { a.tissue = a.sample = 0; }
    __device__ static void process(const Consts&, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned px0, WarpScratch&,
                                   Acc& acc, int tile, bool) {
        if (!active) return;
        uint32_t w[12];
        load_group_smem(buf, true, w);
        keep_sample_group(p.sample, tile, p.groups, (int)(px0 >> 4) + (int)threadIdx.x, w);
        const float2 cr = dup(p.ycoef[0]), cg = dup(p.ycoef[1]), cb = dup(p.ycoef[2]);
        const float bound = p.ybound;
        unsigned mbits = 0;
        // Y = 871 g[R] + 2929 g[G] + 296 g[B]: integers below 2^24, exact in fp32 in any order -- two pixels per FFMA2
        for_each_pair_od_abs(tab, w, [&](int i, float2 g0, float2 g1, float2 g2) {
            const float2 y = __ffma2_rn(g2, cb, __ffma2_rn(g1, cg, __fmul2_rn(g0, cr)));
            or_if_lt(mbits, y.x, bound, 1u << i);
            or_if_lt(mbits, y.y, bound, 2u << i);
        });
        const int g = (int)(px0 >> 4) + (int)threadIdx.x;
        p.mask[(unsigned)tile * (unsigned)p.groups + (unsigned)g] = (unsigned short)mbits;
        const unsigned c = __popc(mbits);
        acc.tissue += c;
        if (is_sample_group(g, p.groups)) acc.sample += c;
    }
    __device__ static void finish_run(const Consts&, const Params& p, const OdAbs, WarpScratch&, Acc& acc, int tile) {
        const unsigned t = warp_sum_u(acc.tissue), sm = warp_sum_u(acc.sample);
        if ((threadIdx.x & 31) == 0) {
            if (t) atomicAdd(&p.state[tile].mom[9], (unsigned long long)t);
            if (sm) atomicAdd(&p.state[tile].mom[0], (unsigned long long)sm);
        }
        acc_init(acc);
    }
};

// get_tissue_mask (stain_utils.py:32-48) as a streaming pass: 3 B/px in through the ring, 1 B/px out (16 bytes per thread,
// coalesced); a tile that shows tissue clears the EMPTY_MASK bit its status word was preset to.
struct MaskOutParams {
    const unsigned short* gamma;
    float ycoef[3], ybound;
    uint8_t* mask_out;            // [B][npx], 1 = tissue
    int32_t* status;              // [B] or null
    int npx;
};
struct MaskOutOp {
    using Consts = Empty;
    using Params = MaskOutParams;
    struct Acc { unsigned any; };
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = (float)p.gamma[i >> 5];
    }
    __device__ static Consts load_consts(const Params&, int) { return Empty{}; }
    __device__ static bool tile_active(const Params&, int) { return true; }
    __device__ static void acc_init(Acc& a) { a.any = 0; }
    __device__ static void process(const Consts&, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned px0, WarpScratch&,
                                   Acc& acc, int tile, bool) {
        if (!active) return;
        uint32_t w[12];
        load_group_smem(buf, true, w);
        const float2 cr = dup(p.ycoef[0]), cg = dup(p.ycoef[1]), cb = dup(p.ycoef[2]);
        const float bound = p.ybound;
        unsigned mbits = 0;
        for_each_pair_od_abs(tab, w, [&](int i, float2 g0, float2 g1, float2 g2) {
            const float2 y = __ffma2_rn(g2, cb, __ffma2_rn(g1, cg, __fmul2_rn(g0, cr)));
            or_if_lt(mbits, y.x, bound, 1u << i);
            or_if_lt(mbits, y.y, bound, 2u << i);
        });
        acc.any |= mbits;
        // four bits -> four 0/1 bytes: the multiply puts bit j of the nibble at bit 8 j (no two products collide)
        uint4 v;
        v.x = ((mbits & 15u) * 0x00204081u) & 0x01010101u;
        v.y = (((mbits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
        v.z = (((mbits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
        v.w = (((mbits >> 12) & 15u) * 0x00204081u) & 0x01010101u;
        stg_stream(reinterpret_cast<uint4*>(p.mask_out + (size_t)tile * p.npx + px0) + threadIdx.x, v);
    }
    __device__ static void finish_run(const Consts&, const Params& p, const OdAbs, WarpScratch&, Acc& acc, int tile) {
        const unsigned any = __ballot_sync(0xffffffffu, acc.any != 0u);
        if ((threadIdx.x & 31) == 0 && any && p.status) atomicAnd(&p.status[tile], ~SB_STATUS_EMPTY_MASK);
        acc.any = 0;
    }
};

// get_concentrations (stain_utils.py:69-78) as a streaming pass: 3 B/px in through the ring, 8 B/px out.  A thread that kept
// "its" 16-pixel group would write 16 float2 = 128 bytes at a 128-byte stride from its neighbour lanes (32 sectors per store
// instruction); here the lanes of a warp take CONSECUTIVE pixels of the warp's 512-pixel stretch of the chunk (three byte loads
// from shared memory: 96 contiguous bytes per warp, broadcast within a word, no bank conflicts), so every store instruction
// writes 256 contiguous bytes.  The arithmetic is lasso2 of the register-staged concentrations_kernel: same bits.
struct ConcOutParams {
    const float* od;
    const double* M;              // [B][2][3]
    double lasso_lambda;
    float* conc_out;              // [B][npx][2]
    int npx;
};
struct ConcOutOp {
    using Consts = LassoK;
    using Params = ConcOutParams;
    struct Acc {};
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = p.od[i >> 5];
    }
    __device__ static Consts load_consts(const Params& p, int tile) {
        double M[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) M[q] = p.M[(size_t)tile * 6 + q];
        LassoK lk;
        make_lasso_consts(M, p.lasso_lambda, lk);
        return lk;
    }
    __device__ static bool tile_active(const Params&, int) { return true; }
    __device__ static void acc_init(Acc&) {}
    __device__ static void process(const Consts& lk, const Params& p, const OdAbs tab, const unsigned char* buf, bool, unsigned px0, WarpScratch&,
                                   Acc&, int tile, bool) {
        const unsigned rest = (unsigned)p.npx - px0;                              // pixels of the tile from this chunk on
        const unsigned n_chunk = rest < (unsigned)(RR_GT * GROUP_PX) ? rest : (unsigned)(RR_GT * GROUP_PX);
        const unsigned w0 = (threadIdx.x & ~31u) * GROUP_PX;                        // the warp's stretch of the chunk: 512 pixels
        float2* out = reinterpret_cast<float2*>(p.conc_out) + (size_t)tile * (size_t)p.npx + px0;
#pragma unroll 4
        for (unsigned i = 0; i < GROUP_PX; ++i) {
            const unsigned j = w0 + i * 32u + (threadIdx.x & 31u);
            if (j < n_chunk) {
                const unsigned char* b = buf + 3u * j;
                const uint32_t r = b[0], g = b[1], bl = b[2];
                float c0, c1;
                lasso2(lk, od_lookup(tab, r, 0u, 0), od_lookup(tab, g, 0u, 0), od_lookup(tab, bl, 0u, 0), c0, c1);
                asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(out + j), "f"(c0), "f"(c1) : "memory");
            }
        }
    }
    __device__ static void finish_run(const Consts&, const Params&, const OdAbs, WarpScratch&, Acc&, int) {}
};

// One full dictionary pass: sparse-code the tissue pixels under the tile's current dictionary, accumulate a a^T and x a^T.
// The partial sums are the ones of the fused kernel, bit for bit: fp32 per thread over the groups of one UNIT (the ring
// hands units to CTAs whole, and thread t of a chunk holds the group thread t of the fused kernel visits), reduced over
// the warp by the same shuffle tree, then fixed point.
struct DictOp {
    using Consts = DictConsts;
    using Params = StreamParams;
    struct Acc { float2 f[9]; long long s[9]; MaskPrefetch pf; };
    static constexpr int kLaneShift = 2;
    __device__ static void fill_table(unsigned char* tab, const Params& p, int tid, int n) {
        for (int i = tid; i < 256 * 32; i += n)
            *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = p.od[i >> 5];
    }
    __device__ static Consts load_consts(const Params& p, int tile) { return p.dconsts[tile]; }
    __device__ static bool tile_active(const Params& p, int tile) { return p.dconsts[tile].mode == 0; }
    __device__ static void acc_init(Acc& a) {
#pragma unroll
        for (int i = 0; i < 9; ++i) { a.f[i] = make_float2(0.f, 0.f); a.s[i] = 0; }
        a.pf.idx = 0xFFFFFFFFu; a.pf.val = 0u;
    }
    template <int LM>
    __device__ static __forceinline__ void add_unit(const LassoK& lk, const OdAbs tab, const uint32_t (&w)[12], uint32_t mbits, float2 (&f)[9]) {
        for_each_pair_od_abs(tab, w, [&](int i, float2 o0, float2 o1, float2 o2) {
            float2 c0, c1;
            lasso2_unit_pair<LM>(lk, o0, o1, o2, c0, c1);
            const bool ma = (mbits & (1u << i)) != 0, mb = (mbits & (2u << i)) != 0;
            c0.x = ma ? c0.x : 0.f; c1.x = ma ? c1.x : 0.f;
            c0.y = mb ? c0.y : 0.f; c1.y = mb ? c1.y : 0.f;
            f[0] = __ffma2_rn(c0, c0, f[0]); f[1] = __ffma2_rn(c0, c1, f[1]); f[2] = __ffma2_rn(c1, c1, f[2]);
            f[3] = __ffma2_rn(o0, c0, f[3]); f[4] = __ffma2_rn(o1, c0, f[4]); f[5] = __ffma2_rn(o2, c0, f[5]);
            f[6] = __ffma2_rn(o0, c1, f[6]); f[7] = __ffma2_rn(o1, c1, f[7]); f[8] = __ffma2_rn(o2, c1, f[8]);
        });
    }
    __device__ static void process(const Consts& k, const Params& p, const OdAbs tab, const unsigned char* buf, bool active, unsigned px0, WarpScratch&,
                                   Acc& acc, int tile, bool unit_end) {
        if (active) {
            const uint32_t mbits = mask_bits_prefetched(p.mask, acc.pf, (unsigned)tile, (unsigned)p.groups, (px0 >> 4) + threadIdx.x);
            uint32_t w[12];
            load_group_smem(buf, true, w);
            if (k.lm == LASSO_UNIT_POS) add_unit<LASSO_UNIT_POS>(k.lk, tab, w, mbits, acc.f);
            else add_unit<LASSO_UNIT_NEG>(k.lk, tab, w, mbits, acc.f);
        }
        if (unit_end) {                                  // warp-uniform
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float v = acc.f[i].x + acc.f[i].y;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                acc.s[i] += to_fix(v, FIX_DL);            // (lane 0 holds the warp's sum)
                acc.f[i] = make_float2(0.f, 0.f);
            }
        }
    }
    __device__ static void finish_run(const Consts&, const Params& p, const OdAbs, WarpScratch&, Acc& acc, int tile) {
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i)
                if (acc.s[i] != 0) atomicAdd(&p.dstate[tile].sums[i], (unsigned long long)acc.s[i]);
        }
        acc_init(acc);
    }
};

// ------------------------------------------------------------------------------------------------ per-tile kernels
struct TileKernelArgs {
    PipeArgs a;
    TileState* state;
    AngleConsts* aconsts;
    ConcConsts* cconsts;
    unsigned* lists;
    int list_cap;                 // entries per key list (slist_cap(npx)); the kernels' second shared list sits at TILE_LIST1_OFFSET
    int* fb_list;                 // tiles for the fused kernel
    int* fb_count;
    const uint4* sample;          // packed 1-in-16 sample of every tile (StreamParams::sample)
};
struct TileShared {
    float2 tab[256];
    PipeShared ps;                // allocated without its fused-kernel-only tail (see PipeShared)
};
// Layout of the per-tile kernels' dynamic shared memory: TileShared up to PipeShared::aa, then the two key lists of
// list_cap full keys each (twice / four times the fused kernel's): 2 x 32 KB for tiles up to a megapixel -- list 0 then
// simply is the 32 KB histogram buffer --, 2 x 64 KB behind the structure for bigger tiles.
constexpr size_t TILE_LIST_OFFSET = (offsetof(TileShared, ps) + offsetof(PipeShared, aa) + 15) & ~(size_t)15;
__host__ __device__ inline size_t tile_shared_bytes(int list_cap) {
    return TILE_LIST_OFFSET + (size_t)list_cap * sizeof(unsigned) * (list_cap * sizeof(unsigned) > 2 * L1_BINS * sizeof(unsigned) ? 2 : 1);
}
struct TileLists { void* l0; void* l1; unsigned bytes; };
__device__ __forceinline__ TileLists tile_lists(unsigned char* smem_raw, PipeShared* sh, int list_cap) {
    const unsigned bytes = (unsigned)list_cap * (unsigned)sizeof(unsigned);
    if (bytes <= 2 * L1_BINS * sizeof(unsigned)) return TileLists{sh->hist, smem_raw + TILE_LIST_OFFSET, bytes};
    return TileLists{smem_raw + TILE_LIST_OFFSET, smem_raw + TILE_LIST_OFFSET + bytes, bytes};
}

// Half-width cap of a sampled bracket (in sample ranks) such that the bracket -- (2 m + 1) sample ranks of 16 pixels plus
// up to two level-1 bins (n / 4096 keys each) of rounding -- stays below ~7000 of the 8192 list entries.
__device__ inline double stream_bracket_cap(unsigned n, int list_cap) {
    const double m = 0.8 * (double)list_cap / 32.0 - 0.5;      // (2 m + 1) sample ranks of 16 pixels fill 80 % of the list
    return m > 80.0 ? m : 80.0;
}
constexpr double STREAM_WIDE_SIGMAS = 6.5;      // the lists of the streaming path afford wider brackets than the fused kernel's 5 sigma
// Bracket edges INSIDE the level-1 bins that hold the bracket's sample ranks: the keys of a bin are taken as uniformly spread
// over it (rank `rem` of `cnt` sample keys in bin `bin`), with 1/16 of a bin of padding.  Snapping the edges to whole bins,
// as the fused kernel does, costs up to two bins of extra keys in the list -- thousands on a megapixel tile, whose tissue
// angles crowd into a few hundred of the 4096 bins.  The bracket only has to CONTAIN the target rank (validated exactly
// after the pass) and fit the list.
__device__ inline unsigned bracket_edge_lo(unsigned bin, unsigned rem, unsigned cnt) {
    const unsigned base = bin << L2_BITS, off = cnt ? (unsigned)(((unsigned long long)rem << L2_BITS) / cnt) : 0u;
    const unsigned pad = 1u << (L2_BITS - 4);
    return base + off > pad ? base + off - pad : 0u;
}
__device__ inline unsigned bracket_edge_hi(unsigned bin, unsigned rem, unsigned cnt) {
    const unsigned base = bin << L2_BITS, off = cnt ? (unsigned)((((unsigned long long)rem + 1u) << L2_BITS) / cnt) + 1u : (1u << L2_BITS);
    const unsigned e = base + off + (1u << (L2_BITS - 4));
    return e < (1u << KEY_BITS) ? e : (1u << KEY_BITS);
}
__device__ __forceinline__ void fill_small_table(float2* tab, const Tables& t) {
    for (int i = threadIdx.x; i < 256; i += NT) tab[i] = make_float2(t.od[i], (float)t.gamma[i]);
}
// Why tiles left the streaming path, counted per process and device (diagnostics: sb_stream_fallbacks):
//   1 tile too small / too little tissue   2 angle sample too small   3 concentration sample too small
//   4 concentration bracket touches zero or the top of the key range   5 angle bracket missed its rank or its list overflowed
//   6 non-unit stain vectors   7 concentration bracket missed / overflowed
__device__ unsigned g_fallback_reasons[8];
// Thread 0: the tile leaves the streaming path.
__device__ inline void tile_to_fallback(const TileKernelArgs& k, int tile, int reason) {
    atomicAdd(&g_fallback_reasons[reason & 7], 1u);
    atomicAdd(&g_fallback_reasons[0], 1u);
    k.state[tile].path = PATH_FALLBACK;
    k.aconsts[tile].mode = 1;
    k.cconsts[tile].mode = 1;
    const int idx = atomicAdd(k.fb_count, 1);
    k.fb_list[idx] = tile;
}
__device__ inline void tile_flagged(const TileKernelArgs& k, int tile, int flags) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    k.state[tile].path = PATH_FLAGGED;
    k.state[tile].flags = flags;
    k.aconsts[tile].mode = 1;
    k.cconsts[tile].mode = 1;
    if (k.a.M_out) for (int i = 0; i < 6; ++i) k.a.M_out[(size_t)tile * 6 + i] = nan;
    if (k.a.mode >= PIPE_FIT && k.a.maxC_out) k.a.maxC_out[(size_t)tile * 2] = k.a.maxC_out[(size_t)tile * 2 + 1] = nan;
    if (k.a.status) k.a.status[tile] = flags;
}
// Loads a tile's global key list j into the 16 KB shared-memory list in the format list_select_pairs expects.
__device__ __forceinline__ void load_list(const KeyList& l, const unsigned* src, unsigned len) {
    for (unsigned i = threadIdx.x; i < len; i += NT) l.put(i, src[i]);
}

// 2: covariance + eigenvectors (macenko_stain_extractor.py:22-27), sample of the angle keys -> brackets (B0 of the fused kernel).
__global__ void __launch_bounds__(NT, 4) plan_angle_kernel(TileKernelArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared* ts = reinterpret_cast<TileShared*>(smem_raw);
    PipeShared* sh = &ts->ps;
    const PipeArgs& a = k.a;
    const int tile = blockIdx.x;
    const int npx = a.npx, G = npx / GROUP_PX;
    const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
    const YCoef yc{a.ycoef[0], a.ycoef[1], a.ycoef[2], a.ybound};
    TileState& st = k.state[tile];
    fill_small_table(ts->tab, a.tab);
    if (threadIdx.x == 0) {
        double t[10];
        for (int i = 0; i < 9; ++i) t[i] = (double)(long long)st.mom[i] * (1.0 / (double)FIX_MOMENT);
        t[9] = (double)st.mom[9];
        const double n = t[9];
        int flags = 0;
        if (n < 1.0) flags |= SB_STATUS_EMPTY_MASK;
        else if (n < 2.0) flags |= SB_STATUS_FEW_TISSUE;
        if (!flags) {
            const double inv = 1.0 / (n - 1.0);
            double c[6];
            c[0] = (t[3] - t[0] * t[0] / n) * inv; c[1] = (t[4] - t[0] * t[1] / n) * inv; c[2] = (t[5] - t[0] * t[2] / n) * inv;
            c[3] = (t[6] - t[1] * t[1] / n) * inv; c[4] = (t[7] - t[1] * t[2] / n) * inv; c[5] = (t[8] - t[2] * t[2] / n) * inv;
            double wv[3], v[3][3];
            jacobi_eig3(c, wv, v);
            int i1 = 0;
            if (wv[1] > wv[i1]) i1 = 1;
            if (wv[2] > wv[i1]) i1 = 2;
            int i2 = -1;
            for (int q = 0; q < 3; ++q) if (q != i1 && (i2 < 0 || wv[q] > wv[i2])) i2 = q;
            const double s1 = v[0][i1] < 0 ? -1.0 : 1.0, s2 = v[0][i2] < 0 ? -1.0 : 1.0;
            bool ok = true;
            for (int q = 0; q < 3; ++q) {
                const double x1 = s1 * v[q][i1], x2 = s2 * v[q][i2];
                sh->V[q] = (float)x1; sh->V[3 + q] = (float)x2;
                st.Vd[q] = x1; st.Vd[3 + q] = x2;
                ok = ok && isfinite(x1) && isfinite(x2);
            }
            if (!ok) flags |= SB_STATUS_DEGENERATE;
        }
        sh->flags = flags;
        st.n_tissue = (unsigned)n;
        sh->s_cnt = 0; sh->s_ok = 0;
        if (flags) tile_flagged(k, tile, flags);
        else if ((unsigned)n < 16384u || npx < 32768) { tile_to_fallback(k, tile, 1); sh->flags = -1; }
    }
    zero_hist(sh);
    if (sh->flags != 0) return;
    const unsigned n_tissue = st.n_tissue;
    const float v00 = sh->V[0], v01 = sh->V[1], v02 = sh->V[2], v10 = sh->V[3], v11 = sh->V[4], v12 = sh->V[5];
    unsigned p_lo[2], p_hi[2];
    { double fr; percentile_index(n_tissue, 100.0 - a.ang_pct, p_lo[0], p_hi[0], fr); percentile_index(n_tissue, a.ang_pct, p_lo[1], p_hi[1], fr); }
    unsigned scnt = 0;
    for_each_packed_sample(k.sample + (size_t)tile * (size_t)(G / SAMPLE_STRIDE) * 3, G / SAMPLE_STRIDE, [&](const uint32_t (&w)[12]) {
        for_each_px_odg_small(ts->tab, w, [&](int, float2 r, float2 g, float2 b) {
            const float px = fmaf(b.x, v02, fmaf(g.x, v01, r.x * v00));
            const float py = fmaf(b.x, v12, fmaf(g.x, v11, r.x * v10));
            if (tissue_y(yc, r.y, g.y, b.y) < yc.bound) { ++scnt; atomicAdd(&sh->hist[angle_key(px, py) >> L2_BITS], 1u); }
        });
    });
    scnt = warp_sum_u(scnt);
    if ((threadIdx.x & 31) == 0 && scnt) atomicAdd(&sh->s_cnt, scnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n_s = sh->s_cnt;
        if (n_s >= 1024u) {
            plan_bracket(n_tissue, n_s, p_lo[0], sh->q_rank[0], sh->q_rank[1], a.bracket_sigmas, a.bracket_pad, stream_bracket_cap(n_tissue, k.list_cap), STREAM_WIDE_SIGMAS);
            plan_bracket(n_tissue, n_s, p_lo[1], sh->q_rank[2], sh->q_rank[3], a.bracket_sigmas, a.bracket_pad, stream_bracket_cap(n_tissue, k.list_cap), STREAM_WIDE_SIGMAS);
            for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
            sh->s_ok = 1;
        } else {
            tile_to_fallback(k, tile, 2);
        }
    }
    __syncthreads();
    if (!sh->s_ok) return;
    select_ranks<L1_BINS>(sh, sh->hist, 1, sh->q_rank, 4, sh->q_bin, sh->q_rem);
    if (threadIdx.x == 0) {
        unsigned ka[2], kb[2];
        for (int j = 0; j < 2; ++j) {
            ka[j] = bracket_edge_lo(sh->q_bin[2 * j], sh->q_rem[2 * j], sh->hist[sh->q_bin[2 * j]]);
            kb[j] = bracket_edge_hi(sh->q_bin[2 * j + 1], sh->q_rem[2 * j + 1], sh->hist[sh->q_bin[2 * j + 1]]);
        }
        // a pixel in the half-plane x > 0 whose diamond coordinate lies safely between the low and the high bracket (64 key
        // units of slack) is "above bracket 0, below bracket 1" without computing its key (as in the fused kernel)
        const double d_lo = diamond_from_key((double)kb[0] + 64.0), d_hi = diamond_from_key((double)ka[1] - 64.0);
        const bool usable = ka[1] >= 64u && d_lo > -0.999 && d_hi < 0.999 && d_lo < d_hi;
        AngleConsts c;
        for (int q = 0; q < 6; ++q) c.v[q] = sh->V[q];
        // the ray of diamond coordinate d in the half-plane x >= 0 is (1 - |d|, d); "counter-clockwise of the low ray" is
        // -d_lo px + (1 - |d_lo|) py > 0, "clockwise of the high ray" is d_hi px - (1 - |d_hi|) py > 0; with px = V0 . od,
        // py = V1 . od (the float eigenvectors the exact keys use) these are two half-spaces of OD space
        for (int q = 0; q < 3; ++q) {
            const double v0 = (double)sh->V[q], v1 = (double)sh->V[3 + q];
            c.n1[q] = usable ? (float)(-d_lo * v0 + (1.0 - fabs(d_lo)) * v1) : 0.f;
            c.n2[q] = usable ? (float)(d_hi * v0 - (1.0 - fabs(d_hi)) * v1) : 0.f;
        }
        c.margin = usable ? 3e-5f : INFINITY;
        c.ka0 = ka[0]; c.kb0 = kb[0]; c.ka1 = ka[1]; c.kb1 = kb[1];
        c.mode = 0; c.pad[0] = c.pad[1] = 0;
        k.aconsts[tile] = c;
        st.brk[0] = ka[0]; st.brk[1] = kb[0]; st.brk[2] = ka[1]; st.brk[3] = kb[1];
    }
}

// C0 of the fused kernel: concentration keys of the 1-in-16 sample under the tile's stain matrix (sh->lk) -> brackets and
// float windows for pass 5.  Whole block calls; the hist buffer is free.
__device__ __forceinline__ void plan_conc_block(const TileKernelArgs& k, int tile, TileShared* ts, PipeShared* sh) {
    const PipeArgs& a = k.a;
    TileState& st = k.state[tile];
    const int npx = a.npx, G = npx / GROUP_PX;
    const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
    // ---- C0: concentration keys of the 1-in-16 sample
    zero_hist(sh);
    const LassoK lk = sh->lk;
    unsigned c_lo, c_hi;
    { double fr; percentile_index((unsigned)npx, a.conc_pct, c_lo, c_hi, fr); }
    unsigned scnt = 0;
    for_each_packed_sample(k.sample + (size_t)tile * (size_t)(G / SAMPLE_STRIDE) * 3, G / SAMPLE_STRIDE, [&](const uint32_t (&w)[12]) {
        scnt += GROUP_PX;
        for_each_px_odg_small(ts->tab, w, [&](int, float2 r, float2 g, float2 b) {
            float c0, c1;
            lasso2(lk, r.x, g.x, b.x, c0, c1);
            atomicAdd(&sh->hist[conc_key(c0) >> L2_BITS], 1u);
            atomicAdd(&sh->hist[L1_BINS + (conc_key(c1) >> L2_BITS)], 1u);
        });
    });
    scnt = warp_sum_u(scnt);
    if ((threadIdx.x & 31) == 0 && scnt) atomicAdd(&sh->s_cnt, scnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned n_s = sh->s_cnt;
        if (n_s >= 1024u) {
            plan_bracket((unsigned)npx, n_s, c_lo, sh->q_rank[0], sh->q_rank[1], a.bracket_sigmas, a.bracket_pad, stream_bracket_cap((unsigned)npx, k.list_cap), STREAM_WIDE_SIGMAS);
            sh->q_rank[2] = sh->q_rank[0]; sh->q_rank[3] = sh->q_rank[1];
            for (int q = 0; q < 4; ++q) { sh->q_bin[q] = 0; sh->q_rem[q] = 0; }
            sh->s_ok = 1;
        } else {
            tile_to_fallback(k, tile, 3);
        }
    }
    __syncthreads();
    if (!sh->s_ok) return;
    select_ranks<L1_BINS>(sh, sh->hist, 1, sh->q_rank, 2, sh->q_bin, sh->q_rem);
    select_ranks<L1_BINS>(sh, sh->hist + L1_BINS, 1, sh->q_rank + 2, 2, sh->q_bin + 2, sh->q_rem + 2);
    if (threadIdx.x == 0) {
        ConcConsts c;
        c.lk = sh->lk;
        unsigned ka[2], kb[2];
        float mid[2], half[2];
        bool ok = true;
        for (int j = 0; j < 2; ++j) {
            const unsigned* hj = sh->hist + j * L1_BINS;
            ka[j] = bracket_edge_lo(sh->q_bin[2 * j], sh->q_rem[2 * j], hj[sh->q_bin[2 * j]]);
            kb[j] = bracket_edge_hi(sh->q_bin[2 * j + 1], sh->q_rem[2 * j + 1], hj[sh->q_bin[2 * j + 1]]);
            // a bracket that touches zero (a stain absent from >= 99 % of the tile) or the top of the key range: robust path
            ok = ok && ka[j] >= 64u && kb[j] + 64u < (1u << KEY_BITS);
            // concentrations outside [lo, hi] (64 key units of slack around the bracket) need no exact key
            const double lo = ok ? (double)float_below(conc_from_key(ka[j] - 64u)) : 0.0;
            const double hi = ok ? (double)float_above(conc_from_key(kb[j] + 64u)) : 1.0;
            mid[j] = (float)(0.5 * (lo + hi));
            half[j] = float_above(fmax(hi - (double)mid[j], (double)mid[j] - lo) * (1.0 + 1e-6));
            ok = ok && lo > 0.0;
        }
        if (!ok) tile_to_fallback(k, tile, 4);
        else {
            c.mid0 = mid[0]; c.half0 = half[0]; c.mid1 = mid[1]; c.half1 = half[1];
            c.ka0 = ka[0]; c.kb0 = kb[0]; c.ka1 = ka[1]; c.kb1 = kb[1];
            c.lm = lasso_mode_of(c.lk.rg00, c.lk.rg11, c.lk.g01);
            c.mode = 0; c.pad = 0;
            k.cconsts[tile] = c;
            st.cbrk[0] = ka[0]; st.cbrk[1] = kb[0]; st.cbrk[2] = ka[1]; st.cbrk[3] = kb[1];
        }
    }
}

// 4: exact angular percentiles from the bracket lists -> stain matrix (macenko_stain_extractor.py:29-44); then the sample of
// the concentrations -> brackets (C0 of the fused kernel).
// MINB = resident CTAs per SM the register budget is cut for: 4 for the small-tile lists (52 KB of shared memory per CTA), 3 otherwise
template <int MINB>
__global__ void __launch_bounds__(NT, MINB) select_angle_kernel(TileKernelArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared* ts = reinterpret_cast<TileShared*>(smem_raw);
    PipeShared* sh = &ts->ps;
    const PipeArgs& a = k.a;
    const int tile = blockIdx.x;
    TileState& st = k.state[tile];
    if (st.path != PATH_STREAM) return;
    const int npx = a.npx, G = npx / GROUP_PX;
    const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
    const unsigned n_tissue = st.n_tissue;
    unsigned p_lo[2], p_hi[2];
    { double fr; percentile_index(n_tissue, 100.0 - a.ang_pct, p_lo[0], p_hi[0], fr); percentile_index(n_tissue, a.ang_pct, p_lo[1], p_hi[1], fr); }
    const unsigned ka0 = st.brk[0], kb0 = st.brk[1], ka1 = st.brk[2], kb1 = st.brk[3];
    const TileLists tl = tile_lists(smem_raw, sh, k.list_cap);
    const KeyList list0{tl.l0, ka0, kb0 - ka0 > LIST_SPAN, tl.bytes}, list1{tl.l1, ka1, kb1 - ka1 > LIST_SPAN, tl.bytes};
    fill_small_table(ts->tab, a.tab);
    if (threadIdx.x == 0) {
        bool ok = true;
        for (int j = 0; j < 2; ++j) {
            sh->l_len[j] = st.len[j]; sh->l_below[j] = st.below[j];
            ok = ok && st.len[j] <= min((j ? list1 : list0).cap(), (unsigned)k.list_cap) && st.below[j] <= p_lo[j] && p_hi[j] < st.below[j] + st.len[j];
        }
        sh->s_ok = ok ? 1 : 0;
        sh->flags = 0;
        if (!ok) tile_to_fallback(k, tile, 5);
    }
    __syncthreads();
    if (!sh->s_ok) return;
    const unsigned* gl = k.lists + (size_t)tile * 2 * k.list_cap;
    load_list(list0, gl, sh->l_len[0]);
    load_list(list1, gl + k.list_cap, sh->l_len[1]);
    __syncthreads();
    {
        const unsigned r_lo[2] = {p_lo[0] - sh->l_below[0], p_lo[1] - sh->l_below[1]};
        const unsigned r_hi[2] = {p_hi[0] - sh->l_below[0], p_hi[1] - sh->l_below[1]};
        unsigned keys[4];
        list_select_pairs(sh, list0, list1, r_lo, r_hi, keys);
        if (threadIdx.x < 4) sh->okey[threadIdx.x] = keys[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < 4) sh->ang[threadIdx.x] = angle_from_key(sh->okey[threadIdx.x]);
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned lo, hi; double fr;
        percentile_index(n_tissue, threadIdx.x == 0 ? 100.0 - a.ang_pct : a.ang_pct, lo, hi, fr);
        const double phi = lerp_np(sh->ang[2 * threadIdx.x], sh->ang[2 * threadIdx.x + 1], fr);
        sh->cs[2 * threadIdx.x] = cos(phi); sh->cs[2 * threadIdx.x + 1] = sin(phi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double c1 = sh->cs[0], s1 = sh->cs[1], c2 = sh->cs[2], s2 = sh->cs[3];
        double v1[3], v2[3];
        for (int q = 0; q < 3; ++q) {
            v1[q] = st.Vd[q] * c1 + st.Vd[3 + q] * s1;
            v2[q] = st.Vd[q] * c2 + st.Vd[3 + q] * s2;
        }
        const bool first = v1[0] > v2[0];
        const double* h = first ? v1 : v2;
        const double* e = first ? v2 : v1;
        const double nh = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), ne = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        bool ok = true;
        for (int q = 0; q < 3; ++q) {
            sh->Msrc[q] = h[q] / nh; sh->Msrc[3 + q] = e[q] / ne;
            ok = ok && isfinite(sh->Msrc[q]) && isfinite(sh->Msrc[3 + q]);
        }
        if (!ok) { sh->flags = SB_STATUS_DEGENERATE; tile_flagged(k, tile, SB_STATUS_DEGENERATE); }
        else {
            for (int q = 0; q < 6; ++q) { st.Msrc[q] = sh->Msrc[q]; if (a.M_out) a.M_out[(size_t)tile * 6 + q] = sh->Msrc[q]; }
            if (a.mode < PIPE_FIT) { k.cconsts[tile].mode = 1; if (a.status) a.status[tile] = 0; }
            else {
                make_lasso_consts(sh->Msrc, a.lasso_lambda, sh->lk);
                const int lm = lasso_mode_of(sh->lk.rg00, sh->lk.rg11, sh->lk.g01);
                if (lm == LASSO_GENERAL) { tile_to_fallback(k, tile, 6); sh->flags = -1; }      // non-unit rows: robust path
            }
        }
        sh->s_cnt = 0; sh->s_ok = 0;
    }
    __syncthreads();
    if (sh->flags != 0 || a.mode < PIPE_FIT) return;
    plan_conc_block(k, tile, ts, sh);
}

// 6: exact 99th percentiles of the two concentrations (normalizer.py:36,47) from the bracket lists.
__global__ void __launch_bounds__(NT) select_conc_kernel(TileKernelArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared* ts = reinterpret_cast<TileShared*>(smem_raw);
    PipeShared* sh = &ts->ps;
    const PipeArgs& a = k.a;
    const int tile = blockIdx.x;
    TileState& st = k.state[tile];
    if (st.path != PATH_STREAM) return;
    const int npx = a.npx;
    unsigned c_lo, c_hi;
    double fr;
    percentile_index((unsigned)npx, a.conc_pct, c_lo, c_hi, fr);
    const unsigned ka0 = st.cbrk[0], kb0 = st.cbrk[1], ka1 = st.cbrk[2], kb1 = st.cbrk[3];
    const TileLists tl = tile_lists(smem_raw, sh, k.list_cap);
    const KeyList list0{tl.l0, ka0, kb0 - ka0 > LIST_SPAN, tl.bytes}, list1{tl.l1, ka1, kb1 - ka1 > LIST_SPAN, tl.bytes};
    if (threadIdx.x == 0) {
        bool ok = true;
        for (int j = 0; j < 2; ++j) {
            sh->l_len[j] = st.clen[j]; sh->l_below[j] = st.cbelow[j];
            ok = ok && st.clen[j] <= min((j ? list1 : list0).cap(), (unsigned)k.list_cap) && st.cbelow[j] <= c_lo && c_hi < st.cbelow[j] + st.clen[j];
        }
        sh->s_ok = ok ? 1 : 0;
        if (!ok) tile_to_fallback(k, tile, 7);
    }
    __syncthreads();
    if (!sh->s_ok) return;
    const unsigned* gl = k.lists + (size_t)tile * 2 * k.list_cap;
    load_list(list0, gl, sh->l_len[0]);
    load_list(list1, gl + k.list_cap, sh->l_len[1]);
    __syncthreads();
    {
        const unsigned r_lo[2] = {c_lo - sh->l_below[0], c_lo - sh->l_below[1]};
        const unsigned r_hi[2] = {c_hi - sh->l_below[0], c_hi - sh->l_below[1]};
        unsigned keys[4];
        list_select_pairs(sh, list0, list1, r_lo, r_hi, keys);
        if (threadIdx.x < 4) sh->okey[threadIdx.x] = keys[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cv[4];
        for (int q = 0; q < 4; ++q) cv[q] = conc_from_key(sh->okey[q]);
        if (a.maxC_out) {
            a.maxC_out[(size_t)tile * 2] = lerp_np(cv[0], cv[1], fr);
            a.maxC_out[(size_t)tile * 2 + 1] = lerp_np(cv[2], cv[3], fr);
        }
        if (a.status) a.status[tile] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ Vahadane per-tile kernels
// One block-coordinate sweep of the dictionary update (Mairal et al. 2010, Alg. 2) from the sums of a pass -- the
// thread-0 code of the fused kernel, operation for operation.  t = (A00, A01, A11, B[:,0] (3), B[:,1] (3)); D rows = atoms.
__device__ inline void dl_sweep(const double* t, const double* D, double* FD) {
    for (int q = 0; q < 6; ++q) FD[q] = D[q];
    const double Aj[2][2] = {{t[0], t[1]}, {t[1], t[2]}};
    for (int j = 0; j < 2; ++j) {
        if (Aj[j][j] > 1e-12) {
            double u[3], nrm = 0.0;
            for (int q = 0; q < 3; ++q) {
                const double Da = FD[q] * Aj[0][j] + FD[3 + q] * Aj[1][j];
                u[q] = (t[3 + 3 * j + q] - Da) / Aj[j][j] + FD[3 * j + q];
                u[q] = u[q] > 0.0 ? u[q] : 0.0;
                nrm += u[q] * u[q];
            }
            nrm = sqrt(nrm);
            const double sc = 1.0 / (nrm > 1.0 ? nrm : 1.0);
            for (int q = 0; q < 3; ++q) FD[3 * j + q] = u[q] * sc;
        }
    }
}
// vahadane_stain_extractor.py:38-43: H first, rows normalised.  Returns false for a non-finite matrix.
__device__ inline bool dl_finish_matrix(const double* D, double* M) {
    const bool swap = D[0] < D[3];
    const double* h = swap ? D + 3 : D;
    const double* e = swap ? D : D + 3;
    const double nh = sqrt(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), ne = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    bool ok = true;
    for (int q = 0; q < 3; ++q) { M[q] = h[q] / nh; M[3 + q] = e[q] / ne; ok = ok && isfinite(M[q]) && isfinite(M[3 + q]); }
    return ok;
}
struct DictKernelArgs {
    TileKernelArgs k;
    DictState* dstate;
    DictConsts* dconsts;
    const unsigned short* mask;
};
// Thread 0: the dictionary iteration of a tile has ended -> stain matrix, outputs; the tile leaves the dictionary passes.
__device__ inline void dl_tile_done(const DictKernelArgs& d, int tile, const double* D) {
    const TileKernelArgs& k = d.k;
    d.dconsts[tile].mode = 1;
    double M[6];
    if (!dl_finish_matrix(D, M)) { tile_flagged(k, tile, SB_STATUS_DEGENERATE); return; }
    for (int q = 0; q < 6; ++q) { k.state[tile].Msrc[q] = M[q]; if (k.a.M_out) k.a.M_out[(size_t)tile * 6 + q] = M[q]; }
    if (k.a.mode < PIPE_FIT && k.a.status) k.a.status[tile] = 0;
}

// Warm start of the dictionary on the 1-in-16 sample (phase 0 of the fused kernel): the sample passes of a tile need no
// other tile, so ONE launch runs them all, tile per CTA, to the residual DL_SAMPLE_TOL.  Most of a tile's time here is
// the serial fp64 dictionary update + Anderson step between passes, so the CTAs are small (128 threads, eight per SM):
// what overlaps one tile's serial step is the other tiles' passes.  The partial sums are those of the fused kernel bit
// for bit: fp32 per lane over one warp step (32 consecutive sample blocks), shuffle tree, fixed point.
constexpr int DLS_NT = 128;
struct __align__(16) DlSampleShared {
    float2 tab[256];
    long long red[DLS_NT / 32][10];
    double tot[10];
    double D[6];
    LassoK lk;
    AAState aa;
    int dl_stop, flags;
};
__global__ void __launch_bounds__(DLS_NT, 8) dl_sample_kernel(DictKernelArgs d) {
    __shared__ DlSampleShared shm;
    DlSampleShared* sh = &shm;
    const TileKernelArgs& k = d.k;
    const PipeArgs& a = k.a;
    const int tile = blockIdx.x;
    const int npx = a.npx, G = npx / GROUP_PX;
    const uint8_t* __restrict__ tin = a.in + (size_t)tile * npx * 3;
    TileState& st = k.state[tile];
    DictState& ds = d.dstate[tile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += DLS_NT) sh->tab[i] = make_float2(a.tab.od[i], (float)a.tab.gamma[i]);
    if (threadIdx.x == 0) {
        const unsigned n_tissue = (unsigned)st.mom[9];
        st.n_tissue = n_tissue;
        sh->flags = 0;
        d.dconsts[tile].mode = 1;
        k.aconsts[tile].mode = 1;
        k.cconsts[tile].mode = 1;
        if (n_tissue < 1u) { sh->flags = SB_STATUS_EMPTY_MASK; tile_flagged(k, tile, SB_STATUS_EMPTY_MASK); }
        // a sample too thin for the warm start (< 1024 tissue pixels): the tile's 4 extra full passes are not worth four
        // launches for the whole batch -- the fused kernel runs it
        else if (a.dl_sample_iters > 0 && !((double)st.mom[0] >= 1024.0)) { sh->flags = -1; tile_to_fallback(k, tile, 1); }
        const double r0[3] = {0.65, 0.70, 0.29}, r1[3] = {0.07, 0.99, 0.11};
        const double n0 = sqrt(r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2]), n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
        for (int q = 0; q < 3; ++q) { sh->D[q] = r0[q] / n0; sh->D[3 + q] = r1[q] / n1; }
        make_dict_lasso_consts(sh->D, a.dl_lambda, sh->lk);
        aa_reset(sh->aa);
        sh->dl_stop = 0;
    }
    __syncthreads();
    if (sh->flags != 0) return;
    const bool use_sample = a.dl_sample_iters > 0 && (double)st.mom[0] >= 1024.0;
    const unsigned short* mrow = d.mask + (size_t)tile * G;
    const uint4* tile_sample = k.sample + (size_t)tile * (size_t)(G / SAMPLE_STRIDE) * 3;
    const int n_blk = G / SAMPLE_STRIDE;                 // sample blocks of the tile (npx is a multiple of 16)
    if (use_sample) {
        for (int it = 0; it < a.dl_sample_iters; ++it) {
            const LassoK lk = sh->lk;
            const int lm = lk.g01 >= 0.f ? LASSO_UNIT_POS : LASSO_UNIT_NEG;
            if (lane < 10) sh->red[warp][lane] = 0;
            __syncwarp();
            // warp step: 32 consecutive sample blocks, one per lane (the same 32 blocks, on the same lanes, as in the fused kernel)
            for (int j0 = (int)(threadIdx.x & ~31u); j0 < n_blk; j0 += DLS_NT) {
                const int j = j0 + lane;
                float2 f[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) f[i] = make_float2(0.f, 0.f);
                if (j < n_blk) {
                    const int g = sample_group_of_block(j);
                    uint32_t w[12];
                    load_sample_block(tile_sample, j, w);
                    const uint32_t mbits = mrow[g];
                    const float2* tb = sh->tab;
                    {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {             // same pixel pairing as for_each_pair_od
                            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
                            const float2 oa[3] = {f2(tb[wa & 255u].x, tb[wa >> 24].x), f2(tb[(wa >> 8) & 255u].x, tb[wb & 255u].x),
                                                  f2(tb[(wa >> 16) & 255u].x, tb[(wb >> 8) & 255u].x)};
                            const float2 ob[3] = {f2(tb[(wb >> 16) & 255u].x, tb[(wc >> 8) & 255u].x), f2(tb[wb >> 24].x, tb[(wc >> 16) & 255u].x),
                                                  f2(tb[wc & 255u].x, tb[wc >> 24].x)};
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float2 o0 = h ? ob[0] : oa[0], o1 = h ? ob[1] : oa[1], o2 = h ? ob[2] : oa[2];
                                const int i = 4 * q + 2 * h;
                                float2 c0, c1;
                                if (lm == LASSO_UNIT_POS) lasso2_unit_pair<LASSO_UNIT_POS>(lk, o0, o1, o2, c0, c1);
                                else lasso2_unit_pair<LASSO_UNIT_NEG>(lk, o0, o1, o2, c0, c1);
                                const bool ma = (mbits & (1u << i)) != 0, mb = (mbits & (2u << i)) != 0;
                                c0.x = ma ? c0.x : 0.f; c1.x = ma ? c1.x : 0.f;
                                c0.y = mb ? c0.y : 0.f; c1.y = mb ? c1.y : 0.f;
                                f[0] = __ffma2_rn(c0, c0, f[0]); f[1] = __ffma2_rn(c0, c1, f[1]); f[2] = __ffma2_rn(c1, c1, f[2]);
                                f[3] = __ffma2_rn(o0, c0, f[3]); f[4] = __ffma2_rn(o1, c0, f[4]); f[5] = __ffma2_rn(o2, c0, f[5]);
                                f[6] = __ffma2_rn(o0, c1, f[6]); f[7] = __ffma2_rn(o1, c1, f[7]); f[8] = __ffma2_rn(o2, c1, f[8]);
                            }
                        }
                    }
                }
                // flush of the warp step: shuffle tree, fixed point, the warp's own accumulator row (no atomics: integer
                // addition is associative, so this equals the fused kernel's one atomic per warp step)
#pragma unroll
                for (int i = 0; i < 9; ++i) {
                    float v = f[i].x + f[i].y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                    if (lane == 0) sh->red[warp][i] += to_fix(v, FIX_DL);
                }
            }
            __syncthreads();
            if (threadIdx.x < 9) {
                long long t = 0;
                for (int wv = 0; wv < DLS_NT / 32; ++wv) t += sh->red[wv][threadIdx.x];
                sh->tot[threadIdx.x] = (double)t * (1.0 / (double)FIX_DL);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double FD[6];
                { double sc[2]; dict_scales(sh->D, sc); dict_scale_sums(sh->tot, sc); }
                dl_sweep(sh->tot, sh->D, FD);
                double rn2 = 0.0;
                for (int q = 0; q < 6; ++q) rn2 += (FD[q] - sh->D[q]) * (FD[q] - sh->D[q]);
                if (a.dl_anderson > 0 && rn2 < DL_SAMPLE_TOL * DL_SAMPLE_TOL) sh->dl_stop = 1;      // this step is still applied
                aa_step(sh->aa, a.dl_anderson, sh->D, FD);
                make_dict_lasso_consts(sh->D, a.dl_lambda, sh->lk);
            }
            __syncthreads();
            if (sh->dl_stop) break;
        }
    }
    if (threadIdx.x == 0) {
        // phase 1 starts: the full passes inherit the difference history of the sample passes (aa_carry), or start afresh
        if (use_sample) aa_carry(sh->aa); else aa_reset(sh->aa);
        for (int q = 0; q < 6; ++q) ds.D[q] = sh->D[q];
        for (int q = 0; q < 10; ++q) ds.sums[q] = 0ull;
        ds.aa = sh->aa;
        ds.it = 0;
        ds.n_it = a.dl_iters;
        if (ds.n_it <= 0) dl_tile_done(d, tile, sh->D);
        else {
            DictConsts c;
            c.lk = sh->lk; c.lm = c.lk.g01 >= 0.f ? LASSO_UNIT_POS : LASSO_UNIT_NEG; c.mode = 0; c.pad = 0;
            d.dconsts[tile] = c;
        }
    }
}

// After every full pass: dictionary update + Anderson step of every tile still iterating.  One warp per tile: the lanes
// move the tile's Anderson state between global and shared memory (coalesced), lane 0 runs the serial fp64 step on it.
constexpr int DL_UPD_THREADS = 32;
__global__ void __launch_bounds__(DL_UPD_THREADS) dl_update_kernel(DictKernelArgs d) {
    __shared__ AAState aa;
    static_assert(sizeof(AAState) % 8 == 0, "AAState is copied as 8-byte words");
    const int tile = blockIdx.x;
    const PipeArgs& a = d.k.a;
    if (d.dconsts[tile].mode != 0) return;
    DictState& ds = d.dstate[tile];
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&ds.aa);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&aa);
        for (int i = threadIdx.x; i < (int)(sizeof(AAState) / 8); i += DL_UPD_THREADS) dst[i] = src[i];
    }
    __syncwarp();
    int done = 0;
    if (threadIdx.x == 0) {
        double t[9], D[6], FD[6];
        for (int q = 0; q < 9; ++q) { t[q] = (double)(long long)ds.sums[q] * (1.0 / (double)FIX_DL); ds.sums[q] = 0ull; }
        for (int q = 0; q < 6; ++q) D[q] = ds.D[q];
        { double sc[2]; dict_scales(D, sc); dict_scale_sums(t, sc); }
        dl_sweep(t, D, FD);
        double rn2 = 0.0;
        for (int q = 0; q < 6; ++q) rn2 += (FD[q] - D[q]) * (FD[q] - D[q]);
        const bool stop = a.dl_anderson > 0 && rn2 < DL_FULL_TOL * DL_FULL_TOL;      // this step is still applied, then the iteration ends
        aa_step(aa, a.dl_anderson, D, FD);
        ds.it += 1;
        for (int q = 0; q < 6; ++q) ds.D[q] = D[q];
        if (stop || ds.it >= ds.n_it) { dl_tile_done(d, tile, D); done = 1; }
        else {
            DictConsts c;
            make_dict_lasso_consts(D, a.dl_lambda, c.lk);
            c.lm = c.lk.g01 >= 0.f ? LASSO_UNIT_POS : LASSO_UNIT_NEG; c.mode = 0; c.pad = 0;
            d.dconsts[tile] = c;
        }
    }
    __syncwarp();
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done) return;
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&aa);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&ds.aa);
        for (int i = threadIdx.x; i < (int)(sizeof(AAState) / 8); i += DL_UPD_THREADS) dst[i] = src[i];
    }
}

// Vahadane, fit / transform: brackets of the concentration pass from the tile's stain matrix (C0).
__global__ void __launch_bounds__(NT, 3) vahadane_plan_conc_kernel(TileKernelArgs k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared* ts = reinterpret_cast<TileShared*>(smem_raw);
    PipeShared* sh = &ts->ps;
    const int tile = blockIdx.x;
    TileState& st = k.state[tile];
    if (st.path != PATH_STREAM) return;
    fill_small_table(ts->tab, k.a.tab);
    if (threadIdx.x == 0) {
        sh->flags = 0;
        for (int q = 0; q < 6; ++q) sh->Msrc[q] = st.Msrc[q];
        make_lasso_consts(sh->Msrc, k.a.lasso_lambda, sh->lk);
        if (lasso_mode_of(sh->lk.rg00, sh->lk.rg11, sh->lk.g01) == LASSO_GENERAL) { tile_to_fallback(k, tile, 6); sh->flags = -1; }
        sh->s_cnt = 0; sh->s_ok = 0;
    }
    __syncthreads();
    if (sh->flags != 0) return;
    plan_conc_block(k, tile, ts, sh);
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int STREAM_SUB_BATCH = 4096;              // tiles per round of passes: bounds the key-list scratch at 256 MB

// status must have been preset to SB_STATUS_EMPTY_MASK on the stream (launch_mask does it).
int launch_mask_stream(const PointArgs& a, int num_sms, cudaStream_t stream) {
    MaskOutParams p{};
    p.gamma = a.tab.gamma; p.ycoef[0] = a.ycoef[0]; p.ycoef[1] = a.ycoef[1]; p.ycoef[2] = a.ycoef[2]; p.ybound = a.ybound;
    p.mask_out = a.mask_out; p.status = a.status; p.npx = a.npx;
    return launch_ring_reduce<MaskOutOp>(RingGeom{a.in, nullptr, a.B, a.npx}, p, num_sms, stream);
}

bool concentrations_stream_eligible(const PointArgs& a) { return a.aligned && a.npx % GROUP_PX == 0 && a.npx >= 4096; }
int launch_concentrations_stream(const PointArgs& a, int num_sms, cudaStream_t stream) {
    ConcOutParams p{};
    p.od = a.tab.od; p.M = a.M; p.lasso_lambda = a.lasso_lambda; p.conc_out = a.conc_out; p.npx = a.npx;
    return launch_ring_reduce<ConcOutOp>(RingGeom{a.in, nullptr, a.B, a.npx}, p, num_sms, stream);
}

int stream_fallback_counters(unsigned out[8], bool reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_fallback_reasons, 8 * sizeof(unsigned));
    if (e == cudaSuccess && reset) {
        const unsigned zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_fallback_reasons, zero, sizeof(zero));
    }
    return (int)e;
}

bool stream_pipeline_eligible(const PipeArgs& a) {
    return (a.method == SB_METHOD_MACENKO || a.method == SB_METHOD_VAHADANE) && a.aligned && a.npx >= 32768 && (a.npx % GROUP_PX) == 0 &&
           a.tile_list == nullptr;
}
static size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }
// Tiles per round of passes: at most STREAM_SUB_BATCH, and at most 2^30 pixels (128 MB of tissue bits).
static int stream_sub_batch(int B, int npx) {
    long long n = (1LL << 30) / (npx > 0 ? npx : 1);
    if (n > STREAM_SUB_BATCH) n = STREAM_SUB_BATCH;
    if (n < 1) n = 1;
    // (beyond a megapixel the lists double: at most 2^30 / npx < 1024 tiles of them, 128 MB)
    return (int)(B < n ? B : n);
}
size_t stream_scratch_bytes(int B, int npx) {
    const size_t n = (size_t)stream_sub_batch(B, npx);
    return up256(n * sizeof(TileState)) + up256(n * sizeof(AngleConsts)) + up256(n * sizeof(ConcConsts)) +
           up256(n * 2 * (size_t)slist_cap(npx) * sizeof(unsigned)) + up256((n + 1) * sizeof(int)) + up256(n * (size_t)(npx / GROUP_PX) * sizeof(unsigned short)) +
           up256(n * sizeof(DictState)) + up256(n * sizeof(DictConsts)) + up256(n * (size_t)(npx / GROUP_PX / SAMPLE_STRIDE) * 48);
}

int launch_stream_pipeline(const PipeArgs& a_all, Scratch& scratch) {
    cudaStream_t st = scratch.st;
    PassTimer pt{scratch.h, st};
    scratch.h->n_pass_ev = 0;
    const int num_sms = scratch.h->num_sms;
    const int nsub = stream_sub_batch(a_all.B, a_all.npx);
    TileState* state = nullptr; AngleConsts* ac = nullptr; ConcConsts* cc = nullptr; unsigned* lists = nullptr; int* fb = nullptr;
    unsigned short* mask = nullptr;
    cudaError_t e;
    if ((e = scratch.get(&state, (size_t)nsub * sizeof(TileState))) != cudaSuccess) return (int)e;
    if ((e = scratch.get(&ac, (size_t)nsub * sizeof(AngleConsts))) != cudaSuccess) return (int)e;
    if ((e = scratch.get(&cc, (size_t)nsub * sizeof(ConcConsts))) != cudaSuccess) return (int)e;
    const int list_cap = slist_cap(a_all.npx);
    if ((e = scratch.get(&lists, (size_t)nsub * 2 * list_cap * sizeof(unsigned))) != cudaSuccess) return (int)e;
    if ((e = scratch.get(&fb, (size_t)(nsub + 1) * sizeof(int))) != cudaSuccess) return (int)e;
    if ((e = scratch.get(&mask, (size_t)nsub * (size_t)(a_all.npx / GROUP_PX) * sizeof(unsigned short))) != cudaSuccess) return (int)e;
    uint4* sample = nullptr;
    if ((e = scratch.get(&sample, (size_t)nsub * (size_t)(a_all.npx / GROUP_PX / SAMPLE_STRIDE) * 48)) != cudaSuccess) return (int)e;
    const bool vahadane = a_all.method == SB_METHOD_VAHADANE;
    DictState* dstate = nullptr; DictConsts* dconsts = nullptr;
    if (vahadane) {
        if ((e = scratch.get(&dstate, (size_t)nsub * sizeof(DictState))) != cudaSuccess) return (int)e;
        if ((e = scratch.get(&dconsts, (size_t)nsub * sizeof(DictConsts))) != cudaSuccess) return (int)e;
    }
    static DeviceOnce once_v;
    if ((e = ensure_dyn_smem(once_v, vahadane_plan_conc_kernel, (int)tile_shared_bytes(SLIST_CAP_MAX))) != cudaSuccess) return (int)e;
    static DeviceOnce once_p, once_a, once_c;
    const int tsm = (int)tile_shared_bytes(SLIST_CAP_MAX), tsm_run = (int)tile_shared_bytes(list_cap);
    if ((e = ensure_dyn_smem(once_p, plan_angle_kernel, tsm)) != cudaSuccess) return (int)e;
    static DeviceOnce once_a4;
    if ((e = ensure_dyn_smem(once_a, select_angle_kernel<3>, tsm)) != cudaSuccess) return (int)e;
    if ((e = ensure_dyn_smem(once_a4, select_angle_kernel<4>, tsm)) != cudaSuccess) return (int)e;
    if ((e = ensure_dyn_smem(once_c, select_conc_kernel, tsm)) != cudaSuccess) return (int)e;
    const size_t tile_bytes = (size_t)a_all.npx * 3;
    for (int t0 = 0; t0 < a_all.B; t0 += nsub) {
        PipeArgs a = a_all;
        a.B = (a_all.B - t0 < nsub) ? a_all.B - t0 : nsub;
        a.in = a_all.in + (size_t)t0 * tile_bytes;
        if (a.M_out) a.M_out += (size_t)t0 * 6;
        if (a.maxC_out) a.maxC_out += (size_t)t0 * 2;
        if (a.status) a.status += t0;
        if ((e = cudaMemsetAsync(state, 0, (size_t)a.B * sizeof(TileState), st)) != cudaSuccess) return (int)e;
        if ((e = cudaMemsetAsync(fb + nsub, 0, sizeof(int), st)) != cudaSuccess) return (int)e;
        StreamParams p{};
        p.state = state; p.aconsts = ac; p.cconsts = cc; p.lists = lists; p.od = a.tab.od; p.gamma = a.tab.gamma;
        p.mask = mask; p.groups = a.npx / GROUP_PX; p.list_cap = list_cap; p.sample = sample;
        p.ycoef[0] = a.ycoef[0]; p.ycoef[1] = a.ycoef[1]; p.ycoef[2] = a.ycoef[2]; p.ybound = a.ybound;
        const RingGeom g{a.in, nullptr, a.B, a.npx};
        TileKernelArgs k{a, state, ac, cc, lists, list_cap, fb, fb + nsub, sample};
        int rc;
        int n_launch = 0;
        if (vahadane) {
            p.dstate = dstate; p.dconsts = dconsts;
            const DictKernelArgs d{k, dstate, dconsts, mask};
            const int unit_chunks = unit_groups(a.npx / GROUP_PX);       // chunk = NT groups: a unit of for_each_unit is this many chunks
            const int n_full = a.dl_iters;       // (tiles without a usable sample, which get 4 more, are fallback tiles)
            pt.mark("ring_reduce<MaskOp>: tissue mask");
            { NvtxRange r("stream: tissue mask"); if ((rc = launch_ring_reduce<MaskOp>(g, p, num_sms, st)) != 0) return rc; }
            pt.mark("dl_sample_kernel: dictionary warm start on the 1-in-16 sample");
            { NvtxRange r("stream: dictionary, sample passes"); dl_sample_kernel<<<a.B, DLS_NT, 0, st>>>(d); }
            for (int it = 0; it < n_full; ++it) {
                NvtxRange r("stream: dictionary, full pass");
                pt.mark("ring_reduce<DictOp>: full dictionary pass");
                if ((rc = launch_ring_reduce<DictOp>(g, p, num_sms, st, unit_chunks)) != 0) return rc;
                pt.mark("dl_update_kernel");
                dl_update_kernel<<<a.B, DL_UPD_THREADS, 0, st>>>(d);
            }
            n_launch = 2 + 2 * n_full;
            if (a.mode >= PIPE_FIT) pt.mark("vahadane_plan_conc_kernel");
            if (a.mode >= PIPE_FIT) { NvtxRange r("stream: plan concentration"); vahadane_plan_conc_kernel<<<a.B, NT, TILE_LIST_OFFSET, st>>>(k); ++n_launch; }
        } else {
            pt.mark("ring_reduce<MomentOp>: mask + moments");
            { NvtxRange r("stream: moments"); if ((rc = launch_ring_reduce<MomentOp>(g, p, num_sms, st)) != 0) return rc; }
            pt.mark("plan_angle_kernel");
            { NvtxRange r("stream: plan angle"); plan_angle_kernel<<<a.B, NT, TILE_LIST_OFFSET, st>>>(k); }
            pt.mark("ring_reduce<AngleOp>: angle brackets");
            { NvtxRange r("stream: angle brackets"); if ((rc = launch_ring_reduce<AngleOp>(g, p, num_sms, st)) != 0) return rc; }
            pt.mark("select_angle_kernel");
            {
                NvtxRange r("stream: select angle");
                if (list_cap <= SLIST_CAP_MAX / 4) select_angle_kernel<4><<<a.B, NT, tsm_run, st>>>(k);
                else select_angle_kernel<3><<<a.B, NT, tsm_run, st>>>(k);
            }
            n_launch = 4;
        }
        if (a.mode >= PIPE_FIT) {
            pt.mark("ring_reduce<ConcOp>: concentration brackets");
            { NvtxRange r("stream: concentration brackets"); if ((rc = launch_ring_reduce<ConcOp>(g, p, num_sms, st)) != 0) return rc; }
            pt.mark("select_conc_kernel");
            { NvtxRange r("stream: select concentration"); select_conc_kernel<<<a.B, NT, tsm_run, st>>>(k); }
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        // tiles the streaming passes did not serve: the fused kernel, driven by the device-side list (no host sync)
        PipeArgs f = a;
        f.cluster_size = 1;
        f.tile_list = fb; f.tile_count = fb + nsub;
        pt.mark("tile_pipeline_kernel: fallback tiles");
        { NvtxRange r("stream: fused fallback"); if ((rc = launch_tile_pipeline(f, num_sms, st)) != 0) return rc; }
        pt.mark("end");
        scratch.h->launches += n_launch + (a.mode >= PIPE_FIT ? 3 : 1);
    }
    return 0;
}

}  // namespace sb
