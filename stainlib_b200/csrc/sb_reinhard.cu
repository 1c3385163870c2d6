// sb_reinhard.cu -- ReinhardStainNormalizer.fit / .transform (normalizer.py:64-94, stain_utils.py:146-194) and
// LuminosityStandardizer.standardize (stain_utils.py:53-67) as STREAMING passes on the TMA ring, one launch per dependency
// level over the whole batch (the per-tile kernel lab_tile_kernel of sb_colour.cu keeps the unaligned tiles):
//
//   R1  rein_ring_kernel<ByteHistOp>     histogram of all 3N channel bytes of every tile                      (3 B/px read)
//   P   rein_plan_kernel                 per tile: 90th percentile -> brightness table folded into the sRGB linearisation
//   R2  rein_ring_kernel<LabStatsOp>     integer sRGB -> CIELAB of the standardised pixels; exact L histogram, integer sums
//                                        of a, a^2, b, b^2; transform: the LAB bytes are parked in the output tile (3 + 3 B/px)
//   S   rein_stats_kernel                per tile: means / stds exactly as cv.meanStdDev sees them; transform: the three
//                                        affine maps + merge_back folded into two 256-entry tables of the inverse path
//   R3  rein_ring_kernel<LabInvOp>       LAB bytes -> mapped -> integer CIELAB -> sRGB, in place in the output tile   (3 + 3 B/px)
//
// Transport as in sb_ring.cuh (persistent CTAs, one per SM; a producer warp streams 24 KB chunks HBM -> shared memory with
// cp.async.bulk on mbarriers; 16 compute warps take one 48-byte group per thread per chunk) with two additions: per-TILE
// lookup tables (the 64 KB lane-replicated table is refilled when the CTA moves to the next tile, behind a named barrier of
// the compute warps) and CTA-shared, lane-replicated histograms (bank = lane: no conflicts) flushed per tile.
//
// Arithmetic.  The forward conversion is OpenCV's fixed-point path (oracle/cv_lab.py) evaluated EXACTLY in fp32: the
// linearised channels are integers <= 2040, so R*1777 + G*1541 + B*778 + 2048 and its two siblings stay below 2^23 and every
// fused multiply-add is exact; ">> 12" is a scaling by 2^-12 (exact) and a round-down add of 2^23, which leaves the table
// index in the mantissa; the L, a, b formulas are again integers below 2^24 in units of 2^-15.  Two pixels share each
// instruction on the packed f32x2 pipe.  The inverse conversion keeps OpenCV's integer arithmetic (its products exceed
// 2^24).  Every floating-point step of the reference is a function of one uint8 and lives in the per-tile tables, computed
// in fp64 by the per-tile kernels with the formulas of lab_tile_kernel -- both paths give the same bytes.
#include "sb_kernels.h"
#include "sb_ring.cuh"
#include "sb_lab.cuh"

namespace sb {

constexpr int RN_GT = 512;                      // compute threads: one 16-pixel group each per chunk
constexpr int RN_CHUNK = RN_GT * 48;
constexpr int RN_BAR_BYTES = 256;

__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RN_GT) : "memory"); }
__device__ __forceinline__ void red_shared_inc(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t r; asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr)); return r; }
__device__ __forceinline__ uint32_t lds_u32_off4(uint32_t addr) { uint32_t r; asm("ld.shared.u32 %0, [%1+4];" : "=r"(r) : "r"(addr)); return r; }
__device__ __forceinline__ float lds_f32(uint32_t addr) { float r; asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr)); return r; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t r; asm("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr)); return r; }
// table row address of byte k of w: T | byte << 8 | lane_base's low byte (one PRMT, as od_lookup)
__device__ __forceinline__ uint32_t row_addr(const OdAbs& t, uint32_t w, int k) { return __byte_perm(w, t.lane_base, 0x6504u | (k << 4)); }

// An Op provides
//   Params, Acc / acc_init, kStore (the chunk is written back: in-place transform + bulk store), kStages, kExtraBytes
//   (small CTA tables in front of the ring), kLaneShift
//   init_static(p, extra, tab, tid, n)   all n threads of the CTA, once
//   tile_begin(p, tile, tab, tid)        the RN_GT compute threads: fill the per-tile table
//   process(p, tab, extra_addr, grp, acc) one 48-byte group in shared memory (in place when kStore)
//   run_flush(p, tile, acc)              every compute thread when the CTA leaves a tile: register accumulators -> global
//   tile_end(p, tile, tab, tid)          the compute threads, behind a barrier: CTA-shared histograms -> global, zeroed
template <class Op>
__global__ void __launch_bounds__(RN_GT + 32, 1) rein_ring_kernel(RingGeom g, typename Op::Params p, int chunks_per_tile, long long total_chunks) {
    constexpr int NST = Op::kStages;
    constexpr int HEAD = RN_BAR_BYTES + Op::kExtraBytes;
    static_assert(HEAD % 128 == 0, "stage alignment");
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t tab_addr = (base + HEAD + 0xFFFFu) & ~0xFFFFu;
    unsigned char* tab_ptr = smem + (tab_addr - base);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + NST;
    unsigned char* extra = smem + RN_BAR_BYTES;
    const int n_front = (int)((tab_addr - base - HEAD) / RN_CHUNK);
    const int n_back = ((int)RING_SMEM_BYTES - (int)(tab_addr - base) - OD_REP_BYTES) / RN_CHUNK;
    if (n_front + n_back < NST) __trap();
    auto stage_ptr = [&](int s) -> unsigned char* {
        return s < n_front ? smem + HEAD + (size_t)s * RN_CHUNK : tab_ptr + OD_REP_BYTES + (size_t)(s - n_front) * RN_CHUNK;
    };
    const size_t tile_bytes = (size_t)g.npx * 3;
    const long long c_begin = total_chunks * blockIdx.x / gridDim.x, c_end = total_chunks * (blockIdx.x + 1) / gridDim.x;
    const int n_local = (int)(c_end - c_begin);
    auto chunk_geom = [&](long long c, int& tile, size_t& off, uint32_t& bytes) {
        tile = (int)(c / chunks_per_tile);
        off = (size_t)(c % chunks_per_tile) * RN_CHUNK;
        const size_t rem = tile_bytes - off;
        bytes = (uint32_t)(rem < (size_t)RN_CHUNK ? rem : (size_t)RN_CHUNK);
    };
    if (threadIdx.x == RN_GT) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], RN_GT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    Op::init_static(p, extra, tab_ptr, (int)threadIdx.x, RN_GT + 32);
    __syncthreads();

    if (threadIdx.x >= RN_GT) {
        // ------------------------------------------------------------------ producer warp (one elected lane)
        if (threadIdx.x == RN_GT) {
            if (Op::kStore) {
                for (int i = 0; i < NST && i < n_local; ++i) {
                    int tile; size_t off; uint32_t bytes;
                    chunk_geom(c_begin + i, tile, off, bytes);
                    mbar_expect_tx(&full[i], bytes);
                    bulk_load(stage_ptr(i), g.in + (size_t)tile * tile_bytes + off, bytes, &full[i]);
                }
                for (int i = 0; i < n_local; ++i) {
                    const int s = i % NST;
                    int tile; size_t off; uint32_t bytes;
                    chunk_geom(c_begin + i, tile, off, bytes);
                    mbar_wait_relaxed(&done[s], (uint32_t)((i / NST) & 1));       // stage s holds the finished output of chunk i
                    bulk_store(g.out + (size_t)tile * tile_bytes + off, stage_ptr(s), bytes);
                    if (i >= 1 && i - 1 + NST < n_local) {                          // refill the stage of chunk i-1 once its store has read it
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        const int ps = (i - 1) % NST;
                        int t2; size_t o2; uint32_t b2;
                        chunk_geom(c_begin + i - 1 + NST, t2, o2, b2);
                        mbar_expect_tx(&full[ps], b2);
                        bulk_load(stage_ptr(ps), g.in + (size_t)t2 * tile_bytes + o2, b2, &full[ps]);
                    }
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            } else {
                for (int n = 0; n < n_local; ++n) {
                    const int s = n % NST;
                    if (n >= NST) mbar_wait_relaxed(&done[s], (uint32_t)(((n / NST) - 1) & 1));   // the slot's previous chunk has been consumed
                    int tile; size_t off; uint32_t bytes;
                    chunk_geom(c_begin + n, tile, off, bytes);
                    mbar_expect_tx(&full[s], bytes);
                    bulk_load(stage_ptr(s), g.in + (size_t)tile * tile_bytes + off, bytes, &full[s]);
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- compute warps
    const OdAbs tab{((threadIdx.x & 31u) << Op::kLaneShift) | ((tab_addr >> 16) << 8)};
    const uint32_t extra_addr = base + RN_BAR_BYTES;
    typename Op::Acc acc;
    Op::acc_init(acc);
    int i = 0;
    while (i < n_local) {
        const int tile = (int)((c_begin + i) / chunks_per_tile);
        const int first_in_tile = (int)((c_begin + i) - (long long)tile * chunks_per_tile);
        int run = chunks_per_tile - first_in_tile;
        if (run > n_local - i) run = n_local - i;
        Op::tile_begin(p, tile, tab_ptr, (int)threadIdx.x);
        compute_bar_sync();                              // table of this tile filled, histograms of the previous tile zeroed
        for (int j = 0; j < run; ++j, ++i) {
            const int s = i % NST;
            const size_t off = (size_t)(first_in_tile + j) * RN_CHUNK;
            const size_t rem = tile_bytes - off;
            const uint32_t bytes = (uint32_t)(rem < (size_t)RN_CHUNK ? rem : (size_t)RN_CHUNK);
            unsigned char* buf = stage_ptr(s);
            mbar_wait(&full[s], (uint32_t)((i / NST) & 1));
            if (threadIdx.x * 48u < bytes) {
                Op::process(p, tab, extra_addr, reinterpret_cast<uint4*>(buf + threadIdx.x * 48u), acc);
                if (Op::kStore) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
        }
        Op::run_flush(p, tile, acc);
        compute_bar_sync();                              // every warp has left the tile: its table and histograms are free
        Op::tile_end(p, tile, tab_ptr, (int)threadIdx.x);
    }
}

template <class Op>
static int launch_rein_ring(const RingGeom& g, const typename Op::Params& p, int num_sms, cudaStream_t stream) {
    static_assert(OD_REP_BYTES + Op::kStages * RN_CHUNK + RN_BAR_BYTES + Op::kExtraBytes + 1024 <= RING_SMEM_BYTES, "ring does not fit");
    static_assert(2 * Op::kStages * 8 <= RN_BAR_BYTES, "barrier area");
    static DeviceOnce once;
    {
        cudaError_t e = ensure_dyn_smem(once, rein_ring_kernel<Op>, RING_SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
    }
    const size_t tile_bytes = (size_t)g.npx * 3;
    const int cpt = (int)((tile_bytes + RN_CHUNK - 1) / RN_CHUNK);
    const long long total = (long long)cpt * g.B;
    int grid = num_sms;
    if ((long long)grid > total) grid = (int)total;
    rein_ring_kernel<Op><<<grid, RN_GT + 32, RING_SMEM_BYTES, stream>>>(g, p, cpt, total);
    return (int)cudaGetLastError();
}

// Lane-replicated histogram rows (256-byte rows of the table area: counter of value v and lane l at v * 256 + row_off +
// l * 4) -> hist[256] in global memory, zeroed behind.  Thread t sums 16 lanes of bin t / 2.
__device__ __forceinline__ void flush_lane_hist(unsigned char* tab, int row_off, unsigned* ghist, int tid) {
    unsigned* row = reinterpret_cast<unsigned*>(tab + (tid >> 1) * OD_ROW_BYTES + row_off + (tid & 1) * 64);
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { s += row[k]; row[k] = 0u; }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if ((tid & 1) == 0 && s) atomicAdd(&ghist[tid >> 1], s);
}

// ------------------------------------------------------------------------------------------------ R1: byte histogram
struct ByteHistParams { unsigned* hist; };               // [B][256], zeroed
struct ByteHistOp {
    using Params = ByteHistParams;
    struct Acc {};
    static constexpr bool kStore = false;
    static constexpr int kStages = 6, kExtraBytes = 0, kLaneShift = 2;
    __device__ static void acc_init(Acc&) {}
    __device__ static void init_static(const Params&, unsigned char*, unsigned char* tab, int tid, int n) {
        for (int i = tid; i < OD_REP_BYTES / 16; i += n) reinterpret_cast<uint4*>(tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __device__ static void tile_begin(const Params&, int, unsigned char*, int) {}
    __device__ static void process(const Params&, const OdAbs tab, uint32_t, uint4* grp, Acc&) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            red_shared_inc(row_addr(tab, w[i], 0)); red_shared_inc(row_addr(tab, w[i], 1));
            red_shared_inc(row_addr(tab, w[i], 2)); red_shared_inc(row_addr(tab, w[i], 3));
        }
    }
    __device__ static void run_flush(const Params&, int, Acc&) {}
    __device__ static void tile_end(const Params& p, int tile, unsigned char* tab, int tid) { flush_lane_hist(tab, 0, p.hist + (size_t)tile * 256, tid); }
};

// ------------------------------------------------------------------------------------------------ R2: forward conversion + statistics
struct LabStatsParams {
    const float* gam2;             // [B][256] linearised channel of the standardised byte (per tile)
    const unsigned short* cbrt;    // [3072]
    unsigned* lhist;               // [B][256] zeroed: histogram of L
    unsigned long long* sums;      // [B][4]   zeroed: sum a, sum a^2, sum b, sum b^2 (bytes as stored, 0..255)
};
constexpr float RN_MAGIC = 8388608.f;                   // 2^23: x + 2^23 rounded down leaves floor(x) in the mantissa
template <bool KEEP>
struct LabStatsOp {
    using Params = LabStatsParams;
    struct Acc { unsigned sa, sa2, sb, sb2; };          // per thread and tile: <= 2^24 / 512 / 16 groups of <= 16 * 255^2 -- fits
    static constexpr bool kStore = KEEP;
    static constexpr int kStages = 5, kExtraBytes = 3072 * 4, kLaneShift = 2;
    __device__ static void acc_init(Acc& a) { a.sa = a.sa2 = a.sb = a.sb2 = 0u; }
    __device__ static void init_static(const Params& p, unsigned char* extra, unsigned char* tab, int tid, int n) {
        for (int i = tid; i < 3072; i += n) reinterpret_cast<float*>(extra)[i] = (float)p.cbrt[i];
        for (int i = tid; i < OD_REP_BYTES / 16; i += n) reinterpret_cast<uint4*>(tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // words 0..31 of a row: the tile's linearisation table, one copy per lane; words 32..63: the L histogram, one counter per lane
    __device__ static void tile_begin(const Params& p, int tile, unsigned char* tab, int tid) {
        const float* t = p.gam2 + (size_t)tile * 256;
        for (int i = tid; i < 256 * 32; i += RN_GT) *reinterpret_cast<float*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 4) = __ldg(t + (i >> 5));
    }
    // companding table lookup: the index sits in the mantissa of u = 2^23 + idx; address = cb + idx * 4 = bits * 4 + (cb - 0x2C000000)
    __device__ static __forceinline__ float cbrt_at(uint32_t kaddr, float u) { return lds_f32(__float_as_uint(u) * 4u + kaddr); }
    // Two pixels: linearised channels -> L, a, b as 2^23 + value (the byte is the low byte of the bits).  OpenCV:
    //   fX = cb[(R*1777 + G*1541 + B*778 + 2048) >> 12], fY = cb[(R*871 + G*2929 + B*296 + 2048) >> 12], fZ = cb[(R*73 + G*448 + B*3575 + 2048) >> 12]
    //   L = (296 fY - 1336934 + 2^14) >> 15, a = (500 (fX - fY) + 128 * 2^15 + 2^14) >> 15, b = (200 (fY - fZ) + 128 * 2^15 + 2^14) >> 15
    __device__ static __forceinline__ void lab_pair(uint32_t kaddr, const float2 R, const float2 G, const float2 B, float2& uL, float2& uA, float2& uB) {
        const float2 M = dup(RN_MAGIC);
        const float2 sx = __ffma2_rn(B, dup(778.f / 4096.f), __ffma2_rn(G, dup(1541.f / 4096.f), __ffma2_rn(R, dup(1777.f / 4096.f), dup(0.5f))));
        const float2 sy = __ffma2_rn(B, dup(296.f / 4096.f), __ffma2_rn(G, dup(2929.f / 4096.f), __ffma2_rn(R, dup(871.f / 4096.f), dup(0.5f))));
        const float2 sz = __ffma2_rn(B, dup(3575.f / 4096.f), __ffma2_rn(G, dup(448.f / 4096.f), __ffma2_rn(R, dup(73.f / 4096.f), dup(0.5f))));
        const float2 ux = __fadd2_rd(sx, M), uy = __fadd2_rd(sy, M), uz = __fadd2_rd(sz, M);
        const float2 fX = f2(cbrt_at(kaddr, ux.x), cbrt_at(kaddr, ux.y));
        const float2 fY = f2(cbrt_at(kaddr, uy.x), cbrt_at(kaddr, uy.y));
        const float2 fZ = f2(cbrt_at(kaddr, uz.x), cbrt_at(kaddr, uz.y));
        const float2 dxy = __fadd2_rn(fX, f2(-fY.x, -fY.y)), dyz = __fadd2_rn(fY, f2(-fZ.x, -fZ.y));
        uL = __fadd2_rd(__ffma2_rn(fY, dup(296.f / 32768.f), dup(-1320550.f / 32768.f)), M);
        uA = __fadd2_rd(__ffma2_rn(dxy, dup(500.f / 32768.f), dup(128.5f)), M);
        uB = __fadd2_rd(__ffma2_rn(dyz, dup(200.f / 32768.f), dup(128.5f)), M);
    }
    __device__ static void process(const Params&, const OdAbs tab, uint32_t extra_addr, uint4* grp, Acc& acc) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        const uint32_t kaddr = extra_addr - 0x2C000000u;
        const OdAbs htab{tab.lane_base | 0x80u};         // the histogram half of the rows
        float2 sa = dup(0.f), sa2 = dup(0.f), sb = dup(0.f), sb2 = dup(0.f);
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
            // pixels: p0=(a0,a1,a2) p1=(a3,b0,b1) p2=(b2,b3,c0) p3=(c1,c2,c3); pairs (p0,p1) and (p2,p3)
            float2 uL[2], uA[2], uB[2];
            lab_pair(kaddr, f2(od_lookup(tab, wa, 0u, 0), od_lookup(tab, wa, 0u, 3)), f2(od_lookup(tab, wa, 0u, 1), od_lookup(tab, wb, 0u, 0)),
                     f2(od_lookup(tab, wa, 0u, 2), od_lookup(tab, wb, 0u, 1)), uL[0], uA[0], uB[0]);
            lab_pair(kaddr, f2(od_lookup(tab, wb, 0u, 2), od_lookup(tab, wc, 0u, 1)), f2(od_lookup(tab, wb, 0u, 3), od_lookup(tab, wc, 0u, 2)),
                     f2(od_lookup(tab, wc, 0u, 0), od_lookup(tab, wc, 0u, 3)), uL[1], uA[1], uB[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 a = __fadd2_rn(uA[h], dup(-RN_MAGIC)), b = __fadd2_rn(uB[h], dup(-RN_MAGIC));      // exact integers 0..255
                sa = __fadd2_rn(sa, a); sa2 = __ffma2_rn(a, a, sa2);
                sb = __fadd2_rn(sb, b); sb2 = __ffma2_rn(b, b, sb2);
                red_shared_inc(row_addr(htab, __float_as_uint(uL[h].x), 0));
                red_shared_inc(row_addr(htab, __float_as_uint(uL[h].y), 0));
            }
            if (KEEP) {
                const uint32_t l0 = __float_as_uint(uL[0].x), a0 = __float_as_uint(uA[0].x), b0 = __float_as_uint(uB[0].x);
                const uint32_t l1 = __float_as_uint(uL[0].y), a1 = __float_as_uint(uA[0].y), b1 = __float_as_uint(uB[0].y);
                const uint32_t l2 = __float_as_uint(uL[1].x), a2 = __float_as_uint(uA[1].x), b2 = __float_as_uint(uB[1].x);
                const uint32_t l3 = __float_as_uint(uL[1].y), a3 = __float_as_uint(uA[1].y), b3 = __float_as_uint(uB[1].y);
                o[3 * q] = pack4(l0, a0, b0, l1);
                o[3 * q + 1] = pack4(a1, b1, l2, a2);
                o[3 * q + 2] = pack4(b2, l3, a3, b3);
            }
        }
        // the group's sums are exact integers (<= 16 * 255^2 < 2^24)
        acc.sa += (unsigned)(sa.x + sa.y); acc.sa2 += (unsigned)(sa2.x + sa2.y);
        acc.sb += (unsigned)(sb.x + sb.y); acc.sb2 += (unsigned)(sb2.x + sb2.y);
        if (KEEP) {
            grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
            grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
            grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
        }
    }
    __device__ static void run_flush(const Params& p, int tile, Acc& acc) {
        unsigned long long v[4] = {acc.sa, acc.sa2, acc.sb, acc.sb2};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
            if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(&p.sums[(size_t)tile * 4 + k], v[k]);
        }
        acc_init(acc);
    }
    __device__ static void tile_end(const Params& p, int tile, unsigned char* tab, int tid) { flush_lane_hist(tab, 128, p.lhist + (size_t)tile * 256, tid); }
};

// ------------------------------------------------------------------------------------------------ R3: map + inverse conversion
// Per tile and LAB byte value v, two words (rein_stats_kernel):
//   word 0 (looked up with L):  y | fy << 16 of the MAPPED lightness (lab2yf of OpenCV); bit 31 = background pixel (mask mode)
//   word 1 (looked up with a for its low half, with b for its high half): adiv of the mapped a | bdiv of the mapped b << 16
struct LabInvParams {
    const uint2* tl_ab;            // [B][256]
    const unsigned char* invg;     // [4096]
    int adiv_bg, bdiv_bg;          // a = b = 128 (background of mask mode)
};
template <bool MASK>
struct LabInvOp {
    using Params = LabInvParams;
    struct Acc {};
    static constexpr bool kStore = true;
    static constexpr int kStages = 5, kExtraBytes = 4096, kLaneShift = 3;
    __device__ static void acc_init(Acc&) {}
    __device__ static void init_static(const Params& p, unsigned char* extra, unsigned char*, int tid, int n) {
        for (int i = tid; i < 4096 / 4; i += n) reinterpret_cast<uint32_t*>(extra)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.invg) + i);
    }
    __device__ static void tile_begin(const Params& p, int tile, unsigned char* tab, int tid) {
        const uint2* t = p.tl_ab + (size_t)tile * 256;
        for (int i = tid; i < 256 * 32; i += RN_GT) *reinterpret_cast<uint2*>(tab + (i >> 5) * OD_ROW_BYTES + (i & 31) * 8) = __ldg(t + (i >> 5));
    }
    __device__ static __forceinline__ int cubic(int t) { return (int)((((unsigned)(t * t) >> 14) * (unsigned)t) >> 14); }
    // One pixel: the three table words -> sRGB bytes (OpenCV's Lab2RGBinteger; lab_inverse of sb_colour.cu).
    __device__ static __forceinline__ void px(const Params& p, uint32_t ig, uint32_t tl, uint32_t ta, uint32_t tb, uint32_t& r, uint32_t& g, uint32_t& b) {
        const int y = (int)(tl & 0xFFFFu);
        const int ify = MASK ? (int)((tl >> 16) & 0x7FFFu) : (int)(tl >> 16);
        int adiv;                                                // sign-extended low half: PRMT with the replicate-sign bit (not __byte_perm, whose selector digits are 3-bit)
        asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(adiv) : "r"(ta));
        int bdiv = (int)tb >> 16;
        if (MASK) { const bool bg = (int)tl < 0; adiv = bg ? p.adiv_bg : adiv; bdiv = bg ? p.bdiv_bg : bdiv; }
        const int tx = ify + adiv, tz = ify - bdiv;
        int x = cubic(tx), z = cubic(tz);
        if (min(tx, tz) <= 3390) {                               // very dark colours: the linear branch of the companding (rare)
            if (tx <= 3390) x = (tx * 108) / 841 - 290;
            if (tz <= 3390) z = (tz * 108) / 841 - 290;
        }
        const int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
        const int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
        const int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
        r = lds_u8(ig + (uint32_t)min(max(ro, 0), 4095));
        g = lds_u8(ig + (uint32_t)min(max(go, 0), 4095));
        b = lds_u8(ig + (uint32_t)min(max(bo, 0), 4095));
    }
    __device__ static void process(const Params& p, const OdAbs tab, uint32_t extra_addr, uint4* grp, Acc&) {
        const uint4 va = grp[0], vb = grp[1], vc = grp[2];
        const uint32_t w[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
            uint32_t c[12];
            px(p, extra_addr, lds_u32(row_addr(tab, wa, 0)), lds_u32_off4(row_addr(tab, wa, 1)), lds_u32_off4(row_addr(tab, wa, 2)), c[0], c[1], c[2]);
            px(p, extra_addr, lds_u32(row_addr(tab, wa, 3)), lds_u32_off4(row_addr(tab, wb, 0)), lds_u32_off4(row_addr(tab, wb, 1)), c[3], c[4], c[5]);
            px(p, extra_addr, lds_u32(row_addr(tab, wb, 2)), lds_u32_off4(row_addr(tab, wb, 3)), lds_u32_off4(row_addr(tab, wc, 0)), c[6], c[7], c[8]);
            px(p, extra_addr, lds_u32(row_addr(tab, wc, 1)), lds_u32_off4(row_addr(tab, wc, 2)), lds_u32_off4(row_addr(tab, wc, 3)), c[9], c[10], c[11]);
            o[3 * q] = pack4(c[0], c[1], c[2], c[3]);
            o[3 * q + 1] = pack4(c[4], c[5], c[6], c[7]);
            o[3 * q + 2] = pack4(c[8], c[9], c[10], c[11]);
        }
        grp[0] = make_uint4(o[0], o[1], o[2], o[3]);
        grp[1] = make_uint4(o[4], o[5], o[6], o[7]);
        grp[2] = make_uint4(o[8], o[9], o[10], o[11]);
    }
    __device__ static void run_flush(const Params&, int, Acc&) {}
    __device__ static void tile_end(const Params&, int, unsigned char*, int) {}
};

// ------------------------------------------------------------------------------------------------ per-tile kernels
// P: 90th percentile of all channel bytes -> brightness standardisation v -> trunc(clip(v * 255 / p)) (stain_utils.py:188-194),
// folded into the sRGB linearisation table.  skip = 1: get_mean_std / LuminosityStandardizer (no brightness step).
__global__ void __launch_bounds__(256) rein_plan_kernel(const unsigned* __restrict__ hist, int npx, const unsigned short* __restrict__ gamma,
                                                        int skip, float* __restrict__ gam2) {
    __shared__ unsigned h[256];
    __shared__ double p;
    const int tile = blockIdx.x, t = threadIdx.x;
    if (skip) { gam2[(size_t)tile * 256 + t] = (float)gamma[t]; return; }
    h[t] = hist[(size_t)tile * 256 + t];
    __syncthreads();
    if (t == 0) p = hist_percentile(h, 3ull * (unsigned long long)npx, 90.0);
    __syncthreads();
    gam2[(size_t)tile * 256 + t] = (float)gamma[trunc_clip_u8((double)t * 255.0 / p)];
}

enum ReinMode { REIN_STATS = 0, REIN_TRANSFORM = 1, REIN_LUMINOSITY = 2 };
struct ReinStatsArgs {
    const unsigned* lhist;         // [B][256]
    const unsigned long long* sums;// [B][4]
    int npx, mode;
    const double* tmeans;          // [3] (transform)
    const double* tstds;           // [3]
    double* means_out;             // [B,3] (stats)
    double* stds_out;
    int mask_background, lmax;
    double percentile;             // luminosity mode
    const int* lab2yf;             // [512]
    uint2* tl_ab;                  // [B][256]
    int32_t* status;
};
__device__ __forceinline__ int lab_adiv(int a) { return ((5 * a * 53687 + 128) >> 13) - 128 * 16384 / 500; }
__device__ __forceinline__ int lab_bdiv(int b) { return ((b * 41943 + 16) >> 9) - 128 * 16384 / 200 + 1; }
// S: means and population standard deviations exactly as cv.meanStdDev sees the planes of lab_split (stain_utils.py:146-186:
// I1 = float32(L) / float32(2.55), I2 = a - 128, I3 = b - 128, double accumulators), then the per-channel affine maps
// (normalizer.py:81-83) followed by merge_back's scale / offset, clip and truncation (stain_utils.py:160-172), composed with
// the first table step of the inverse conversion.  Same formulas as lab_tile_kernel.
__global__ void __launch_bounds__(256) rein_stats_kernel(ReinStatsArgs a) {
    __shared__ unsigned h[256];
    __shared__ double stat[6];
    __shared__ double pl;
    __shared__ unsigned any_tissue;
    const int tile = blockIdx.x, t = threadIdx.x;
    h[t] = a.lhist[(size_t)tile * 256 + t];
    if (t == 0) any_tissue = 0u;
    __syncthreads();
    const double n = (double)a.npx;
    if (a.mode == REIN_LUMINOSITY) {
        if (t == 0) pl = hist_percentile(h, (unsigned long long)a.npx, a.percentile);
        __syncthreads();
        const int L2 = trunc_clip_u8(255.0 * (double)t / pl);
        a.tl_ab[(size_t)tile * 256 + t] = make_uint2((uint32_t)a.lab2yf[2 * L2] | ((uint32_t)a.lab2yf[2 * L2 + 1] << 16),
                                                     ((uint32_t)lab_adiv(t) & 0xFFFFu) | ((uint32_t)lab_bdiv(t) << 16));
        return;
    }
    if (t == 0) {
        double s = 0.0, sq = 0.0;
        for (int v = 0; v < 256; ++v) {
            const double q = (double)((float)v / 2.55f);
            const double hv = (double)h[v];
            s += hv * q; sq += hv * q * q;
        }
        const double mean = s / n;
        double var = sq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[0] = mean; stat[3] = sqrt(var);
    } else if (t < 3) {
        // sum (v - 128) and sum (v - 128)^2 from the integer sums of v and v^2: exact, as the histogram-weighted double sums are
        const long long s1 = (long long)a.sums[(size_t)tile * 4 + 2 * (t - 1)], s2 = (long long)a.sums[(size_t)tile * 4 + 2 * (t - 1) + 1];
        const double s = (double)(s1 - 128LL * a.npx), sq = (double)(s2 - 256LL * s1 + 16384LL * a.npx);
        const double mean = s / n;
        double var = sq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[t] = mean; stat[3 + t] = sqrt(var);
    }
    __syncthreads();
    if (a.mode == REIN_STATS) {
        if (t < 3) { a.means_out[(size_t)tile * 3 + t] = stat[t]; a.stds_out[(size_t)tile * 3 + t] = stat[3 + t]; }
        return;
    }
    unsigned char cm[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double q = c == 0 ? (double)((float)t / 2.55f) : (double)(t - 128);
        const double m = (q - stat[c]) * (a.tstds[c] / stat[3 + c]) + a.tmeans[c];
        cm[c] = trunc_clip_u8(c == 0 ? m * 2.55 : m + 128.0);
    }
    const bool use_mask = a.mask_background != 0;
    const bool bg = use_mask && t > a.lmax;              // background: L = clip((254 + 0) * 2.55) = 255, a = b = 128 (normalizer.py:85-90)
    const int L2 = bg ? 255 : cm[0];
    a.tl_ab[(size_t)tile * 256 + t] = make_uint2((uint32_t)a.lab2yf[2 * L2] | ((uint32_t)a.lab2yf[2 * L2 + 1] << 16) | (bg ? 0x80000000u : 0u),
                                                 ((uint32_t)lab_adiv(cm[1]) & 0xFFFFu) | ((uint32_t)lab_bdiv(cm[2]) << 16));
    if (a.status) {
        if (use_mask && t <= a.lmax && h[t]) atomicOr(&any_tissue, 1u);
        __syncthreads();
        if (t == 0) a.status[tile] = (!use_mask || any_tissue) ? 0 : SB_STATUS_EMPTY_MASK;
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool reinhard_ring_eligible(const void* in, const void* out, int npx) {
    // (small tiles: the per-tile table refill -- 16 stores per thread and tile -- stops being negligible against one chunk of work)
    return (((uintptr_t)in | (uintptr_t)out) % 16 == 0) && npx % GROUP_PX == 0 && npx >= 16384;
}

// mode: ReinMode.  REIN_STATS: brightness (unless skip_brightness) + means/stds.  REIN_TRANSFORM: full transform into out.
// REIN_LUMINOSITY: LuminosityStandardizer into out.
int launch_reinhard_ring(sb_handle* h, const uint8_t* in, uint8_t* out, int B, int npx, int mode, int skip_brightness, const double* tmeans,
                         const double* tstds, double* means_out, double* stds_out, int mask_background, int lmax, double percentile,
                         int32_t* status, cudaStream_t st) {
    Scratch scratch(h, st);
    PassTimer pt{h, st};
    h->n_pass_ev = 0;
    const size_t nb = (size_t)B * 256;
    unsigned* hist = nullptr; unsigned* lhist = nullptr; float* gam2 = nullptr; unsigned long long* sums = nullptr; uint2* tl_ab = nullptr;
    cudaError_t e;
    // one allocation: [lhist | hist | sums] are zeroed together
    unsigned char* z = nullptr;
    const size_t z_bytes = nb * 4 * 2 + (size_t)B * 32;
    if ((e = scratch.get(&z, z_bytes)) != cudaSuccess) return SB_ERR_CUDA;
    lhist = reinterpret_cast<unsigned*>(z); hist = lhist + nb; sums = reinterpret_cast<unsigned long long*>(z + nb * 8);
    if ((e = scratch.get(&gam2, nb * sizeof(float))) != cudaSuccess) return SB_ERR_CUDA;
    if (mode != REIN_STATS && (e = scratch.get(&tl_ab, nb * sizeof(uint2))) != cudaSuccess) return SB_ERR_CUDA;
    if (cudaMemsetAsync(z, 0, z_bytes, st) != cudaSuccess) return SB_ERR_CUDA;
    const int skip = skip_brightness || mode == REIN_LUMINOSITY;
    int launches = 0;
    if (!skip) {
        NvtxRange r("reinhard: byte histogram");
        pt.mark("rein_ring<ByteHistOp>: histogram of all channel bytes");
        if (launch_rein_ring<ByteHistOp>(RingGeom{in, nullptr, B, npx}, ByteHistParams{hist}, h->num_sms, st) != 0) return SB_ERR_CUDA;
        ++launches;
    }
    pt.mark("rein_plan_kernel");
    rein_plan_kernel<<<B, 256, 0, st>>>(hist, npx, h->tab.gamma, skip, gam2);
    {
        NvtxRange r("reinhard: LAB statistics");
        pt.mark(mode == REIN_STATS ? "rein_ring<LabStatsOp>: forward LAB + statistics (read only)" : "rein_ring<LabStatsOp>: forward LAB + statistics, LAB bytes parked");
        const LabStatsParams p{gam2, h->tab.cbrt, lhist, sums};
        const int rc = mode == REIN_STATS ? launch_rein_ring<LabStatsOp<false>>(RingGeom{in, nullptr, B, npx}, p, h->num_sms, st)
                                          : launch_rein_ring<LabStatsOp<true>>(RingGeom{in, out, B, npx}, p, h->num_sms, st);
        if (rc != 0) return SB_ERR_CUDA;
    }
    ReinStatsArgs sa{};
    sa.lhist = lhist; sa.sums = sums; sa.npx = npx; sa.mode = mode; sa.tmeans = tmeans; sa.tstds = tstds; sa.means_out = means_out; sa.stds_out = stds_out;
    sa.mask_background = mask_background; sa.lmax = lmax; sa.percentile = percentile; sa.lab2yf = h->tab.lab2yf; sa.tl_ab = tl_ab; sa.status = status;
    pt.mark("rein_stats_kernel");
    rein_stats_kernel<<<B, 256, 0, st>>>(sa);
    launches += 3;
    if (mode != REIN_STATS) {
        NvtxRange r("reinhard: map + inverse conversion");
        pt.mark("rein_ring<LabInvOp>: map + inverse LAB, in place");
        LabInvParams p{tl_ab, h->tab.invgamma, 0, 0};
        p.adiv_bg = ((5 * 128 * 53687 + 128) >> 13) - 128 * 16384 / 500;
        p.bdiv_bg = ((128 * 41943 + 16) >> 9) - 128 * 16384 / 200 + 1;
        const RingGeom g{out, out, B, npx};               // in place: the LAB bytes parked by R2
        const int rc = (mode == REIN_TRANSFORM && mask_background) ? launch_rein_ring<LabInvOp<true>>(g, p, h->num_sms, st)
                                                                   : launch_rein_ring<LabInvOp<false>>(g, p, h->num_sms, st);
        if (rc != 0) return SB_ERR_CUDA;
        ++launches;
    }
    pt.mark("end");
    if (cudaGetLastError() != cudaSuccess) return SB_ERR_CUDA;
    h->launches += launches;
    return SB_OK;
}

}  // namespace sb
