// match_kernel.cuh — the hot path: persistent, warp-specialised sm_100a kernel that, for a list of
// (pair, query-strip) work items, computes the u8 x u8 -> s32 Gram tile on the 5th-gen tensor cores
// (tcgen05.mma.kind::i8, operands TMA-staged in 128B-swizzled shared memory, accumulators in TMEM) and reduces it
// in a fused epilogue to
//   * per query row: exact top-2 of d = ||q||^2 + ||r||^2 - 2 q.r  (lowest reference index on ties), and
//   * per reference row: the best query row (lowest query index on ties; global u64 atomicMin),
// so the N x M distance matrix never leaves the SM.
//
// Replaces the arithmetic of flann_find_nearest_neighbors_index(k=2) at
// /root/reference/SfM/src/graph/fine_matching_graph.cc:99 (and slam_gps.cc:463, feature_matching.cpp:44,336,409)
// with the brute-force/mutual-best semantics of the declared GPU matchers (SiftGPU.h:303-308, cudaSift sift.h:97).
//
// Roles per CTA ((4*STRIPS + 2) warps, one CTA per SM, persistent over work items):
//   warps 0 .. 4*STRIPS-1  epilogue: warp w owns TMEM lanes 32*(w%4).. of query strip w/4 (thread = query row)
//   warp  4*STRIPS         TMA producer (one elected lane)
//   warp  4*STRIPS+1       TMEM allocator + MMA issuer (one elected lane)
// Pipelines (all mbarrier based):
//   A ring (2 deep)   : query strips of a work item, STRIPS x [128 rows x 128 B]
//   B ring (STAGES)   : reference tiles [TILE_N rows x 128 B]
//   norm ring         : ||r||^2 of the tile's reference rows (no empty barrier: STAGES+2 slots cannot be
//                       overrun because the producer is throttled by B-empty, which trails TMEM-empty)
//   TMEM ring (2 deep): STRIPS x TILE_N int32 accumulator columns per buffer
#pragma once
#include <cstdint>
#include <climits>
#include <cuda.h>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

namespace msfm {

constexpr int kDim = 128;                 // bytes per packed descriptor row
constexpr int kStripRows = 128;           // MMA M
constexpr uint32_t kNormPad = 0x3FFFFFFFu; // ||r||^2 stored for pad rows: never wins a minimum
constexpr int kAlignRows = 256;           // every image starts on / is padded to a multiple of this many rows
constexpr int kBoxRows = 64;              // rows per TMA box (strips and reference tiles are loaded as 64-row boxes)

struct PairDesc {
    int32_t ref_img, qry_img;  // tensor-map slots
    int32_t ref_rows, qry_rows;
    int64_t ref_off, qry_off;  // first row in the arenas
    int64_t knn_off;           // first row of this pair in the per-batch kNN scratch
    int64_t col_off;           // first row of this pair in the per-batch column-best scratch
};

struct WorkItem {
    int32_t pair;  // index into PairDesc[]
    int32_t row0;  // first query row of the item (multiple of STRIPS*128)
};

struct MatchKernelParams {
    const CUtensorMap *maps;      // [max_images] one 2-D map per image: {128 B, rows}, box {128 B, kBoxRows rows}, SW128
    const uint32_t *norms;        // arena of squared norms (pad rows = kNormPad)
    const PairDesc *pairs;
    const WorkItem *items;
    int32_t n_items;
    unsigned long long *stats;    // optional debug counter (slow-path group visits); null in production
    uint32_t debug_flags;         // bit 0: skip the exact slow path (timing experiments only; results are wrong)
    int4 *knn;                    // [sum qry_rows] {id0, id1, d0, d1}; id = -1 / d = INT_MAX when absent
    unsigned long long *colbest;  // [sum ref_rows] (d << 32 | query row), initialised to ~0
};

template <int STRIPS, int TILE_N, int STAGES>
struct MatchKernelCfg {
    static constexpr int kEpiWarps = 4 * STRIPS;
    static constexpr int kThreads = (kEpiWarps + 2) * 32;
    // The producer may refill a norm slot once the MMA that shares its B stage has completed; that MMA was issued after
    // the epilogue released the TMEM buffer two tiles earlier, and the release happens one chunk before the epilogue
    // stops reading that tile's norms: STAGES + 3 slots can therefore never be overrun.
    static constexpr int kNormSlots = STAGES + 3;
    static constexpr int kTmemBufs = 2;
    static constexpr int kTmemCols = kTmemBufs * STRIPS * TILE_N;
    static constexpr int kABytes = STRIPS * kStripRows * kDim;  // one A buffer
    static constexpr int kBBytes = TILE_N * kDim;               // one B stage
    static constexpr int kSmemA = 0;
    static constexpr int kSmemB = kSmemA + 2 * kABytes;
    static constexpr int kSmemNorm = kSmemB + STAGES * kBBytes;
    static constexpr int kSmemBar = kSmemNorm + kNormSlots * TILE_N * 4;
    static constexpr int kNumBars = 2 + 2 + STAGES + STAGES + kNormSlots + 2 + 2;
    static constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
    static constexpr int kSmemBytes = kSmemTmemPtr + 16;
    static constexpr int kSmemAlloc = kSmemBytes + 1024;  // slack for manual 1024-byte alignment
    static_assert(kTmemCols == 32 || kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512,
                  "TMEM allocation must be a power of two in [32, 512] columns");
    static_assert(TILE_N % kBoxRows == 0 && TILE_N <= 256, "reference tile is loaded as 64-row TMA boxes; UMMA N <= 256");
    static_assert(kThreads <= 1024, "too many warps");
    static_assert(kAlignRows % TILE_N == 0, "image padding must cover whole reference tiles");
};

template <int STRIPS, int TILE_N, int STAGES, bool COLBEST>
__global__ void __launch_bounds__(MatchKernelCfg<STRIPS, TILE_N, STAGES>::kThreads, 1)
match_pairs_kernel(const MatchKernelParams p) {
    using Cfg = MatchKernelCfg<STRIPS, TILE_N, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    uint8_t *sA = smem + Cfg::kSmemA;
    uint8_t *sB = smem + Cfg::kSmemB;
    uint32_t *sNorm = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemNorm);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemBar);
    uint64_t *a_full = bars;                       // [2]
    uint64_t *a_empty = a_full + 2;                // [2]
    uint64_t *b_full = a_empty + 2;                // [STAGES]
    uint64_t *b_empty = b_full + STAGES;           // [STAGES]
    uint64_t *n_full = b_empty + STAGES;           // [kNormSlots]
    uint64_t *t_full = n_full + Cfg::kNormSlots;   // [2]
    uint64_t *t_empty = t_full + 2;                // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemTmemPtr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == Cfg::kEpiWarps && lane == 0) {
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < Cfg::kNormSlots; ++i) ptx::mbar_init(&n_full[i], 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], Cfg::kEpiWarps); }
        ptx::fence_mbar_init();
    }
    if (warp == Cfg::kEpiWarps + 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == Cfg::kEpiWarps) {
        // =========================================================== TMA producer
        if (ptx::elect_one()) {
            uint32_t g = 0, a = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++a) {
                const WorkItem wi = p.items[item];
                const PairDesc pd = p.pairs[wi.pair];
                const CUtensorMap *qmap = p.maps + pd.qry_img;
                const CUtensorMap *rmap = p.maps + pd.ref_img;
                const uint32_t abuf = a & 1;
                ptx::mbar_wait(&a_empty[abuf], ((a >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&a_full[abuf], Cfg::kABytes);
#pragma unroll
                for (int s = 0; s < STRIPS * kStripRows / kBoxRows; ++s)
                    ptx::tma_load_2d(sA + abuf * Cfg::kABytes + s * kBoxRows * kDim, qmap, &a_full[abuf], 0,
                                     wi.row0 + s * kBoxRows);
                const int ntiles = (pd.ref_rows + TILE_N - 1) / TILE_N;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const uint32_t st = g % STAGES;
                    ptx::mbar_wait(&b_empty[st], ((g / STAGES) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&b_full[st], Cfg::kBBytes);
#pragma unroll
                    for (int h = 0; h < TILE_N / kBoxRows; ++h)
                        ptx::tma_load_2d(sB + st * Cfg::kBBytes + h * kBoxRows * kDim, rmap, &b_full[st], 0,
                                         t * TILE_N + h * kBoxRows);
                    const uint32_t ns = g % Cfg::kNormSlots;
                    ptx::mbar_arrive_expect_tx(&n_full[ns], TILE_N * 4);
                    ptx::bulk_load_1d(sNorm + ns * TILE_N, p.norms + pd.ref_off + (int64_t)t * TILE_N, TILE_N * 4, &n_full[ns]);
                }
            }
        }
    } else if (warp == Cfg::kEpiWarps + 1) {
        // =========================================================== MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(kStripRows, TILE_N, 0, 0);
            uint32_t g = 0, a = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++a) {
                const WorkItem wi = p.items[item];
                const PairDesc pd = p.pairs[wi.pair];
                const uint32_t abuf = a & 1;
                ptx::mbar_wait(&a_full[abuf], (a >> 1) & 1);
                const uint32_t a_addr = ptx::smem_u32(sA + abuf * Cfg::kABytes);
                const int ntiles = (pd.ref_rows + TILE_N - 1) / TILE_N;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    const uint32_t st = g % STAGES;
                    const uint32_t buf = g & 1;
                    ptx::mbar_wait(&b_full[st], (g / STAGES) & 1);
                    ptx::mbar_wait(&t_empty[buf], ((g >> 1) & 1) ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t b_addr = ptx::smem_u32(sB + st * Cfg::kBBytes);
#pragma unroll
                    for (int s = 0; s < STRIPS; ++s) {
                        const uint32_t d_tmem = tmem_base + buf * (STRIPS * TILE_N) + s * TILE_N;
#pragma unroll
                        for (int k = 0; k < kDim / 32; ++k) {
                            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + s * kStripRows * kDim + k * 32);
                            const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32);
                            ptx::mma_i8_ss(d_tmem, da, db, idesc, k > 0 ? 1u : 0u);
                        }
                    }
                    ptx::mma_commit(&b_empty[st]);  // B stage reusable once these MMAs have read it
                    ptx::mma_commit(&t_full[buf]);  // accumulators ready for the epilogue
                }
                ptx::mma_commit(&a_empty[abuf]);    // A buffer reusable
            }
        }
    } else {
        // =========================================================== epilogue (thread = query row)
        // Scores s = 2*acc - ||r||^2 (maximise; d = ||q||^2 - s).  Fast path: the running maximum of the RAW accumulators of
        // a group of 8 columns is compared with T = floor((S1 + min||r||^2 over the tile) / 2): if no lane of the warp
        // exceeds it, no column of the group can enter any lane's top-2 (2*acc - nb_j <= 2*acc - nbmin <= S1) and the group
        // costs ~0.7 instructions per element.  Otherwise the group is re-scanned exactly (slow path).  Ties never displace
        // (strict >) and columns are visited in ascending order, hence lowest-index tie-breaking.
        const int strip = warp >> 2;
        const int quarter = warp & 3;
        const int row_local = strip * kStripRows + quarter * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + strip * TILE_N;
        constexpr int kChunks = TILE_N / 32;
        constexpr int kAbsent = -0x20000000;  // scores below this are pad columns / "no neighbour"
        uint32_t g = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const WorkItem wi = p.items[item];
            const PairDesc pd = p.pairs[wi.pair];
            const int q = wi.row0 + row_local;
            const bool valid = q < pd.qry_rows;
            const uint32_t na = valid ? p.norms[pd.qry_off + q] : 0u;
            // column key = (na + 2^23 - 2 acc) * 32 + lane  (29 bits); invalid rows sit above every valid key
            const uint32_t kcol = valid ? (((na + (1u << 23)) << 5) | (uint32_t)lane) : ((1u << 29) | (uint32_t)lane);
            // rows past the image end hold zeros; park their state where nothing can trigger the slow path
            int S0 = valid ? INT_MIN : 0x20000000, S1 = S0, J0 = -1, J1 = -1;
            const int ntiles = (pd.ref_rows + TILE_N - 1) / TILE_N;
            for (int t = 0; t < ntiles; ++t, ++g) {
                const uint32_t buf = g & 1;
                const uint32_t ns = g % Cfg::kNormSlots;
                ptx::mbar_wait(&n_full[ns], (g / Cfg::kNormSlots) & 1);
                const uint32_t *nb = sNorm + ns * TILE_N;
                uint32_t nbmin = nb[lane];
#pragma unroll
                for (int k = 1; k < TILE_N / 32; ++k) nbmin = min(nbmin, nb[lane + 32 * k]);
                nbmin = __reduce_min_sync(0xFFFFFFFFu, nbmin);
                int T = (S1 + (int)nbmin) >> 1;
                ptx::mbar_wait(&t_full[buf], (g >> 1) & 1);
                ptx::tc_fence_after();
                const uint32_t tile_taddr = lane_taddr + buf * (STRIPS * TILE_N);
                uint32_t acc[2][32];
                ptx::tmem_ld_32x32b_x32(tile_taddr, acc[0]);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    ptx::tmem_ld_wait();
                    if (c + 1 < kChunks) {
                        ptx::tmem_ld_32x32b_x32(tile_taddr + (c + 1) * 32, acc[(c + 1) & 1]);
                    } else {
                        // every TMEM read of this buffer has landed in registers: hand it back to the MMA warp early
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
                    }
                    const uint32_t(&a)[32] = acc[c & 1];
                    const int jbase = t * TILE_N + c * 32;
                    int m[4];
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        const int m1 = __vimax3_s32((int)a[8 * gq], (int)a[8 * gq + 1], (int)a[8 * gq + 2]);
                        const int m2 = __vimax3_s32((int)a[8 * gq + 3], (int)a[8 * gq + 4], (int)a[8 * gq + 5]);
                        m[gq] = __vimax3_s32((int)a[8 * gq + 6], (int)a[8 * gq + 7], max(m1, m2));
                    }
                    const int mall = max(max(m[0], m[1]), max(m[2], m[3]));
                    if (__any_sync(0xFFFFFFFFu, mall > T) && !(p.debug_flags & 1u)) {
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            if (__any_sync(0xFFFFFFFFu, m[gq] > T)) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    // only columns that can still beat some lane's second best are scored exactly
                                    if (__any_sync(0xFFFFFFFFu, (int)a[8 * gq + k] > T)) {
                                        const int sc = 2 * (int)a[8 * gq + k] - (int)nb[c * 32 + 8 * gq + k];
                                        const int j = jbase + 8 * gq + k;
                                        if (sc > S1) {
                                            if (sc > S0) { S1 = S0; J1 = J0; S0 = sc; J0 = j; }
                                            else { S1 = sc; J1 = j; }
                                        }
                                        T = (S1 + (int)nbmin) >> 1;
                                    }
                                }
                                if (p.stats && lane == 0) atomicAdd(p.stats, 1ull);
                            }
                        }
                    }
                    if (COLBEST) {
                        uint32_t mycol = 0xFFFFFFFFu;
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const uint32_t key = kcol - 64u * a[k];
                            const uint32_t r = __reduce_min_sync(0xFFFFFFFFu, key);
                            if (lane == k) mycol = r;
                        }
                        const int j = jbase + lane;
                        if (j < pd.ref_rows && mycol < (1u << 29)) {
                            const uint32_t d = (mycol >> 5) - (1u << 23) + nb[c * 32 + lane];
                            const uint32_t qsrc = (uint32_t)(wi.row0 + strip * kStripRows + quarter * 32) + (mycol & 31u);
                            const unsigned long long val = ((unsigned long long)d << 32) | qsrc;
                            unsigned long long *dst = p.colbest + pd.col_off + j;
                            if (val < *reinterpret_cast<volatile unsigned long long *>(dst)) atomicMin(dst, val);
                        }
                    }
                }
            }
            if (valid) {
                int4 out;
                out.x = (S0 > kAbsent) ? J0 : -1;
                out.y = (S1 > kAbsent) ? J1 : -1;
                out.z = (S0 > kAbsent) ? (int)na - S0 : INT_MAX;
                out.w = (S1 > kAbsent) ? (int)na - S1 : INT_MAX;
                p.knn[pd.knn_off + q] = out;
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == Cfg::kEpiWarps + 1) ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace msfm
