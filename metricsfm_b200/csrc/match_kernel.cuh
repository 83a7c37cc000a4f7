// match_kernel.cuh — the hot path: persistent, warp-specialised sm_100a kernel that, for a list of
// (pair, query-strip-group) work items, computes the u8 x u8 -> s32 Gram tile on the 5th-gen tensor cores
// (tcgen05.mma.kind::i8, operands TMA-staged in 128B-swizzled shared memory, accumulators in TMEM) and reduces it
// in a fused epilogue to
//   * per query row: exact top-2 of d = ||q||^2 + ||r||^2 - 2 q.r  (lowest reference index on ties),
// so the N x M distance matrix never leaves the SM.  The mutual cross-check ("best query of a reference row") is the
// same reduction with the roles of the two images swapped; the host schedules those items into the same launch.
//
// Replaces the arithmetic of flann_find_nearest_neighbors_index(k=2) at
// /root/reference/SfM/src/graph/fine_matching_graph.cc:99 (and slam_gps.cc:463, feature_matching.cpp:44,336,409)
// with the brute-force/mutual-best semantics of the declared GPU matchers (SiftGPU.h:303-308, cudaSift sift.h:97).
//
// Roles per CTA (4*STRIPS*CSPLIT epilogue warps + 1 + STRIPS auxiliary warps padded to whole warpgroups; one CTA per SM,
// persistent over work items):
//   epilogue warps  warp w owns TMEM lanes 32*(w%4).. of query strip (w/4)%STRIPS and the column share w/(4*STRIPS)
//                   of every tile (thread = query row x column share)
//   next warp       TMA producer (one elected lane)
//   next STRIPS     MMA issuers, one per query strip (one elected lane each; the first also owns the TMEM allocation):
//                   spreads the issue instructions over the four SM sub-partitions and lets every strip advance as soon
//                   as ITS accumulator buffer is free
//   (padding warps) idle; the auxiliary warpgroups hand their registers to the epilogue warps (setmaxnreg)
//   The service warps share issue slots with the epilogue warps of their sub-partition: their tile loops are unrolled
//   over the B ring (immediate addresses) and they sleep between polls (see the comment at the role dispatch).
// MODE 1 ("collect", float regime) replaces the top-2 merge by "append every element under the row's fixed threshold
// to an event list"; everything else is shared.
// Work distribution (template parameter DYN):
//   DYN = false  static: CTA b takes items b, b + grid, ...; every role walks that list on its own.
//   DYN = true   the loader draws items from a device counter (atomicAdd) and publishes each index in a small shared-memory
//                ring (s_item / i_full barriers) as soon as it knows it, i.e. one to two items ahead of the MMAs; issuers and
//                epilogue warps read it there (-1 ends the CTA).  Items go to whichever CTA is free: a CTA that cannot be
//                placed at launch (a collective or a packer kernel occupies its SM) simply finds nothing left.
// Pipelines (all mbarrier based):
//   A ring (2 deep)     : query strips of a work item, STRIPS x [128 rows x 128 B]
//   B ring (STAGES)     : reference tiles [TILE_N rows x 128 B]
//   column-key ring     : ckey_j = -8*||r_j||^2 + (7 - j%8) of the tile's reference rows + the tile's smallest norm; they
//                         complete the B stage's "full" barrier together with the tile (no barriers of their own: the
//                         epilogue reads them behind the tile's MMAs; for the slot count see MatchKernelCfg::kKeySlots)
//   TMEM                : TBUFS x STRIPS accumulator blocks of TILE_N int32 columns, each with its own full/empty
//                         barrier pair, so the MMA of tile t+1 runs under the epilogue of tile t
#pragma once
#include <cstdint>
#include <climits>
#include <cuda.h>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

namespace msfm {

constexpr int kDim = 128;        // bytes per packed descriptor row
constexpr int kStripRows = 128;  // MMA M
constexpr int kAlignRows = 256;  // every image starts on / is padded to a multiple of this many rows
constexpr int kBoxRows = 64;     // rows per TMA box (strips and reference tiles are loaded as 64-row boxes)
constexpr int kKeyTileRows = 64; // granularity of the per-tile minimum-norm table (MatchKernelParams::tilemin) = TILE_N

// Per-row side table ("column keys").  For row j with squared norm nb:  ckey = -8*nb + (7 - j%8).
//   * exact packed score key of element (q, j):  16*acc + ckey = 8*(2*acc - nb) + (7 - j%8)
//     (max key = best score, then lowest column inside a group of 8)
//   * nb = (7 - ckey) >> 3
// Pad rows carry kNormPad so that they lose against every real column.
constexpr uint32_t kNormPad = 0x0FFFFFF0u;
__host__ __device__ constexpr int32_t make_ckey(uint32_t nb, int j) { return -8 * (int32_t)nb + (7 - (j & 7)); }
__host__ __device__ constexpr int32_t ckey_to_norm(int32_t ckey) { return (7 - ckey) >> 3; }

struct PairDesc {
    int32_t ref_img, qry_img;  // tensor-map slots (the slot after the last image maps the candidate scratch)
    int32_t ref_rows, qry_rows;
    int64_t ref_off, qry_off;  // first row in the column-key arena (query side: candidate scratch when cand_idx >= 0)
    int64_t knn_off;           // first row of this pair in the per-batch kNN scratch
    int32_t qry_row_base;      // row coordinate of query row 0 in the query tensor map (0 for whole-image maps)
    int32_t cand_idx;          // >= 0: the query rows are this pair's gathered mutual-check candidates and their
                               // number is counts[cand_idx] (known only on the device); -1: ordinary image rows
    float fscale2;             // float regime: (quantisation scale)^2 when both images keep their float rows, else 0
    int32_t pad_;
    int64_t col_off;           // first entry of this pair in the per-batch column-best table (mutual check, one per ref row)
};

struct WorkItem {
    int32_t pair;  // index into PairDesc[]
    int32_t row0;  // first query row of the item (multiple of STRIPS*128)
};

struct MatchKernelParams {
    const CUtensorMap *maps;      // [max_images] one 2-D map per image: {128 B, rows}, box {128 B, kBoxRows rows}, SW128
    const int32_t *ckeys;         // arena of column keys (see make_ckey)
    const int4 *tilemin;          // [arena rows / kKeyTileRows] {smallest squared norm of the tile's rows, -, -, -}: refreshed per batch
                                  // by tile_min_kernel; images start on multiples of kAlignRows, so tiles never straddle images
    const int32_t *cand_ckeys;    // column keys of the gathered candidate rows (query side of cand_idx >= 0 pairs)
    const int32_t *cand_d0;       // squared distance of each candidate to the query row that proposed it
    const int32_t *counts;        // per-pair candidate counts
    const PairDesc *pairs;
    const WorkItem *items;
    int32_t n_items;
    uint32_t debug_flags;         // timing experiments only (results are wrong): 1 = skip the exact phase,
                                  // 2 = no epilogue work at all, 4 = TMEM loads + hand-back only
    unsigned long long *stats;    // optional debug counter (slow-path group visits); null in production
    int4 *knn;                    // [sum qry_rows][CSPLIT] partial {id0, id1, d0, d1} per column share; id = -1 /
                                  // d = INT_MAX when absent; consumers merge the shares with merge_knn_shares()
    // MODE 1 ("collect", float regime): instead of a top-2 the kernel lists every (query row, reference row) whose
    // squared distance is <= cand_d0[row], as {scratch row, reference row, cand_idx, 0}
    int4 *events;
    unsigned int *event_count;    // total events seen (may exceed event_cap: the consumer then falls back)
    uint32_t event_cap;
    const unsigned int *gate;     // optional: the whole launch returns at once when *gate == 0 (mutual twin pass with no
                                  // pair left for the tensor path, see select_candidates_kernel)
    unsigned int *next_item;      // work-item counter of this launch (zero at launch): CTAs draw items from it, so a CTA that
                                  // starts late (another kernel sits on its SM) or meets cheap items does not hold the launch up
    uint32_t prune_q8;            // "dead row" rule of the forward pass (see the epilogue): a row whose two nearest neighbours
                                  // so far satisfy d0 > (prune_q8 / 256) * d1 only follows its NEAREST neighbour exactly.
                                  // kNoPrune = off (exact 2-NN of every row: msfm_knn2, twin and collect passes)
};
constexpr uint32_t kNoPrune = 4096;  // ((d1 >> 8) + 1) * 4096 >= 16 * d1 > d0 for every row, and fits in 32 bits (d < 2^25)

template <int STRIPS, int TILE_N, int STAGES, int CSPLIT, int TBUFS>
struct MatchKernelCfg {
    // CSPLIT warps share one (strip, 32-row quarter) and split the tile's columns between them; TBUFS accumulator
    // buffers per strip let the MMA of tile t+1 run under the epilogue of tile t.
    static constexpr int kEpiWarps = 4 * STRIPS * CSPLIT;
    static constexpr int kAuxWarps = (1 + STRIPS + 3) / 4 * 4;       // TMA producer + one MMA issuer per strip, padded to warpgroups
    static constexpr int kThreads = (kEpiWarps + kAuxWarps) * 32;
    // Registers: the launch gives every thread 65536 / kThreads (a multiple of 8); the auxiliary warpgroups hand most of
    // their share to the epilogue warps (setmaxnreg), which hold a 64-column accumulator tile each.
    static constexpr int kAuxRegs = 32;
    static constexpr int kLaunchRegs = 65536 / kThreads / 8 * 8;
    static constexpr int kEpiRegs = (kLaunchRegs * kThreads - kAuxWarps * 32 * kAuxRegs) / (kEpiWarps * 32) / 8 * 8;
    static constexpr int kColsPerWarp = TILE_N / CSPLIT;
    // The producer may refill a key slot once the MMAs that share its B stage have completed; those were issued after
    // every epilogue warp released the accumulator buffer TBUFS tiles earlier; a warp releases a buffer before it has
    // finished reading that tile's keys, but it must finish before it can release the next one:
    // tiles <= t-STAGES-TBUFS-1 are done when tile t is loaded.  +1 for margin.
    static constexpr int kKeySlots = 2 * STAGES;  // >= STAGES + TBUFS + 1; two rounds of the B ring, so slot = stage + STAGES * (round & 1)
    static_assert(2 * STAGES >= STAGES + TBUFS + 1 && STAGES % (2 * TBUFS) == 0, "ring position must fix key slot, accumulator buffer and its phase");
    static constexpr int kTmemCols = TBUFS * STRIPS * TILE_N;
    static constexpr int kABytes = STRIPS * kStripRows * kDim;  // one A buffer
    static constexpr int kBBytes = TILE_N * kDim;               // one B stage
    static constexpr int kSmemA = 0;
    static constexpr int kSmemB = kSmemA + 2 * kABytes;
    static constexpr int kSmemKey = kSmemB + STAGES * kBBytes;
    static constexpr int kSmemTmin = kSmemKey + kKeySlots * TILE_N * 4;        // [kKeySlots] int4: the tile's smallest reference norm
    static constexpr int kSmemShare = kSmemTmin + kKeySlots * 16;              // [STRIPS*128 rows][CSPLIT] int4
    static constexpr int kSmemBar = kSmemShare + STRIPS * kStripRows * CSPLIT * 16;
    static constexpr int kItemSlots = 8;  // published work-item indices (the loader runs at most two items ahead of the MMAs)
    static constexpr int kNumBars = 2 + 2 + STAGES + STAGES + 2 * TBUFS * STRIPS + kItemSlots;
    static constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
    static constexpr int kSmemItems = kSmemTmemPtr + 16;
    static constexpr int kSmemBytes = kSmemItems + kItemSlots * 4;
    static constexpr int kSmemAlloc = kSmemBytes + 1024;  // slack for manual 1024-byte alignment
    static_assert(kTmemCols == 32 || kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512,
                  "TMEM allocation must be a power of two in [32, 512] columns");
    static_assert(TILE_N % kBoxRows == 0 && TILE_N <= 256, "reference tile is loaded as 64-row TMA boxes; UMMA N <= 256");
    static_assert(kAlignRows % TILE_N == 0, "image padding must cover whole reference tiles");
    static_assert(TILE_N == kKeyTileRows, "the minimum-norm table is kept per reference tile");
    static_assert(CSPLIT == 1 || CSPLIT == 2, "one or two warps per (strip, quarter)");
    static_assert(kColsPerWarp % 16 == 0 && kColsPerWarp / 8 <= 32, "one flag bit per group of 8 columns");
    static_assert(kThreads <= 1024, "too many warps");
    static_assert(kSmemAlloc <= 227 * 1024, "shared memory budget");
};

// Top-2 (largest, second largest) of eight distinct keys: 17 min/max operations, depth 4, no branches.
__device__ __forceinline__ void top2_of8(const int (&k)[8], int &g0, int &g1) {
    const int h0 = max(k[0], k[1]), l0 = min(k[0], k[1]);
    const int h1 = max(k[2], k[3]), l1 = min(k[2], k[3]);
    const int h2 = max(k[4], k[5]), l2 = min(k[4], k[5]);
    const int h3 = max(k[6], k[7]), l3 = min(k[6], k[7]);
    const int H0 = max(h0, h1), L0 = __vimax3_s32(min(h0, h1), l0, l1);
    const int H1 = max(h2, h3), L1 = __vimax3_s32(min(h2, h3), l2, l3);
    g0 = max(H0, H1);
    g1 = __vimax3_s32(min(H0, H1), L0, L1);
}

// Merge a group's two best (score, column) candidates a >= b into the running (S0,J0) >= (S1,J1).  Candidates come
// from higher column indices than anything in the running state, so they must beat it strictly.
__device__ __forceinline__ void merge_top2(int sa, int ja, int sb, int jb, int &S0, int &J0, int &S1, int &J1) {
    const bool t0 = sa > S0, t1 = sa > S1, u = sb > S0;
    const int x1s = u ? sb : S0, x1j = u ? jb : J0;    // second place when a takes first
    const int y1s = t1 ? sa : S1, y1j = t1 ? ja : J1;  // second place when a does not
    S1 = t0 ? x1s : y1s;
    J1 = t0 ? x1j : y1j;
    S0 = t0 ? sa : S0;
    J0 = t0 ? ja : J0;
}

// The score a row prunes against: its second best, or — "dead row" rule, see the epilogue — its best when the running pair
// already violates d0 <= rho * d1 (rho = rho8 / 256, rounded so that the test errs on the live side; d = na - S).
// Placeholder states (nothing or one neighbour seen, pad rows) never count as dead.
__device__ __forceinline__ int prune_score(int S0, int S1, int na, int rho8, int absent) {
    const int d0 = (int)((uint32_t)na - (uint32_t)S0), d1 = (int)((uint32_t)na - (uint32_t)S1);
    const bool dead = (S1 > absent) && (d0 > (d1 >> 8) * rho8 + rho8);
    return dead ? S0 : S1;
}

template <int STRIPS, int TILE_N, int STAGES, int CSPLIT, int TBUFS, bool DEBUG, int MODE = 0, bool DYN = false>
__global__ void __launch_bounds__(MatchKernelCfg<STRIPS, TILE_N, STAGES, CSPLIT, TBUFS>::kThreads, 1)
match_pairs_kernel(const MatchKernelParams p) {
    using Cfg = MatchKernelCfg<STRIPS, TILE_N, STAGES, CSPLIT, TBUFS>;
    // No static shared memory is declared, so the dynamic window starts at shared offset 0 (1024-byte aligned, as
    // the 128B-swizzled operand tiles require); checked once below.
    extern __shared__ __align__(1024) uint8_t smem[];

    uint8_t *sA = smem + Cfg::kSmemA;
    uint8_t *sB = smem + Cfg::kSmemB;
    int32_t *sKey = reinterpret_cast<int32_t *>(smem + Cfg::kSmemKey);
    int4 *sTmin = reinterpret_cast<int4 *>(smem + Cfg::kSmemTmin);
    int4 *sShare = reinterpret_cast<int4 *>(smem + Cfg::kSmemShare);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Cfg::kSmemBar);
    uint64_t *a_full = bars;                         // [2]
    uint64_t *a_empty = a_full + 2;                  // [2]
    uint64_t *b_full = a_empty + 2;                  // [STAGES]
    uint64_t *b_empty = b_full + STAGES;             // [STAGES]
    uint64_t *t_full = b_empty + STAGES;             // [TBUFS][STRIPS]
    uint64_t *t_empty = t_full + TBUFS * STRIPS;     // [TBUFS][STRIPS]
    uint64_t *i_full = t_empty + TBUFS * STRIPS;     // [kItemSlots] (DYN) item index published
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(smem + Cfg::kSmemTmemPtr);
    volatile int32_t *s_item = reinterpret_cast<volatile int32_t *>(smem + Cfg::kSmemItems);

    const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction: lets the compiler keep
                                                                            // everything derived from it in uniform registers
    const int lane = threadIdx.x & 31;
    if (p.gate != nullptr && *p.gate == 0u) return;  // uniform over the grid: nothing was routed to this pass
    if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

    if (warp == Cfg::kEpiWarps && lane == 0) {
        // A buffers and B stages are released by the commits of all STRIPS MMA issuers
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], STRIPS); }
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], STRIPS); }
        for (int i = 0; i < TBUFS * STRIPS; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 4 * CSPLIT); }
        for (int i = 0; i < Cfg::kItemSlots; ++i) ptx::mbar_init(&i_full[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == Cfg::kEpiWarps + 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    if (warp < Cfg::kEpiWarps)  // threshold-sharing slots {score, item tag} start out "no information"
        for (int i = threadIdx.x; i < STRIPS * kStripRows * CSPLIT; i += Cfg::kEpiWarps * 32) sShare[i] = make_int4(0, 0, -1, 0);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= Cfg::kEpiWarps) {
      ptx::setmaxnreg_dec<Cfg::kAuxRegs>();
      // Aux warp x = warp - kEpiWarps: x = 0 is the TMA loader, x = 1 .. STRIPS are the MMA issuers (one per SM
      // sub-partition), the rest is padding up to whole warpgroups (setmaxnreg) and goes straight to the final barrier.
      // Measured with idle dummy warps: a warp that polls an mbarrier (or merely loops on nanosleep(20)) slows the four
      // epilogue warps of its sub-partition by 7 %, a warp that sleeps 1 us between polls costs nothing, and every
      // instruction a service warp executes per tile costs its sub-partition ~1.4 cycles.  Hence: long sleeps where the
      // latency is hidden anyway (the loader runs STAGES tiles ahead, the issuers have a spare accumulator buffer), and
      // tile loops that walk the B ring unrolled, so that stage / barrier / key-slot addresses are immediates.
      // Spreading the loads over four loader warps was measured too: more total service work, slower.
#ifndef MSFM_LOADER_SLEEP
#define MSFM_LOADER_SLEEP 1000
#endif
#ifndef MSFM_ISSUER_SLEEP
#define MSFM_ISSUER_SLEEP 200
#endif
      constexpr unsigned kLoaderSleepNs = MSFM_LOADER_SLEEP, kIssuerSleepNs = MSFM_ISSUER_SLEEP;
#define LOADER_WAIT(bar, par) ptx::mbar_wait_backoff<kLoaderSleepNs>(bar, par)
#define ISSUER_WAIT(bar, par) ptx::mbar_wait_backoff<kIssuerSleepNs>(bar, par)
      const int x = warp - Cfg::kEpiWarps;
      if (x == 0) {
        // =========================================================== TMA loader
        // The tile loop walks the B ring (unrolled over its STAGES positions, so stage, barrier and key-slot addresses are
        // immediates); items are switched inside it.  Every instruction here is taken from the issue slots of the
        // epilogue warps on the same sub-partition (~1.4 cycles each per tile), hence the lean loop and the long sleeps.
        if (ptx::elect_one()) {
            int cur = (int)blockIdx.x - (int)gridDim.x, left = 0, t = 0;
            uint32_t a = 0, round = 0;  // items started, trips round the B ring
            const CUtensorMap *rmap = nullptr;
            const int32_t *keyp = nullptr;
            const int4 *tminp = nullptr;
            auto next_item = [&]() -> bool {  // next item with work; fetches its query strips
                for (;;) {
                    const int idx = DYN ? (int)atomicAdd(p.next_item, 1u) : (cur += (int)gridDim.x);
                    if (idx >= p.n_items) {
                        if (DYN) {  // end marker
                            s_item[a % Cfg::kItemSlots] = -1;
                            ptx::mbar_arrive(&i_full[a % Cfg::kItemSlots]);
                        }
                        return false;
                    }
                    const WorkItem wi = p.items[idx];
                    const PairDesc pd = p.pairs[wi.pair];
                    if (pd.cand_idx >= 0 && wi.row0 >= p.counts[pd.cand_idx]) continue;  // item past the candidate list
                    if (DYN) {  // published before the strips are even requested: the other roles prefetch the descriptors
                        s_item[a % Cfg::kItemSlots] = idx;
                        ptx::mbar_arrive(&i_full[a % Cfg::kItemSlots]);
                    }
                    const CUtensorMap *qmap = p.maps + pd.qry_img;
                    const uint32_t abuf = a & 1;
                    LOADER_WAIT(&a_empty[abuf], ((a >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&a_full[abuf], Cfg::kABytes);
#pragma unroll
                    for (int s = 0; s < STRIPS * kStripRows / kBoxRows; ++s)
                        ptx::tma_load_2d(sA + abuf * Cfg::kABytes + s * kBoxRows * kDim, qmap, &a_full[abuf], 0,
                                         pd.qry_row_base + wi.row0 + s * kBoxRows);
                    rmap = p.maps + pd.ref_img;
                    keyp = p.ckeys + pd.ref_off;
                    tminp = p.tilemin + pd.ref_off / kKeyTileRows;
                    left = (pd.ref_rows + TILE_N - 1) / TILE_N;
                    t = 0;
                    ++a;
                    return true;
                }
            };
            bool more = next_item();
            while (more) {
#pragma unroll
                for (int st = 0; st < STAGES; ++st) {
                    if (!more) break;
                    LOADER_WAIT(&b_empty[st], (round & 1) ^ 1);
                    // one barrier for the tile, its column keys and its minimum norm: the MMAs are issued behind it and the
                    // epilogue reads the keys behind the MMAs, so it needs no barrier of its own for them
                    ptx::mbar_arrive_expect_tx(&b_full[st], Cfg::kBBytes + TILE_N * 4 + 16);
#pragma unroll
                    for (int h = 0; h < TILE_N / kBoxRows; ++h)
                        ptx::tma_load_2d(sB + st * Cfg::kBBytes + h * kBoxRows * kDim, rmap, &b_full[st], 0, t + h * kBoxRows);
                    const uint32_t ks = st + STAGES * (round & 1);
                    ptx::bulk_load_1d(sKey + ks * TILE_N, keyp + t, TILE_N * 4, &b_full[st]);
                    ptx::bulk_load_1d(sTmin + ks, tminp++, 16, &b_full[st]);
                    t += TILE_N;  // first reference row of the next tile
                    if (--left == 0) more = next_item();
                }
                ++round;
            }
        }
      } else if (x >= 1 && x <= STRIPS) {
        // =========================================================== MMA issuers: warp kEpiWarps + 1 + s feeds strip s
        // One issuer per strip (one elected lane each): the strips' MMAs are no longer issued in a fixed round-robin order
        // (a strip whose accumulator buffer is free never queues behind one that is still being drained), and the issuing
        // instructions are spread over the four SM sub-partitions instead of loading one of them.
        const int s = x - 1;
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_i8(kStripRows, TILE_N, 0, 0);
            // Same loop shape as the loaders: ring position = stage, accumulator buffer (st % TBUFS) and its phase
            // ((st / TBUFS) & 1) are compile-time, descriptors are 32-bit low words (68 -> ~35 instructions per tile).
            const uint32_t b_lo0 = ptx::smem_u32(sB) >> 4;
            const uint32_t a_lo0 = (ptx::smem_u32(sA) + s * kStripRows * kDim) >> 4;
            const uint32_t d_tmem0 = tmem_base + s * TILE_N;
            uint64_t *t_full_s = t_full + s, *t_empty_s = t_empty + s;
            int cur = (int)blockIdx.x - (int)gridDim.x, left = 0;
            uint32_t a = 0, abuf = 0, a_lo = 0, round = 0;
            auto next_item = [&]() -> bool {  // next item with work; waits for its query strips
                for (;;) {
                    int idx;
                    if (DYN) {
                        ISSUER_WAIT(&i_full[a % Cfg::kItemSlots], (a / Cfg::kItemSlots) & 1);
                        idx = s_item[a % Cfg::kItemSlots];
                        if (idx < 0) return false;
                    } else {
                        idx = (cur += (int)gridDim.x);
                        if (idx >= p.n_items) return false;
                    }
                    const WorkItem wi = p.items[idx];
                    const PairDesc pd = p.pairs[wi.pair];
                    if (!DYN && pd.cand_idx >= 0 && wi.row0 >= p.counts[pd.cand_idx]) continue;
                    abuf = a & 1;
                    ISSUER_WAIT(&a_full[abuf], (a >> 1) & 1);
                    a_lo = a_lo0 + abuf * (Cfg::kABytes >> 4);
                    left = (pd.ref_rows + TILE_N - 1) / TILE_N;
                    ++a;
                    return true;
                }
            };
            bool more = next_item();
            while (more) {
#pragma unroll
                for (int st = 0; st < STAGES; ++st) {
                    if (!more) break;
                    constexpr int kBufStride = STRIPS;  // barriers / accumulators are laid out [TBUFS][STRIPS]
                    const int buf = st % TBUFS;
                    long long i0 = 0, i1 = 0, i2 = 0;
                    const bool iprof = DEBUG && (p.debug_flags & 8u) && p.stats != nullptr;
                    if (iprof) i0 = clock64();
                    ISSUER_WAIT(&b_full[st], round & 1);
                    if (iprof) i1 = clock64();
                    ISSUER_WAIT(t_empty_s + buf * kBufStride, ((st / TBUFS) & 1) ^ 1);  // accumulator drained by the epilogue
                    if (iprof) {
                        i2 = clock64();
                        atomicAdd(p.stats + 40 + 2 * s, (unsigned long long)(i1 - i0));  // issuer: waiting for the B tile
                        atomicAdd(p.stats + 41 + 2 * s, (unsigned long long)(i2 - i1));  // issuer: waiting for the accumulator buffer
                    }
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = d_tmem0 + buf * kBufStride * TILE_N;
                    const uint32_t b_lo = b_lo0 + st * (Cfg::kBBytes >> 4);
                    ptx::mma_i8_ss_lo<false>(d_tmem, a_lo, b_lo, idesc);
#pragma unroll
                    for (int k = 1; k < kDim / 32; ++k) ptx::mma_i8_ss_lo<true>(d_tmem, a_lo + k * 2, b_lo + k * 2, idesc);
                    ptx::mma_commit(t_full_s + buf * kBufStride);  // this strip's accumulators are ready
                    ptx::mma_commit(&b_empty[st]);                 // one of STRIPS arrivals that free the B stage
                    if (--left == 0) {
                        ptx::mma_commit(&a_empty[abuf]);           // one of STRIPS arrivals that free the A buffer
                        more = next_item();
                    }
                }
                ++round;
            }
        }
      }
    } else {
        // =========================================================== epilogue (thread = query row x column share)
        ptx::setmaxnreg_inc<(Cfg::kEpiRegs > 232 ? 232 : Cfg::kEpiRegs)>();
        // Scores s = 2*acc - ||r||^2 (maximise; d = ||q||^2 - s).
        // Phase 1 (filter): stream this warp's share of the strip's accumulator tile through registers once and keep
        //   only the maximum RAW accumulator of every group of 8 columns; a group is flagged when that maximum exceeds
        //   T = floor((theta + min||r||^2 over the share) / 2) in any lane.  theta is a lower bound of the row's final
        //   second-best score, so an unflagged group cannot enter the row's top-2
        //   (2*acc - nb_j <= 2*acc - nbmin <= theta) and costs ~0.45 instructions per element.
        // Phase 2 (exact): flagged groups (a few per tile) are re-read from TMEM and scored exactly with packed keys
        //   (branch-free top-2-of-8 network + merge), in ascending column order; ties never displace (strict >), hence
        //   lowest-index tie-breaking inside a share.  Then the accumulator is handed back to the MMA warp.
        // Column shares (CSPLIT = 2): the two threads of a row keep independent top-2 states over their own columns
        //   (merged by the consumer kernels) but publish their two best scores;
        //   theta = max(own S1, partner S1 - 1, min(own S0, partner S0 - 1)).
        //   The "- 1" keeps every element that merely TIES the partner's second best in play, because the partner's
        //   columns may lie to the right of it; elements strictly below two known scores can never reach the top-2.
        const int share = warp / (4 * STRIPS);
        const int strip = (warp >> 2) % STRIPS;
        const int quarter = warp & 3;
        const int row_local = strip * kStripRows + quarter * 32 + lane;
        const uint32_t warp_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + strip * TILE_N + share * Cfg::kColsPerWarp;
        constexpr int kCols = Cfg::kColsPerWarp;
        constexpr int kAbsent = -0x08000000;  // scores below this are pad columns / "no neighbour"
        // everything the tile loop touches in shared memory, as 32-bit shared addresses
        const uint32_t my_slot = ptx::smem_u32(sShare + row_local * CSPLIT + share);
        const uint32_t peer_slot = ptx::smem_u32(sShare + row_local * CSPLIT + (share ^ (CSPLIT - 1)));
        const uint32_t key_base = ptx::smem_u32(sKey) + share * kCols * 4;
        const uint32_t tmin_base = ptx::smem_u32(sTmin);
        const uint32_t t_full_base = ptx::smem_u32(t_full + strip);
        const uint32_t t_empty_base = ptx::smem_u32(t_empty + strip);
        // ring positions are carried incrementally (no divisions in the tile loop)
        // ring positions all derive from ONE running tile counter g (TBUFS and kKeySlots are powers of two)
        static_assert((TBUFS & (TBUFS - 1)) == 0 && (Cfg::kKeySlots & (Cfg::kKeySlots - 1)) == 0, "ring sizes must be powers of two");
        uint32_t g = 0, a = 0;
        const uint32_t i_full_base = ptx::smem_u32(i_full);
        for (int cur = blockIdx.x;; cur += gridDim.x) {
            int item = cur;
            if (DYN) {  // published by the loader one to two items ahead of the MMAs
                ptx::mbar_wait_a(i_full_base + (a % Cfg::kItemSlots) * 8, (a / Cfg::kItemSlots) & 1);
                item = s_item[a % Cfg::kItemSlots];
                if (item < 0) break;
            } else if (item >= p.n_items) {
                break;
            }
            const WorkItem wi = p.items[item];
            const PairDesc pd = p.pairs[wi.pair];
            const int qry_rows = pd.cand_idx >= 0 ? p.counts[pd.cand_idx] : pd.qry_rows;
            if (!DYN && wi.row0 >= qry_rows) continue;  // item past the candidate list (never published when DYN)
            const int q = wi.row0 + row_local;
            const bool valid = q < qry_rows;
            const int na = valid ? ckey_to_norm((pd.cand_idx >= 0 ? p.cand_ckeys : p.ckeys)[pd.qry_off + q]) : 0;
            // rows past the image end hold zeros; park their state where nothing can flag a group
            int S0 = valid ? INT_MIN : 0x20000000, S1 = S0, J0 = -1, J1 = -1;
            // Mutual-check items only need the nearest row, and one row at distance d0 is known to exist (the query
            // row that proposed this candidate): start both slots just below its score so that only rows at least as
            // close are ever scored exactly.  The placeholders carry id -1 and are dropped by the consumers.
            // Collect mode keeps that threshold for the whole row: every reference row at distance <= cand_d0 is listed.
            if (pd.cand_idx >= 0 && valid) S0 = S1 = na - p.cand_d0[pd.qry_off + q] - 1;
            // "Dead row" rule (forward items of msfm_match_pairs only).  The caller keeps a row only if d0/d1 < ratio.  While
            // the running pair violates that with a margin (d0 > rho * d1, rho just above every ratio the caller tests:
            // the row is "dead"), only an element closer than d0 can change the verdict: it becomes the nearest neighbour and
            // the old one — the exact minimum of everything before it, because elements under d0 are never skipped — the
            // second, so the pair is exact again.  Elements between d0 and d1 would only push a dead row further from the
            // threshold, so the pruning threshold of a dead row is its BEST score instead of its second best: about half of
            // the running-top-2 records of such rows (most rows of an image pair) never reach the exact phase.  A row that
            // ends dead reports d1 := d0: "rejected", and still a valid lower bound of its distance to every row but nn0,
            // which is what the mutual check's dangerous-row bound needs (select_candidates_kernel).
            const int rho8 = (MODE == 0 && CSPLIT == 1 && pd.cand_idx < 0) ? (int)p.prune_q8 : (int)kNoPrune;  // shares keep partial states
            const int ntiles = (pd.ref_rows + TILE_N - 1) / TILE_N;
            long long acc_wait = 0, acc_load = 0, acc_p1 = 0, acc_p2 = 0, acc_hot = 0, acc_first = 0;
            int theta = prune_score(S0, S1, na, rho8, kAbsent);  // pruning score of the row, refreshed after every hot tile
            int jtile = share * kCols;  // first column of this warp's share in the current tile
            for (int t = 0; t < ntiles; ++t, jtile += TILE_N) {
                // ---- pull this warp's whole share of the accumulator tile into registers and release TMEM at once
                long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                const bool prof = DEBUG && (p.debug_flags & 8u) && p.stats != nullptr;
                if (prof) c0 = clock64();
                const uint32_t buf = g % TBUFS, t_phase = (g / TBUFS) & 1;
                const uint32_t ks = g % Cfg::kKeySlots;
                ptx::mbar_wait_a(t_full_base + buf * (STRIPS * 8), t_phase);
                ptx::tc_fence_after();
                if (prof) c1 = clock64();
                const uint32_t tile_taddr = warp_taddr + buf * (STRIPS * TILE_N);
                uint32_t acc[kCols / 16][16];
#pragma unroll
                for (int c = 0; c < kCols / 16; ++c) ptx::tmem_ld_32x32b_x16(tile_taddr + c * 16, acc[c]);
                // ---- while the loads fly: column keys of the tile, smallest reference norm, pruning threshold
                // (the tile's keys and minimum norm landed before its MMAs were issued: they share the B stage's barrier)
                const uint32_t ck = key_base + ks * (TILE_N * 4);
                const int nbmin = ptx::lds_s32(tmin_base + ks * 16);  // smallest reference norm of the tile (tile_min_kernel)
                int th = theta;  // dead rows are pruned against their best score, the others against their second best
                if (CSPLIT > 1) {
                    // peer's {S0 - 1, S1 - 1, item tag}: the row's final second best is >= its own S1, >= the peer's S1
                    // and >= min(S0_own, S0_peer); peer scores count minus one (see above)
                    const int4 peer = ptx::lds_v4_volatile(peer_slot);
                    if ((uint32_t)peer.z == a) th = __vimax3_s32(th, peer.y, min(S0, peer.x));
                }
                const int T = (th + nbmin) >> 1;
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_a(t_empty_base + buf * (STRIPS * 8));
                if (prof) c2 = clock64();
                if (DEBUG && (p.debug_flags & 6u)) {  // experiments: 2/4 = TMEM loads + hand-back only
                    if (acc[0][0] == 0x7fffffffu) S1 = 0;
                } else {
                    // ---- phase 1: can any column of this tile still matter to some lane?  (maximum of the raw accumulators
                    //      per group of 8 columns, then of the tile: one vote; most tiles late in an item end here)
                    int gm[kCols / 8];
#pragma unroll
                    for (int gq = 0; gq < kCols / 8; ++gq) {
                        const uint32_t *v = &acc[gq / 2][8 * (gq & 1)];
                        const int m1 = __vimax3_s32((int)v[0], (int)v[1], (int)v[2]);
                        const int m2 = __vimax3_s32((int)v[3], (int)v[4], (int)v[5]);
                        const int m3 = __vimax3_s32((int)v[6], (int)v[7], m1);
                        gm[gq] = max(m2, m3);
                    }
                    int tm = gm[0];
#pragma unroll
                    for (int gq = 1; gq + 1 < kCols / 8; gq += 2) tm = __vimax3_s32(tm, gm[gq], gm[gq + 1]);
                    tm = max(tm, gm[kCols / 8 - 1]);
                    const bool tile_hot = __any_sync(0xFFFFFFFFu, tm > T);
                    if (prof) c3 = clock64();
                    bool touched = false;
                    if (tile_hot) {
                        // ---- which groups?  (one vote and one branch per group: only paid by tiles that have one)
                        bool hot[kCols / 8];
#pragma unroll
                        for (int gq = 0; gq < kCols / 8; ++gq) hot[gq] = __any_sync(0xFFFFFFFFu, gm[gq] > T);
                        // ---- phase 2: exact scoring of the hot groups straight from the registers
#pragma unroll
                        for (int gq = 0; gq < kCols / 8; ++gq) {
                            if (hot[gq] && !(DEBUG && (p.debug_flags & 1u))) {
                                const uint32_t *v = &acc[gq / 2][8 * (gq & 1)];
                                const int4 c0 = ptx::lds_v4(ck + gq * 32);
                                const int4 c1 = ptx::lds_v4(ck + gq * 32 + 16);
                                const int key[8] = {16 * (int)v[0] + c0.x, 16 * (int)v[1] + c0.y, 16 * (int)v[2] + c0.z,
                                                    16 * (int)v[3] + c0.w, 16 * (int)v[4] + c1.x, 16 * (int)v[5] + c1.y,
                                                    16 * (int)v[6] + c1.z, 16 * (int)v[7] + c1.w};
                                const int jb8 = jtile + gq * 8;
                                if (MODE == 1) {
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
                                        if ((key[i] >> 3) > S1) {
                                            const unsigned slot = atomicAdd(p.event_count, 1u);
                                            if (slot < p.event_cap) p.events[slot] = make_int4((int)(pd.qry_off + q), jb8 + i, pd.cand_idx, 0);
                                        }
                                } else {
                                    int g0, g1;
                                    top2_of8(key, g0, g1);
                                    merge_top2(g0 >> 3, jb8 | (g0 & 7), g1 >> 3, jb8 | (g1 & 7), S0, J0, S1, J1);  // columns stay packed: see the item's end
                                }
                                touched = true;
                                if (prof) ++acc_hot;
                            }
                        }
                        // the row's pruning score for the next tiles (only a hot tile can change it)
                        theta = prune_score(S0, S1, na, rho8, kAbsent);
                    }
                    if (CSPLIT > 1 && touched)
                        ptx::sts_v4_volatile(my_slot, max(S0, INT_MIN + 1) - 1, max(S1, INT_MIN + 1) - 1, (int)a, 0);
                }
                if (prof) {  // per-warp cycle accounting of the tile loop (debug flag 8)
                    const long long c4 = clock64();
                    acc_wait += c1 - c0;  // waiting for the accumulator
                    if (t == 0) acc_first = c1 - c0;  // ... of the item's first tile (the item-switch bubble)
                    acc_load += c2 - c1;  // TMEM loads + keys + threshold + hand-back
                    acc_p1 += c3 - c2;    // phase 1
                    acc_p2 += c4 - c3;    // phase 2 + publish
                }
                ++g;
            }
            if (DEBUG && (p.debug_flags & 8u) && p.stats != nullptr && lane == 0) {
                atomicAdd(p.stats + 0, (unsigned long long)acc_hot);
                atomicAdd(p.stats + 6, (unsigned long long)acc_first);
                atomicAdd(p.stats + 7, 1ull);  // warp-items
                atomicAdd(p.stats + 1, (unsigned long long)acc_wait);
                atomicAdd(p.stats + 2, (unsigned long long)acc_load);
                atomicAdd(p.stats + 3, (unsigned long long)acc_p1);
                atomicAdd(p.stats + 4, (unsigned long long)acc_p2);
                atomicAdd(p.stats + 5, (unsigned long long)ntiles);
                atomicAdd(p.stats + 8 + 2 * quarter, (unsigned long long)acc_wait);  // per SM sub-partition
                atomicAdd(p.stats + 9 + 2 * quarter, (unsigned long long)(acc_load + acc_p1 + acc_p2));
            }
            if (valid && MODE == 0) {
                if (prune_score(S0, S1, na, rho8, kAbsent) != S1) S1 = S0;  // ended dead: d1 := d0 (lower bound)
                int4 out;
                out.x = (S0 > kAbsent && J0 >= 0) ? (J0 ^ 7) : -1;  // low three bits hold 7 - (column mod 8), as in the packed keys;
                                                                   // seeded placeholders (mutual twin items) keep id -1
                out.y = (S1 > kAbsent && J1 >= 0) ? (J1 ^ 7) : -1;
                out.z = (S0 > kAbsent) ? na - S0 : INT_MAX;
                out.w = (S1 > kAbsent) ? na - S1 : INT_MAX;
                p.knn[(pd.knn_off + q) * CSPLIT + share] = out;
            }
            ++a;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == Cfg::kEpiWarps + 1) ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace msfm
