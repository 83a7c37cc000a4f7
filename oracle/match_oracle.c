/*
 * match_oracle.c — CPU restatement of MetricSfM's pairwise SIFT-128 matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under metricsfm_b200/ may link, import or
 * call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline — never as the product path.
 *
 * Parity status: PINNED against the reference's own vendored exact engine
 * (nanoflann KD-tree, SfM/src/utils/nanoflann.hpp, compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/) — see
 * tests/test_oracle.py::test_oracle_vs_reference_nanoflann and the committed
 * fixtures under tests/golden/ generated with that binary.  The production
 * FLANN/OpenCV KD-forest (fine_matching_graph.cc:72-99) is approximate,
 * randomized, un-vendored and has no tests upstream: parity against THAT engine
 * is unpinned by construction; what is pinned is the exact 2-NN + ratio
 * arithmetic every reference call site applies to the kNN result.
 *
 * What is restated (reference file:line, relative to /root/reference/SfM/src):
 *   - squared L2 over 128 components, accumulated in index order
 *       utils/nanoflann.hpp:376-383   (L2_Simple_Adaptor::evalMetric)
 *   - sorted 2-slot result set, strict '>' insertion
 *       utils/nanoflann.hpp:116-140   (KNNResultSet::addPoint; with
 *       NANOFLANN_FIRST_MATCH the lowest index wins ties — the rule the
 *       SiftGPU row/col max shaders also use)
 *   - FLANN result layout ids[2q..2q+1], dists[2q..2q+1] (squared)
 *       graph/fine_matching_graph.cc:96-99
 *   - ratio test  dis[0]/dis[1] < th   in fp32, strict
 *       feature/feature_matching.cpp:45-46, :332-337, :491-495
 *       graph/fine_matching_graph.cc:118-131 (0.6 "good" / 0.85 "all")
 *   - too-few-keypoints gate (<20 on either side => no result)
 *       feature/feature_matching.cpp:28-33
 *   - pair orientation / output order
 *       fine_matching_graph.cc:121,127  (ref index first, ascending query m)
 *       feature_matching.cpp:56-64      (query index first, ascending query i)
 *   - mutual best match (row best == col best), lowest index on ties
 *       thirdparty/siftgpu/include/siftgpu/SiftGPU.h:303-308 + GLSL (§2.3)
 *   - descriptor scale feeding the quantiser
 *       feature/feature_extractor_vl_sift.cpp:199-203 (512 x unit norm)
 *       feature/feature_extractor_cuda_sift.cpp:75-80 (unit norm)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_DIM 128

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* q = min(255, max(0, rint(x * scale))): the packer's quantisation rule. */
void oracle_quantize_f32(const float *src, int64_t rows, int64_t src_stride_floats, float scale, uint8_t *dst) {
    for (int64_t r = 0; r < rows; ++r) {
        const float *s = src + r * src_stride_floats;
        uint8_t *d = dst + r * ORACLE_DIM;
        for (int k = 0; k < ORACLE_DIM; ++k) {
            float v = rintf(s[k] * scale);
            if (!(v > 0.0f)) v = 0.0f; /* also maps NaN to 0 */
            if (v > 255.0f) v = 255.0f;
            d[k] = (uint8_t)v;
        }
    }
}

static inline int32_t sqdist_u8(const uint8_t *a, const uint8_t *b) {
    int32_t acc = 0;
    for (int k = 0; k < ORACLE_DIM; ++k) {
        int32_t d = (int32_t)a[k] - (int32_t)b[k];
        acc += d * d;
    }
    return acc;
}

/* fp32 accumulation in index order, exactly nanoflann.hpp:376-383 */
static inline float sqdist_f32(const float *a, const float *b) {
    float result = 0.0f;
    for (int k = 0; k < ORACLE_DIM; ++k) {
        const float diff = a[k] - b[k];
        result += diff * diff;
    }
    return result;
}

/*
 * Exact 2-NN of every query row in the reference set, integer regime.
 *   ref   [M x 128] u8, query [N x 128] u8
 *   ids   [2N]  (nn0, nn1) or -1 when absent
 *   dists [2N]  (float)d0, (float)d1 ; +inf when absent
 * Lowest reference index wins ties for both neighbours.
 */
void oracle_knn2_u8(const uint8_t *ref, int32_t M, const uint8_t *query, int32_t N, int32_t *ids, float *dists) {
#pragma omp parallel for schedule(static)
    for (int32_t q = 0; q < N; ++q) {
        const uint8_t *qa = query + (int64_t)q * ORACLE_DIM;
        int64_t d0 = INT64_MAX, d1 = INT64_MAX;
        int32_t i0 = -1, i1 = -1;
        for (int32_t j = 0; j < M; ++j) {
            int64_t d = sqdist_u8(qa, ref + (int64_t)j * ORACLE_DIM);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        ids[2 * q] = i0;
        ids[2 * q + 1] = i1;
        dists[2 * q] = i0 >= 0 ? (float)d0 : INFINITY;
        dists[2 * q + 1] = i1 >= 0 ? (float)d1 : INFINITY;
    }
}

/* Same on float descriptors (the reference's native container: CV_32FC1 N x 128). */
void oracle_knn2_f32(const float *ref, int32_t M, const float *query, int32_t N, int32_t *ids, float *dists) {
#pragma omp parallel for schedule(static)
    for (int32_t q = 0; q < N; ++q) {
        const float *qa = query + (int64_t)q * ORACLE_DIM;
        float d0 = INFINITY, d1 = INFINITY;
        int32_t i0 = -1, i1 = -1;
        for (int32_t j = 0; j < M; ++j) {
            float d = sqdist_f32(qa, ref + (int64_t)j * ORACLE_DIM);
            if (i0 < 0 || d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (i1 < 0 || d < d1) { d1 = d; i1 = j; }
        }
        ids[2 * q] = i0;
        ids[2 * q + 1] = i1;
        dists[2 * q] = i0 >= 0 ? d0 : INFINITY;
        dists[2 * q + 1] = i1 >= 0 ? d1 : INFINITY;
    }
}

/* For every reference row j: the query row with the smallest distance (lowest q on ties). */
void oracle_colbest_u8(const uint8_t *ref, int32_t M, const uint8_t *query, int32_t N, int32_t *col_best, float *col_dist) {
#pragma omp parallel for schedule(static)
    for (int32_t j = 0; j < M; ++j) {
        const uint8_t *rb = ref + (int64_t)j * ORACLE_DIM;
        int64_t best = INT64_MAX;
        int32_t bi = -1;
        for (int32_t q = 0; q < N; ++q) {
            int64_t d = sqdist_u8(query + (int64_t)q * ORACLE_DIM, rb);
            if (d < best) { best = d; bi = q; }
        }
        col_best[j] = bi;
        if (col_dist) col_dist[j] = bi >= 0 ? (float)best : INFINITY;
    }
}

void oracle_colbest_f32(const float *ref, int32_t M, const float *query, int32_t N, int32_t *col_best, float *col_dist) {
#pragma omp parallel for schedule(static)
    for (int32_t j = 0; j < M; ++j) {
        const float *rb = ref + (int64_t)j * ORACLE_DIM;
        float best = INFINITY;
        int32_t bi = -1;
        for (int32_t q = 0; q < N; ++q) {
            float d = sqdist_f32(query + (int64_t)q * ORACLE_DIM, rb);
            if (bi < 0 || d < best) { best = d; bi = q; }
        }
        col_best[j] = bi;
        if (col_dist) col_dist[j] = bi >= 0 ? best : INFINITY;
    }
}

/*
 * Ratio test (+ optional max-distance gate, optional mutual check) over a FLANN-layout
 * kNN result, ascending query index.
 *   orientation 0: emit (nn0, q)  — fine_matching_graph.cc:121,127 (ref image first)
 *   orientation 1: emit (q, nn0)  — feature_matching.cpp:60-61     (query image first)
 *   col_best: NULL disables the mutual check.
 *   max_dist_sq <= 0 disables the distance gate (SiftGPU distmax analogue on squared L2).
 *   ratio_good > 0: good_flags[k] = 1 iff the same match also passes ratio_good
 *                   (the dual 0.6/0.85 lists of fine_matching_graph.cc:118-131).
 *   reject_gt: 0 = accept iff d0/d1 < ratio (feature_matching.cpp:45-46, fine_matching_graph.cc:118-129);
 *              1 = SLAMGPS::FeatureMatching's rule, slam_gps.cc:470-477: `if (ratio > th) continue;`, i.e. accept iff
 *                  !(d0/d1 > ratio) — non-strict, and 0/0 = NaN is accepted.
 * Returns the number of matches, or -1 when the <min_keypoints gate rejects the pair
 * (feature_matching.cpp:30-33 returns false).
 */
static int ratio_pass(float r, float th, int32_t reject_gt) { return reject_gt ? !(r > th) : (r < th); }

int32_t oracle_ratio_select_rule(const int32_t *ids, const float *dists, int32_t M, int32_t N, float ratio, float max_dist_sq,
                                 const int32_t *col_best, int32_t min_keypoints, int32_t orientation, float ratio_good,
                                 int32_t reject_gt, int32_t *out_pairs /* [N][2] */, uint8_t *good_flags /* [N] or NULL */) {
    if (M < min_keypoints || N < min_keypoints) return -1;
    int32_t n = 0;
    for (int32_t q = 0; q < N; ++q) {
        const int32_t i0 = ids[2 * q], i1 = ids[2 * q + 1];
        if (i0 < 0 || i1 < 0) continue; /* fewer than two reference points: no ratio exists */
        const float d0 = dists[2 * q], d1 = dists[2 * q + 1];
        const float r = d0 / d1; /* IEEE fp32 divide; 0/0 = NaN compares false */
        if (!ratio_pass(r, ratio, reject_gt)) continue;
        if (max_dist_sq > 0.0f && !(d0 < max_dist_sq)) continue;
        if (col_best && col_best[i0] != q) continue;
        if (orientation == 0) { out_pairs[2 * n] = i0; out_pairs[2 * n + 1] = q; }
        else { out_pairs[2 * n] = q; out_pairs[2 * n + 1] = i0; }
        if (good_flags) good_flags[n] = (ratio_good > 0.0f && ratio_pass(r, ratio_good, reject_gt)) ? 1 : 0;
        ++n;
    }
    return n;
}

int32_t oracle_ratio_select(const int32_t *ids, const float *dists, int32_t M, int32_t N, float ratio, float max_dist_sq,
                            const int32_t *col_best, int32_t min_keypoints, int32_t orientation, float ratio_good,
                            int32_t *out_pairs /* [N][2] */, uint8_t *good_flags /* [N] or NULL */) {
    return oracle_ratio_select_rule(ids, dists, M, N, ratio, max_dist_sq, col_best, min_keypoints, orientation, ratio_good, 0,
                                    out_pairs, good_flags);
}

/*
 * Whole per-pair path on u8 descriptors: 2-NN, optional column best, ratio select.
 * Scratch-free convenience used by the tests and the CPU baseline.
 */
int32_t oracle_match_pair_u8(const uint8_t *ref, int32_t M, const uint8_t *query, int32_t N, float ratio, float max_dist_sq,
                             int32_t mutual, int32_t min_keypoints, int32_t orientation, float ratio_good, int32_t *ids,
                             float *dists, int32_t *out_pairs, uint8_t *good_flags) {
    if (M < min_keypoints || N < min_keypoints) return -1;
    oracle_knn2_u8(ref, M, query, N, ids, dists);
    int32_t *col_best = NULL;
    if (mutual) {
        col_best = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
        oracle_colbest_u8(ref, M, query, N, col_best, NULL);
    }
    int32_t n = oracle_ratio_select(ids, dists, M, N, ratio, max_dist_sq, col_best, min_keypoints, orientation, ratio_good,
                                    out_pairs, good_flags);
    free(col_best);
    return n;
}
