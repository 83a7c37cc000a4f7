// ref_nanoflann_harness.cc — thin driver around the REFERENCE's own vendored exact kNN engine.
//
// TEST INFRASTRUCTURE ONLY (see oracle/match_oracle.c header).  This file contains no reference
// source: it #includes the reference headers where they lie (-I/root/reference/SfM/src, set by
// oracle/Makefile) and is compiled into oracle/_ref/libref_nanoflann.so, which is git-ignored.
//
// It reproduces the one matcher overload of the reference that is exact and fully in-tree:
//   FeatureMatching::KNNMatchingWithGeoVerify(kp1, my_kd_tree_t* kd_tree1, kp2, descriptors2, matches)
//     SfM/src/feature/feature_matching.cpp:319-342  — for each row i of image 2:
//     kd_tree1->knnSearch(row, 2, id, dis); ratio = dis[0]/dis[1]; ratio < 0.5 -> match (id[0], i)
//   my_kd_tree_t = KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<float, SiftList<float>>, SiftList<float>, 128>
//     SfM/src/basic_structs.h:260 ; SiftList in SfM/src/utils/nanoflann_utils.h:37-80
// The query loop is threaded with OpenMP the way the production caller threads its partner loop
// (graph/fine_matching_graph.cc:87); flags follow SfM/CMakeLists.txt:17-19 (-O3 -march=native).
#include <cstdint>
#include <cstddef>
#include <stdexcept>
#include <vector>
#include <cmath>
#include <limits>

#include "utils/nanoflann.hpp"
#include "utils/nanoflann_utils.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
typedef objectsfm::SiftList<float> sift_list_t;
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, sift_list_t>, sift_list_t, 128> kd_tree_t;

struct RefIndex {
    sift_list_t list;
    kd_tree_t *tree;
};
}  // namespace

extern "C" {

int ref_nanoflann_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// The reference sizes its OpenMP team from the machine (no omp_set_num_threads anywhere under SfM/src); launchers such as
// torchrun export OMP_NUM_THREADS=1, so the benchmark arm sets the team size explicitly.
void ref_nanoflann_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// Build the KD-tree on image 1's descriptors (rows x 128 float, contiguous).
void *ref_nanoflann_build(const float *desc1, int32_t rows) {
    RefIndex *ix = new RefIndex();
    ix->list.pts.reserve(rows);
    for (int32_t i = 0; i < rows; ++i)
        ix->list.pts.push_back(sift_list_t::SiftData(const_cast<float *>(desc1) + (size_t)i * 128, i));
    ix->tree = new kd_tree_t(128, ix->list, nanoflann::KDTreeSingleIndexAdaptorParams(10));
    ix->tree->buildIndex();
    return ix;
}

void ref_nanoflann_free(void *handle) {
    RefIndex *ix = static_cast<RefIndex *>(handle);
    if (!ix) return;
    delete ix->tree;
    delete ix;
}

// 2-NN of every row of image 2 in image 1's tree; FLANN layout out (ids[2i..], squared dists[2i..]).
void ref_nanoflann_knn2(void *handle, const float *desc2, int32_t rows2, int32_t *ids, float *dists) {
    RefIndex *ix = static_cast<RefIndex *>(handle);
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t i = 0; i < rows2; ++i) {
        size_t id[2] = {(size_t)-1, (size_t)-1};
        float dis[2] = {std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity()};
        size_t found = ix->tree->knnSearch(desc2 + (size_t)i * 128, 2, id, dis);
        ids[2 * i] = found > 0 ? (int32_t)id[0] : -1;
        ids[2 * i + 1] = found > 1 ? (int32_t)id[1] : -1;
        dists[2 * i] = found > 0 ? dis[0] : std::numeric_limits<float>::infinity();
        dists[2 * i + 1] = found > 1 ? dis[1] : std::numeric_limits<float>::infinity();
    }
}

// The whole overload feature_matching.cpp:319-342: gate, kNN, ratio, (id1, id2) list ascending id2.
// Returns the number of matches or -1 when the gate rejects (reference returns false).
int32_t ref_nanoflann_match(const float *desc1, int32_t rows1, const float *desc2, int32_t rows2, float th_ratio,
                            int32_t th_reject, int32_t *out_pairs /* [rows2][2] */) {
    if (rows1 < th_reject || rows2 < th_reject) return -1;
    void *h = ref_nanoflann_build(desc1, rows1);
    std::vector<int32_t> ids((size_t)rows2 * 2);
    std::vector<float> dis((size_t)rows2 * 2);
    ref_nanoflann_knn2(h, desc2, rows2, ids.data(), dis.data());
    ref_nanoflann_free(h);
    int32_t n = 0;
    for (int32_t i = 0; i < rows2; ++i) {
        float ratio = dis[2 * i] / dis[2 * i + 1];
        if (ratio < th_ratio) {
            out_pairs[2 * n] = ids[2 * i];
            out_pairs[2 * n + 1] = i;
            ++n;
        }
    }
    return n;
}

}  // extern "C"
