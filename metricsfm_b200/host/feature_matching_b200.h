// feature_matching_b200.h — C++ host mirror of MetricSfM's matcher interface on top of the C ABI (include/msfm_match.h).
//
// Same containers in (std::vector<cv::KeyPoint>, cv::Mat CV_32FC1 N x 128), same per-pair match lists out
// (std::vector<std::pair<int,int>>), same bool error convention, so the SfM tracking / triangulation stages
// (sfm_incremental.cc:224-915, slam_gps.cc:557-668) consume the results unchanged:
//
//   reference                                                             here
//   FeatureMatching::KNNMatching(kp1, d1, kp2, d2, matches)               FeatureMatchingB200::KNNMatching
//     SfM/src/feature/feature_matching.h:33-35, .cpp:24-65                  (index on image 2, ratio 0.5, (i1,i2) ascending i1)
//   FeatureMatching::KNNMatchingWithGeoVerify(kp1, kp2, id, dis, matches) FeatureMatchingB200::KNN2 fills id/dis
//     feature_matching.h:57-58, .cpp:477-501                                ([2*N2] FLANN layout, squared L2)
//   FeatureMatchingCudaSift::Run(kp1, d1, kp2, d2, matches)               FeatureMatchingB200::Run
//     feature_matching_cuda_sift.h:34-36, .cpp:21-108 (kNN part)            (mutual best match, ratio as given)
//   FineMatchingGraph::BuildMatchGraph kNN + ratio loops                  MatchGraphB200::{AddImage, MatchPairs}
//     graph/fine_matching_graph.cc:58-133                                    (descriptors staged once, whole pair list batched)
//
// Everything here is a thin marshalling layer: all arithmetic runs in the CUDA library behind msfm_match.h.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "../../include/msfm_match.h"
#include "cv_standin.h"

namespace objectsfm {

// Float descriptors are quantised q = min(255, max(0, rint(x * scale))): 1 for VLSIFT's 512-scaled rows
// (feature_extractor_vl_sift.cpp:201-203), 512 for unit-norm rows (feature_extractor_cuda_sift.cpp:75-80).
struct MatcherB200Options {
    int device = 0;
    float descriptor_scale = 1.0f;
    float th_ratio = 0.5f;   // feature_matching.cpp:27
    int th_reject = 20;      // feature_matching.cpp:28
    bool mutual = false;     // the CPU paths have none; SiftMatchGPU::GetSiftMatch defaults to 1 (SiftGPU.h:308)
};

class FeatureMatchingB200 {
public:
    explicit FeatureMatchingB200(const MatcherB200Options &opt = MatcherB200Options());
    ~FeatureMatchingB200();
    FeatureMatchingB200(const FeatureMatchingB200 &) = delete;
    FeatureMatchingB200 &operator=(const FeatureMatchingB200 &) = delete;

    bool ok() const { return ctx_ != nullptr; }
    const std::string &last_error() const { return err_; }

    // FeatureMatching::KNNMatching: kd-index on descriptors2, every row of descriptors1 queried, ratio < th_ratio,
    // matches resized to (i1, i2) ascending i1; false when either image has < th_reject keypoints.
    bool KNNMatching(std::vector<cv::KeyPoint> &kp1, cv::Mat &descriptors1, std::vector<cv::KeyPoint> &kp2, cv::Mat &descriptors2,
                     std::vector<std::pair<int, int>> &matches);
    // FeatureMatchingCudaSift::Run kNN part: same as above with the mutual-best-match rule of the declared GPU matchers.
    bool Run(std::vector<cv::KeyPoint> &kp1, cv::Mat &descriptors1, std::vector<cv::KeyPoint> &kp2, cv::Mat &descriptors2,
             std::vector<std::pair<int, int>> &matches);
    // FeatureMatching::KNNMatchingWithGeoVerify(kp1, d1, kp2, d2, matches) (feature_matching.cpp:67-150): ratio matches as in
    // KNNMatching, then two verification passes (epipolar threshold 3 px, then 1 px): fewer than th_reject matches =>
    // false; a least-squares homography close to the identity (diagonal within 0.01 of 0.995) => false (no parallax);
    // RANSAC-F, outliers dropped.  The matches are APPENDED (the reference push_backs).  RANSAC runs on the GPU
    // (msfm_geo_ransac); the homography is a normalised DLT without OpenCV's final Levenberg-Marquardt polish, which the
    // 0.01 test does not resolve.
    bool KNNMatchingWithGeoVerify(std::vector<cv::KeyPoint> &kp1, cv::Mat &descriptors1, std::vector<cv::KeyPoint> &kp2,
                                  cv::Mat &descriptors2, std::vector<std::pair<int, int>> &matches);
    // Index on image 1, 2-NN of every row of image 2 in FLANN layout: id[2*N2], dis[2*N2] (squared L2) — exactly the
    // arrays KNNMatchingWithGeoVerify(kp1, kp2, id, dis, matches) and fine_matching_graph.cc:96-99 consume.
    bool KNN2(cv::Mat &descriptors1, cv::Mat &descriptors2, int *id, float *dis);

private:
    bool Match(cv::Mat &d1, cv::Mat &d2, bool mutual, std::vector<std::pair<int, int>> &matches);
    bool Upload(int slot, cv::Mat &d);
    msfm_ctx *ctx_ = nullptr;
    MatcherB200Options opt_;
    std::string err_;
};

// Least-squares homography pt2 ~ H pt1 (normalised DLT, smallest eigenvector of A^T A by Jacobi sweeps), scaled so that
// H[8] = 1; false when the system is degenerate.  Stand-in for cv::findHomography(pt1, pt2) with method 0
// (feature_matching.cpp:115).
bool FindHomographyDLT(const std::vector<cv::Point2f> &pt1, const std::vector<cv::Point2f> &pt2, double H[9]);

// Batched form of FineMatchingGraph::BuildMatchGraph's matching loops: stage every image once (replaces the per-idx1
// flann_build_index and the per-pair disk re-reads, fine_matching_graph.cc:69-91), then match the whole candidate pair
// list; per pair the "all" list (ratio < th_all) with a flag for the "good" subset (ratio < th_good), pairs
// (id_in_idx1, id_in_idx2) ascending id_in_idx2 as in fine_matching_graph.cc:116-133.
class MatchGraphB200 {
public:
    MatchGraphB200(int device, int max_images, long long total_rows, float descriptor_scale = 1.0f);
    ~MatchGraphB200();
    MatchGraphB200(const MatchGraphB200 &) = delete;
    MatchGraphB200 &operator=(const MatchGraphB200 &) = delete;

    bool ok() const { return ctx_ != nullptr; }
    const std::string &last_error() const { return err_; }
    bool AddImage(int idx, cv::Mat &descriptors);   // once per image
    bool ReleaseImage(int idx);

    struct PairMatches {
        bool ok = false;                              // false: < 20 keypoints on either side
        std::vector<std::pair<int, int>> matches_all;  // ratio < th_all
        std::vector<unsigned char> is_good;            // 1 iff also ratio < th_good
    };
    // match_graph_init[idx1] = partner list, as produced by InitialMatchingGraph (initial_matching_graph.h:64).
    bool MatchPairs(const std::vector<std::vector<int>> &match_graph_init, std::vector<std::vector<PairMatches>> &out,
                    float th_good = 0.6f, float th_all = 0.85f, bool mutual = false, int th_reject = 20);

private:
    msfm_ctx *ctx_ = nullptr;
    float scale_;
    std::vector<int> rows_;
    std::string err_;
};

}  // namespace objectsfm
