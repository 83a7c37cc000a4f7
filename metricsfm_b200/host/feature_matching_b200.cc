// feature_matching_b200.cc — see feature_matching_b200.h.  Pure marshalling over the C ABI.
#include "feature_matching_b200.h"

#include <cstring>

namespace objectsfm {

namespace {
bool upload_mat(msfm_ctx *ctx, int slot, cv::Mat &d, float scale, std::string &err) {
    msfm_status st;
    if (d.rows > 0 && d.cols != MSFM_DIM) {
        err = "descriptors must have 128 columns";
        return false;
    }
    if (d.type() == CV_32FC1)
        st = msfm_upload_f32(ctx, slot, d.ptr<float>(0), d.rows, d.rows ? (long long)(d.step / sizeof(float)) : MSFM_DIM, scale);
    else if (d.type() == CV_8UC1)
        st = msfm_upload_u8(ctx, slot, d.ptr<unsigned char>(0), d.rows, d.rows ? (long long)d.step : MSFM_DIM);
    else {
        err = "descriptors must be CV_32FC1 or CV_8UC1";
        return false;
    }
    if (st != MSFM_OK) err = msfm_last_error(ctx);
    return st == MSFM_OK;
}
}  // namespace

FeatureMatchingB200::FeatureMatchingB200(const MatcherB200Options &opt) : opt_(opt) {
    msfm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = opt.device;
    cfg.max_images = 2;
    cfg.arena_rows = 2 * 1000192;  // two images of up to idx_max_per_image rows (basic_structs.h:171)
    msfm_status st = msfm_create(&cfg, &ctx_);
    if (st != MSFM_OK) {
        ctx_ = nullptr;
        err_ = msfm_status_string(st);
    }
}

FeatureMatchingB200::~FeatureMatchingB200() { msfm_destroy(ctx_); }

bool FeatureMatchingB200::Upload(int slot, cv::Mat &d) { return upload_mat(ctx_, slot, d, opt_.descriptor_scale, err_); }

bool FeatureMatchingB200::Match(cv::Mat &d1, cv::Mat &d2, bool mutual, std::vector<std::pair<int, int>> &matches) {
    if (!ctx_) return false;
    msfm_release_all(ctx_);
    if (!Upload(0, d1) || !Upload(1, d2)) return false;
    // index on image 2, queries = rows of image 1 (feature_matching.cpp:35-44)
    msfm_pair pair = {1, 0};
    msfm_params prm;
    prm.ratio = opt_.th_ratio;
    prm.ratio_good = 0.f;
    prm.max_dist_sq = 0.f;
    prm.mutual = mutual ? 1 : 0;
    prm.min_keypoints = opt_.th_reject;
    prm.orientation = 1;  // (i1, i2), ascending i1 (feature_matching.cpp:56-64)
    prm.rescore_band = 0.f;
    std::vector<int32_t> buf((size_t)(d1.rows > 0 ? d1.rows : 1) * 2);
    int64_t offsets[2] = {0, 0};
    int32_t okflag = 0;
    msfm_result res;
    res.offsets = offsets;
    res.ok = &okflag;
    res.matches = reinterpret_cast<int32_t(*)[2]>(buf.data());
    res.good = nullptr;
    res.match_capacity = d1.rows;
    msfm_status st = msfm_match_pairs(ctx_, &pair, 1, &prm, &res);
    if (st != MSFM_OK) {
        err_ = msfm_last_error(ctx_);
        return false;
    }
    if (!okflag) return false;  // th_reject gate (feature_matching.cpp:30-33)
    matches.resize((size_t)offsets[1]);  // the reference resizes (feature_matching.cpp:54)
    for (size_t k = 0; k < matches.size(); ++k) matches[k] = std::pair<int, int>(buf[2 * k], buf[2 * k + 1]);
    return true;
}

bool FeatureMatchingB200::KNNMatching(std::vector<cv::KeyPoint> &kp1, cv::Mat &descriptors1, std::vector<cv::KeyPoint> &kp2,
                                      cv::Mat &descriptors2, std::vector<std::pair<int, int>> &matches) {
    if ((int)kp1.size() < opt_.th_reject || (int)kp2.size() < opt_.th_reject) return false;
    return Match(descriptors1, descriptors2, opt_.mutual, matches);
}

bool FeatureMatchingB200::Run(std::vector<cv::KeyPoint> &kp1, cv::Mat &descriptors1, std::vector<cv::KeyPoint> &kp2,
                              cv::Mat &descriptors2, std::vector<std::pair<int, int>> &matches) {
    if ((int)kp1.size() < opt_.th_reject || (int)kp2.size() < opt_.th_reject) return false;
    return Match(descriptors1, descriptors2, true, matches);
}

bool FeatureMatchingB200::KNN2(cv::Mat &descriptors1, cv::Mat &descriptors2, int *id, float *dis) {
    if (!ctx_) return false;
    msfm_release_all(ctx_);
    if (!Upload(0, descriptors1) || !Upload(1, descriptors2)) return false;
    msfm_status st = msfm_knn2(ctx_, 0, 1, id, dis);
    if (st != MSFM_OK) err_ = msfm_last_error(ctx_);
    return st == MSFM_OK;
}

MatchGraphB200::MatchGraphB200(int device, int max_images, long long total_rows, float descriptor_scale) : scale_(descriptor_scale) {
    msfm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = device;
    cfg.max_images = max_images;
    cfg.arena_rows = total_rows + 256ll * max_images;  // every image is padded to a multiple of 256 rows
    rows_.assign(max_images, -1);
    msfm_status st = msfm_create(&cfg, &ctx_);
    if (st != MSFM_OK) {
        ctx_ = nullptr;
        err_ = msfm_status_string(st);
    }
}

MatchGraphB200::~MatchGraphB200() { msfm_destroy(ctx_); }

bool MatchGraphB200::AddImage(int idx, cv::Mat &descriptors) {
    if (!ctx_ || idx < 0 || idx >= (int)rows_.size()) return false;
    if (!upload_mat(ctx_, idx, descriptors, scale_, err_)) return false;
    rows_[idx] = descriptors.rows;
    return true;
}

bool MatchGraphB200::ReleaseImage(int idx) {
    if (!ctx_ || idx < 0 || idx >= (int)rows_.size()) return false;
    rows_[idx] = -1;
    return msfm_release(ctx_, idx) == MSFM_OK;
}

bool MatchGraphB200::MatchPairs(const std::vector<std::vector<int>> &match_graph_init, std::vector<std::vector<PairMatches>> &out,
                                float th_good, float th_all, bool mutual, int th_reject) {
    if (!ctx_) return false;
    std::vector<msfm_pair> pairs;
    long long capacity = 0;
    for (size_t idx1 = 0; idx1 < match_graph_init.size(); ++idx1)
        for (int idx2 : match_graph_init[idx1]) {
            if (idx2 < 0 || idx2 >= (int)rows_.size() || rows_[idx1] < 0 || rows_[idx2] < 0) {
                err_ = "pair list names an image that was not added";
                return false;
            }
            msfm_pair p = {(int32_t)idx1, (int32_t)idx2};  // index on idx1, queries = rows of idx2 (fine_matching_graph.cc:81,99)
            pairs.push_back(p);
            capacity += rows_[idx2];
        }
    std::vector<int64_t> offsets(pairs.size() + 1);
    std::vector<int32_t> okflags(pairs.size());
    std::vector<int32_t> buf((size_t)(capacity > 0 ? capacity : 1) * 2);
    std::vector<uint8_t> good((size_t)(capacity > 0 ? capacity : 1));
    msfm_params prm;
    prm.ratio = th_all;
    prm.ratio_good = th_good;
    prm.max_dist_sq = 0.f;
    prm.mutual = mutual ? 1 : 0;
    prm.min_keypoints = th_reject;
    prm.orientation = 0;  // (ptid1, ptid2) ascending ptid2 (fine_matching_graph.cc:121,127)
    prm.rescore_band = 0.f;
    msfm_result res;
    res.offsets = offsets.data();
    res.ok = okflags.data();
    res.matches = reinterpret_cast<int32_t(*)[2]>(buf.data());
    res.good = good.data();
    res.match_capacity = capacity;
    msfm_status st = msfm_match_pairs(ctx_, pairs.data(), (int64_t)pairs.size(), &prm, &res);
    if (st != MSFM_OK) {
        err_ = msfm_last_error(ctx_);
        return false;
    }
    out.assign(match_graph_init.size(), std::vector<PairMatches>());
    size_t p = 0;
    for (size_t idx1 = 0; idx1 < match_graph_init.size(); ++idx1) {
        out[idx1].resize(match_graph_init[idx1].size());
        for (size_t j = 0; j < match_graph_init[idx1].size(); ++j, ++p) {
            PairMatches &pm = out[idx1][j];
            pm.ok = okflags[p] != 0;
            for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k) {
                pm.matches_all.push_back(std::pair<int, int>(buf[2 * k], buf[2 * k + 1]));
                pm.is_good.push_back(good[k]);
            }
        }
    }
    return true;
}

}  // namespace objectsfm
