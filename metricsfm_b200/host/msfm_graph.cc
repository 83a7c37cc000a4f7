// msfm_graph.cc — FineMatchingGraph::BuildMatchGraph rebuilt around the GPU matcher (include/msfm_graph.h).
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/msfm_graph.h"
#include "../../include/msfm_match.h"
#include "../../include/msfm_multi.h"
#include "../../include/msfm_store.h"

namespace {

int fail(char *err, size_t cap, int code, const char *fmt, ...) {
    if (err && cap) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, cap, fmt, ap);
        va_end(ap);
    }
    return code;
}

// MSFM_GRAPH_VERBOSE=1: wall-clock of every stage on stderr.
struct StageClock {
    bool on = getenv("MSFM_GRAPH_VERBOSE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        const auto n = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[msfm graph] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

struct Image {
    msfm_feature_info info{};
    std::vector<float> xy;  // centred keypoints
    bool needed = false, loaded = false;
};

}  // namespace

extern "C" int msfm_build_match_graph(const char *fold, int32_t num_imgs, const int64_t *offsets, const int32_t *list,
                                      const msfm_graph_options *opt, msfm_verify_fn verify, void *user, char *err, size_t err_cap) {
    if (!fold || num_imgs < 0 || !offsets || !opt) return fail(err, err_cap, -1, "null argument");
    if (err && err_cap) err[0] = '\0';

    // ---- resume: which images still have to be matched (fine_matching_graph.cc:49-54)
    std::vector<int32_t> missing((size_t)num_imgs);
    int32_t n_missing = 0;
    if (msfm_match_index_missing(fold, num_imgs, missing.data(), &n_missing) != 0) return fail(err, err_cap, -2, "match_index.txt unreadable");
    if (n_missing == 0) return 0;
    missing.resize(n_missing);
    std::vector<char> is_missing((size_t)num_imgs, 0);
    for (int32_t i : missing) is_missing[i] = 1;
    std::vector<int32_t> existing;
    for (int32_t i = 0; i < num_imgs; ++i)
        if (!is_missing[i]) existing.push_back(i);
    std::vector<int32_t> graph((size_t)num_imgs * num_imgs, 0);
    if (msfm_graph_recover(fold, num_imgs, existing.data(), (int32_t)existing.size(), graph.data()) != 0)
        return fail(err, err_cap, -2, "existing match files unreadable");

    StageClock clk;
    clk.lap("resume state");
    // ---- the images this run touches, their feature headers
    std::vector<Image> imgs((size_t)num_imgs);
    bool any_pair = false;
    for (int32_t idx1 : missing)
        for (int64_t j = offsets[idx1]; j < offsets[idx1 + 1]; ++j) {
            const int32_t idx2 = list[j];
            if (idx2 < 0 || idx2 >= num_imgs) return fail(err, err_cap, -1, "partner index %d of image %d out of range", idx2, idx1);
            imgs[idx1].needed = imgs[idx2].needed = true;
            any_pair = true;
        }
    int64_t arena_rows = 0;
    char path[4096];
    for (int32_t i = 0; i < num_imgs; ++i) {
        if (!imgs[i].needed) continue;
        if (msfm_feature_path(fold, i, path, sizeof path) != 0 || msfm_feature_stat(path, &imgs[i].info) != 0)
            return fail(err, err_cap, -2, "feature file of image %d missing or malformed", i);
        const msfm_feature_info &fi = imgs[i].info;
        if (fi.desc_rows > 0 && (fi.desc_cols != 128 || (fi.desc_type != 5 && fi.desc_type != 0)))
            return fail(err, err_cap, -3, "image %d: descriptors must be N x 128 CV_32FC1 or CV_8UC1 (got cols %d type %d)", i, fi.desc_cols, fi.desc_type);
        if (fi.desc_rows != fi.num_pts) return fail(err, err_cap, -3, "image %d: %d keypoints but %d descriptor rows", i, fi.num_pts, fi.desc_rows);
        arena_rows += (fi.desc_rows + 255) / 256 * 256 + 256;
    }
    clk.lap("feature headers");

    // ---- stage every needed image in HBM once (replaces the per-idx1 flann_build_index and the per-pair re-reads):
    //      feature files are read into page-locked staging, one chunk while the previous chunk's host->device copies run
    const int n_dev = opt->n_devices > 1 ? opt->n_devices : 1;
    msfm_ctx *ctx = nullptr;
    msfm_multi *mm = nullptr;
    if (any_pair) {
        // each row costs 132 B packed (+ 512 B when the float rows are retained for re-scoring)
        const int64_t need = arena_rows * (132 + (opt->rescore_band > 0.0f ? 512 : 0));
        int64_t free_b = 0, total_b = 0;
        if (msfm_device_memory(opt->device, &free_b, &total_b) == MSFM_OK && need > free_b)
            return fail(err, err_cap, -4, "the descriptor table of this run needs %.1f GB of HBM (%lld rows), device %d has %.1f GB free: "
                                          "match the collection in several runs over subsets of match_graph_init (resume keeps what is done)",
                        need / 1e9, (long long)arena_rows, opt->device, free_b / 1e9);
        if (n_dev > 1) {
            if (opt->rescore_band > 0.0f) return fail(err, err_cap, -1, "rescore_band (retained float rows) is single-GPU only");
            msfm_multi_config mc;
            memset(&mc, 0, sizeof mc);
            mc.n_devices = n_dev;
            mc.max_images = num_imgs;
            mc.arena_rows = arena_rows > 0 ? arena_rows : 256;
            msfm_status st = msfm_multi_create(&mc, &mm);
            if (st != MSFM_OK) return fail(err, err_cap, -4, "msfm_multi_create: %s", msfm_status_string(st));
            ctx = msfm_multi_context(mm, 0);  // geo-verification runs on the first device
        } else {
            msfm_config cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.device = opt->device;
            cfg.max_images = num_imgs;
            cfg.arena_rows = arena_rows > 0 ? arena_rows : 256;
            cfg.keep_float = opt->rescore_band > 0.0f ? 1 : 0;
            msfm_status st = msfm_create(&cfg, &ctx);
            if (st != MSFM_OK) return fail(err, err_cap, -4, "msfm_create: %s", msfm_status_string(st));
        }
    }
    clk.lap("create context(s)");
    void *stage[2] = {nullptr, nullptr};
    auto bail = [&](int code, const std::string &msg) {
        if (mm) msfm_multi_destroy(mm);
        else if (ctx) msfm_destroy(ctx);
        msfm_host_free(stage[0]);
        msfm_host_free(stage[1]);
        return fail(err, err_cap, code, "%s", msg.c_str());
    };
    auto sync_uploads = [&]() { return mm ? msfm_multi_sync(mm) : msfm_sync(ctx); };
    if (ctx) {
        const size_t kStageBytes = (size_t)256 << 20;
        size_t largest = 0;
        for (int32_t i = 0; i < num_imgs; ++i)
            if (imgs[i].needed) largest = std::max(largest, (size_t)imgs[i].info.desc_rows * 128 * imgs[i].info.desc_elem_size);
        const size_t cap = std::max(kStageBytes, largest + 16);
        for (void *&p : stage)
            if (msfm_host_alloc(cap, &p) != MSFM_OK) return bail(-4, "cannot allocate page-locked staging memory");
        clk.lap("page-locked staging buffers");
        double read_ms = 0.0, wait_ms = 0.0;
        int which = 0;
        int32_t i = 0;
        while (i < num_imgs) {
            // fill one staging buffer with as many consecutive needed images of one descriptor type as fit
            char *base = static_cast<char *>(stage[which]);
            const auto t_read0 = std::chrono::steady_clock::now();
            size_t used = 0;
            std::vector<int32_t> ids, nrows;
            std::vector<const void *> ptrs;
            int type = -1;
            while (i < num_imgs) {
                if (!imgs[i].needed) { ++i; continue; }
                const msfm_feature_info &fi = imgs[i].info;
                const size_t bytes = (size_t)fi.desc_rows * 128 * fi.desc_elem_size;
                if (type >= 0 && (fi.desc_type != type || used + bytes > cap)) break;
                type = fi.desc_type;
                imgs[i].xy.resize((size_t)fi.num_pts * 2);
                msfm_feature_path(fold, i, path, sizeof path);
                if (msfm_feature_read(path, &fi, nullptr, nullptr, imgs[i].xy.data(), base + used, (int64_t)128 * fi.desc_elem_size) != 0)
                    return bail(-2, "feature file of image " + std::to_string(i) + " truncated");
                ids.push_back(i);
                nrows.push_back(fi.desc_rows);
                ptrs.push_back(base + used);
                used += (bytes + 255) / 256 * 256;
                ++i;
            }
            if (ids.empty()) break;
            read_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_read0).count();
            msfm_status st;
            const int32_t n = (int32_t)ids.size();
            if (type == 0)
                st = mm ? msfm_multi_upload_u8(mm, n, ids.data(), reinterpret_cast<const uint8_t *const *>(ptrs.data()), nrows.data())
                        : msfm_upload_u8_batch_async(ctx, n, ids.data(), reinterpret_cast<const uint8_t *const *>(ptrs.data()), nrows.data(), nullptr);
            else
                st = mm ? msfm_multi_upload_f32(mm, n, ids.data(), reinterpret_cast<const float *const *>(ptrs.data()), nrows.data(), opt->descriptor_scale)
                        : msfm_upload_f32_batch_async(ctx, n, ids.data(), reinterpret_cast<const float *const *>(ptrs.data()), nrows.data(), opt->descriptor_scale);
            if (st != MSFM_OK) return bail(-4, std::string("upload: ") + (mm ? msfm_multi_last_error(mm) : msfm_last_error(ctx)));
            which ^= 1;
            // the other buffer is filled next: its copies (queued one round ago) must have left it
            const auto t_wait0 = std::chrono::steady_clock::now();
            if (sync_uploads() != MSFM_OK) return bail(-4, "upload synchronisation failed");
            wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wait0).count();
        }
        if (sync_uploads() != MSFM_OK) return bail(-4, "upload synchronisation failed");
        if (clk.on) fprintf(stderr, "[msfm graph]   of which: reading feature files %.1f ms, waiting for uploads %.1f ms\n", read_ms, wait_ms);
        msfm_host_free(stage[0]);
        msfm_host_free(stage[1]);
        stage[0] = stage[1] = nullptr;
    }
    clk.lap("read + upload");

    // ---- the missing images in chunks of consecutive idx1 (host match buffers stay bounded; match_index.txt advances
    //      chunk by chunk, so an interrupted run resumes where it stopped)
    const int64_t chunk_rows = opt->max_batch_rows > 0 ? opt->max_batch_rows : (32ll << 20);
    std::vector<const float *> xy_ptr((size_t)num_imgs, nullptr);
    std::vector<int32_t> npts((size_t)num_imgs, 0);
    for (int32_t i = 0; i < num_imgs; ++i)
        if (imgs[i].needed) { xy_ptr[i] = imgs[i].xy.data(); npts[i] = imgs[i].info.num_pts; }
    std::vector<msfm_pair> pairs;
    std::vector<int64_t> moff;
    std::vector<int32_t> okflags, mbuf, geo_ok, inl, keep, kept;
    std::vector<uint8_t> gbuf, geo_keep;
    int64_t pair_base = 0;  // global index of the chunk's first pair: keeps the RANSAC streams independent of the chunking
    size_t mi = 0;
    while (mi < missing.size()) {
        const size_t chunk_first = mi;
        pairs.clear();
        int64_t capacity = 0;
        while (mi < missing.size()) {
            const int32_t idx1 = missing[mi];
            int64_t need = 0;
            for (int64_t j = offsets[idx1]; j < offsets[idx1 + 1]; ++j) need += imgs[list[j]].info.desc_rows;
            if (mi > chunk_first && capacity + need > chunk_rows) break;
            for (int64_t j = offsets[idx1]; j < offsets[idx1 + 1]; ++j) pairs.push_back({idx1, list[j]});  // index on idx1, queries = idx2 rows
            capacity += need;
            ++mi;
        }
        moff.assign(pairs.size() + 1, 0);
        okflags.assign(pairs.size(), 0);
        mbuf.resize((size_t)(capacity > 0 ? capacity : 1) * 2);
        gbuf.resize((size_t)(capacity > 0 ? capacity : 1));
        if (!pairs.empty()) {
            msfm_params prm;
            memset(&prm, 0, sizeof prm);
            prm.ratio = opt->th_all;
            prm.ratio_good = opt->th_good;
            prm.mutual = opt->mutual;
            prm.min_keypoints = opt->min_keypoints;
            prm.orientation = 0;  // (ptid1, ptid2) ascending ptid2 (fine_matching_graph.cc:121,127)
            prm.rescore_band = opt->rescore_band;
            msfm_result res;
            res.offsets = moff.data();
            res.ok = okflags.data();
            res.matches = reinterpret_cast<int32_t(*)[2]>(mbuf.data());
            res.good = gbuf.data();
            res.match_capacity = capacity;
            msfm_status st = mm ? msfm_multi_match_pairs(mm, pairs.data(), (int64_t)pairs.size(), &prm, &res)
                                : msfm_match_pairs(ctx, pairs.data(), (int64_t)pairs.size(), &prm, &res);
            if (st != MSFM_OK) return bail(-4, std::string("msfm_match_pairs: ") + (mm ? msfm_multi_last_error(mm) : msfm_last_error(ctx)));
        }
        clk.lap("msfm_match_pairs");
        // ---- GeoVerificationFundamental for the chunk on the GPU (fine_matching_graph.cc:137-153)
        const bool gpu_geo = opt->geo_verify && !verify && !pairs.empty();
        if (gpu_geo) {
            msfm_geo_params gp;
            memset(&gp, 0, sizeof gp);
            gp.th_epipolar = 3.0f;   // utils/geo_verification.cc:45,66
            gp.min_points = 30;      // :33
            gp.min_inliers = 30;     // :53
            gp.iters = 1024;
            gp.seed = opt->geo_seed;
            gp.pair_index_base = pair_base;
            geo_ok.assign(pairs.size(), 0);
            inl.assign(pairs.size(), 0);
            geo_keep.assign((size_t)(moff[pairs.size()] > 0 ? moff[pairs.size()] : 1), 0);
            msfm_status st = msfm_geo_verify(ctx, pairs.data(), (int64_t)pairs.size(), moff.data(), reinterpret_cast<const int32_t(*)[2]>(mbuf.data()),
                                             gbuf.data(), xy_ptr.data(), npts.data(), num_imgs, &gp, geo_ok.data(), inl.data(), geo_keep.data(), nullptr);
            if (st != MSFM_OK) return bail(-4, std::string("msfm_geo_verify: ") + msfm_last_error(ctx));
            clk.lap("msfm_geo_verify");
        }
        // ---- verification seam + output, in the reference's order (idx1 ascending over the missing list, partners in
        //      list order); match_index.txt gets its line when idx1 is complete (fine_matching_graph.cc:137-191)
        size_t p = 0;
        for (size_t k1 = chunk_first; k1 < mi; ++k1) {
            const int32_t idx1 = missing[k1];
            for (int64_t j = offsets[idx1]; j < offsets[idx1 + 1]; ++j, ++p) {
                const int32_t idx2 = list[j];
                const int32_t n = (int32_t)(moff[p + 1] - moff[p]);
                const int32_t(*m)[2] = reinterpret_cast<const int32_t(*)[2]>(mbuf.data()) + moff[p];
                const uint8_t *g = gbuf.data() + moff[p];
                if (!okflags[p]) continue;
                keep.resize((size_t)n);
                int32_t n_keep = 0;
                int accept = 1;
                if (verify) {
                    accept = verify(user, idx1, idx2, imgs[idx1].xy.data(), imgs[idx1].info.num_pts, imgs[idx2].xy.data(),
                                    imgs[idx2].info.num_pts, m, g, n, keep.data(), &n_keep);
                } else if (gpu_geo) {
                    accept = geo_ok[p];
                    for (int32_t k = 0; k < n; ++k)
                        if (geo_keep[moff[p] + k]) keep[n_keep++] = k;
                } else {
                    int32_t n_good = 0;
                    for (int32_t k = 0; k < n; ++k) n_good += g[k];
                    accept = n_good >= opt->min_good;
                    for (int32_t k = 0; k < n; ++k) keep[k] = k;
                    n_keep = n;
                }
                if (!accept) continue;
                kept.resize((size_t)n_keep * 2);
                for (int32_t k = 0; k < n_keep; ++k) { kept[2 * k] = m[keep[k]][0]; kept[2 * k + 1] = m[keep[k]][1]; }
                if (msfm_match_append(fold, idx1, idx2, reinterpret_cast<const int32_t(*)[2]>(kept.data()), n_keep) != 0)
                    return bail(-2, "cannot append to the match file of image " + std::to_string(idx1));
                graph[(size_t)idx1 * num_imgs + idx2] = n_keep;
            }
            if (msfm_match_index_append(fold, idx1) != 0) return bail(-2, "cannot append to match_index.txt");
        }
        pair_base += (int64_t)pairs.size();
        clk.lap("verify seam + match files");
    }
    if (mm) msfm_multi_destroy(mm);
    else if (ctx) msfm_destroy(ctx);
    clk.lap("msfm_destroy");
    if (msfm_graph_write(fold, num_imgs, graph.data()) != 0) return fail(err, err_cap, -2, "cannot write graph_matching.txt");
    return 0;
}
