// msfm_store.cc — the reference's on-disk formats around the matching hot path (include/msfm_store.h).
// Host-only C++; every function cites the reference lines whose byte layout / text layout it reproduces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/msfm_store.h"

namespace {

std::string join(const char *fold, const std::string &name) { return std::string(fold) + "//" + name; }  // database.cc:360

int copy_path(const std::string &s, char *out, size_t cap) {
    if (!out || cap < s.size() + 1) return MSFM_STORE_ERR_CAPACITY;
    memcpy(out, s.c_str(), s.size() + 1);
    return MSFM_STORE_OK;
}

struct File {
    FILE *f = nullptr;
    File(const std::string &p, const char *mode) : f(fopen(p.c_str(), mode)) {}
    ~File() { if (f) fclose(f); }
    bool rd(void *dst, size_t n) { return n == 0 || fread(dst, 1, n, f) == n; }
    bool wr(const void *src, size_t n) { return n == 0 || fwrite(src, 1, n, f) == n; }
};

// OpenCV type code -> bytes per element (CV_MAT_DEPTH = type & 7, channels = (type >> 3) + 1).
int elem_size_of(int32_t type) {
    static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type & 7] * ((type >> 3) + 1);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------ <idx>_feature
int msfm_feature_path(const char *fold, int32_t idx, char *out, size_t cap) {
    if (!fold) return MSFM_STORE_ERR_ARG;
    return copy_path(join(fold, std::to_string(idx) + "_feature"), out, cap);
}

// Field order of Database::ReadinImageFeatures, database.cc:373-419.
int msfm_feature_stat(const char *path, msfm_feature_info *info) {
    if (!path || !info) return MSFM_STORE_ERR_ARG;
    File fp(path, "rb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    memset(info, 0, sizeof *info);
    if (!fp.rd(&info->rows, 4) || !fp.rd(&info->cols, 4) || !fp.rd(&info->zoom_ratio, 4) || !fp.rd(&info->f_mm, 4) ||
        !fp.rd(&info->f_pixel, 4) || !fp.rd(&info->gps_latitude, 4) || !fp.rd(&info->gps_longitude, 4))
        return MSFM_STORE_ERR_FORMAT;
    if (!fp.rd(&info->maker_len, 4) || info->maker_len < 0 || fseek(fp.f, info->maker_len, SEEK_CUR) != 0) return MSFM_STORE_ERR_FORMAT;
    if (!fp.rd(&info->model_len, 4) || info->model_len < 0 || fseek(fp.f, info->model_len, SEEK_CUR) != 0) return MSFM_STORE_ERR_FORMAT;
    if (!fp.rd(&info->num_pts, 4) || info->num_pts < 0) return MSFM_STORE_ERR_FORMAT;
    info->keypoints_offset = ftell(fp.f);
    if (fseek(fp.f, (long)info->num_pts * 8, SEEK_CUR) != 0) return MSFM_STORE_ERR_FORMAT;
    if (!fp.rd(&info->desc_rows, 4) || !fp.rd(&info->desc_cols, 4) || !fp.rd(&info->desc_type, 4)) return MSFM_STORE_ERR_FORMAT;
    if (info->desc_rows < 0 || info->desc_cols < 0) return MSFM_STORE_ERR_FORMAT;
    info->desc_elem_size = elem_size_of(info->desc_type);
    info->desc_offset = ftell(fp.f);
    // the payload must be complete
    if (fseek(fp.f, 0, SEEK_END) != 0) return MSFM_STORE_ERR_FORMAT;
    const int64_t need = info->desc_offset + (int64_t)info->desc_rows * info->desc_cols * info->desc_elem_size;
    if (ftell(fp.f) < need) return MSFM_STORE_ERR_FORMAT;
    return MSFM_STORE_OK;
}

int msfm_feature_read(const char *path, const msfm_feature_info *info, char *maker, char *model, float *xy, void *desc,
                      int64_t desc_row_stride_bytes) {
    if (!path || !info) return MSFM_STORE_ERR_ARG;
    File fp(path, "rb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    const long maker_off = 7 * 4 + 4, model_off = maker_off + info->maker_len + 4;
    if (maker) {
        if (fseek(fp.f, maker_off, SEEK_SET) != 0 || !fp.rd(maker, info->maker_len)) return MSFM_STORE_ERR_FORMAT;
        maker[info->maker_len] = '\0';
    }
    if (model) {
        if (fseek(fp.f, model_off, SEEK_SET) != 0 || !fp.rd(model, info->model_len)) return MSFM_STORE_ERR_FORMAT;
        model[info->model_len] = '\0';
    }
    if (xy) {
        if (fseek(fp.f, (long)info->keypoints_offset, SEEK_SET) != 0 || !fp.rd(xy, (size_t)info->num_pts * 8)) return MSFM_STORE_ERR_FORMAT;
    }
    if (desc) {
        const int64_t row_bytes = (int64_t)info->desc_cols * info->desc_elem_size;
        if (desc_row_stride_bytes < row_bytes) return MSFM_STORE_ERR_ARG;
        if (fseek(fp.f, (long)info->desc_offset, SEEK_SET) != 0) return MSFM_STORE_ERR_FORMAT;
        if (desc_row_stride_bytes == row_bytes) {
            if (!fp.rd(desc, (size_t)(row_bytes * info->desc_rows))) return MSFM_STORE_ERR_FORMAT;
        } else {
            for (int32_t r = 0; r < info->desc_rows; ++r)
                if (!fp.rd(static_cast<char *>(desc) + r * desc_row_stride_bytes, (size_t)row_bytes)) return MSFM_STORE_ERR_FORMAT;
        }
    }
    return MSFM_STORE_OK;
}

// Database::WriteoutImageFeature, database.cc:490-541.
int msfm_feature_write(const char *path, const msfm_feature_info *info, const char *maker, const char *model,
                       const float *xy_pixel, const void *desc, int64_t desc_row_stride_bytes) {
    if (!path || !info || (info->num_pts > 0 && !xy_pixel) || (info->desc_rows > 0 && info->desc_cols > 0 && !desc)) return MSFM_STORE_ERR_ARG;
    const int32_t maker_len = maker ? (int32_t)strlen(maker) : 0, model_len = model ? (int32_t)strlen(model) : 0;
    const int64_t row_bytes = (int64_t)info->desc_cols * elem_size_of(info->desc_type);
    if (desc && desc_row_stride_bytes < row_bytes) return MSFM_STORE_ERR_ARG;
    File fp(path, "wb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    bool ok = fp.wr(&info->rows, 4) && fp.wr(&info->cols, 4) && fp.wr(&info->zoom_ratio, 4) && fp.wr(&info->f_mm, 4) &&
              fp.wr(&info->f_pixel, 4) && fp.wr(&info->gps_latitude, 4) && fp.wr(&info->gps_longitude, 4);
    ok = ok && fp.wr(&maker_len, 4) && fp.wr(maker, maker_len) && fp.wr(&model_len, 4) && fp.wr(model, model_len);
    ok = ok && fp.wr(&info->num_pts, 4);
    std::vector<float> centred((size_t)info->num_pts * 2);
    for (int32_t i = 0; i < info->num_pts; ++i) {  // "points are centralized", database.cc:522-527 (double arithmetic)
        centred[2 * i + 0] = (float)(xy_pixel[2 * i + 0] - info->cols / 2.0);
        centred[2 * i + 1] = (float)(xy_pixel[2 * i + 1] - info->rows / 2.0);
    }
    ok = ok && fp.wr(centred.data(), centred.size() * 4);
    ok = ok && fp.wr(&info->desc_rows, 4) && fp.wr(&info->desc_cols, 4) && fp.wr(&info->desc_type, 4);
    for (int32_t r = 0; ok && r < info->desc_rows; ++r) ok = fp.wr(static_cast<const char *>(desc) + r * desc_row_stride_bytes, (size_t)row_bytes);
    return ok ? MSFM_STORE_OK : MSFM_STORE_ERR_FORMAT;
}

// ------------------------------------------------------------------------------------------------ <idx1>_match
int msfm_match_path(const char *fold, int32_t idx1, char *out, size_t cap) {
    if (!fold) return MSFM_STORE_ERR_ARG;
    return copy_path(join(fold, std::to_string(idx1) + "_match"), out, cap);
}

// FineMatchingGraph::WriteOutMatches, fine_matching_graph.cc:247-272.
int msfm_match_append(const char *fold, int32_t idx1, int32_t idx2, const int32_t (*pairs)[2], int32_t n) {
    if (!fold || n < 0 || (n > 0 && !pairs)) return MSFM_STORE_ERR_ARG;
    if (n == 0) return MSFM_STORE_OK;  // the reference returns before touching the file
    File fp(join(fold, std::to_string(idx1) + "_match"), "ab");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    return fp.wr(&idx2, 4) && fp.wr(&n, 4) && fp.wr(pairs, (size_t)n * 8) ? MSFM_STORE_OK : MSFM_STORE_ERR_FORMAT;
}

// Graph::QueryMatch, graph.cc:92-121.
int msfm_match_read(const char *fold, int32_t idx1, int32_t *idx2, int64_t *offsets, int32_t record_cap, int32_t (*pairs)[2],
                    int64_t pair_cap, int32_t *n_records, int64_t *n_pairs) {
    if (!fold || !n_records || !n_pairs) return MSFM_STORE_ERR_ARG;
    *n_records = 0;
    *n_pairs = 0;
    File fp(join(fold, std::to_string(idx1) + "_match"), "rb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    int32_t id = 0, num = 0;
    bool fits = true;
    while (fp.rd(&id, 4)) {
        if (!fp.rd(&num, 4) || num < 0) return MSFM_STORE_ERR_FORMAT;
        const bool store = idx2 && offsets && *n_records < record_cap && (num == 0 || (pairs && *n_pairs + num <= pair_cap));
        if (store) {
            idx2[*n_records] = id;
            offsets[*n_records] = *n_pairs;
            if (!fp.rd(pairs + *n_pairs, (size_t)num * 8)) return MSFM_STORE_ERR_FORMAT;
            offsets[*n_records + 1] = *n_pairs + num;
        } else {
            fits = false;
            if (fseek(fp.f, (long)num * 8, SEEK_CUR) != 0) return MSFM_STORE_ERR_FORMAT;
        }
        *n_records += 1;
        *n_pairs += num;
    }
    return fits ? MSFM_STORE_OK : MSFM_STORE_ERR_CAPACITY;
}

// ------------------------------------------------------------------------------------------------ match_index.txt
// FineMatchingGraph::CheckMissingMatchingFile, fine_matching_graph.cc:209-244.
int msfm_match_index_missing(const char *fold, int32_t num_imgs, int32_t *missing, int32_t *n_missing) {
    if (!fold || num_imgs < 0 || !missing || !n_missing) return MSFM_STORE_ERR_ARG;
    std::vector<char> done((size_t)num_imgs, 0);
    File fp(join(fold, "match_index.txt"), "r");
    if (fp.f) {
        int idx;
        while (fscanf(fp.f, "%d", &idx) == 1)
            if (idx >= 0 && idx < num_imgs) done[idx] = 1;
    }
    *n_missing = 0;
    for (int32_t i = 0; i < num_imgs; ++i)
        if (!done[i]) missing[(*n_missing)++] = i;
    return MSFM_STORE_OK;
}

// `of_match_index << idx1 << std::endl`, fine_matching_graph.cc:57,191.
int msfm_match_index_append(const char *fold, int32_t idx1) {
    if (!fold) return MSFM_STORE_ERR_ARG;
    File fp(join(fold, "match_index.txt"), "a");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    return fprintf(fp.f, "%d\n", idx1) > 0 ? MSFM_STORE_OK : MSFM_STORE_ERR_FORMAT;
}

// ------------------------------------------------------------------------------------------------ graph_matching.txt
// FineMatchingGraph::WriteOutMatchGraph, fine_matching_graph.cc:275-292: "<count> " per entry, newline per row.
int msfm_graph_write(const char *fold, int32_t num_imgs, const int32_t *graph) {
    if (!fold || num_imgs < 0 || (num_imgs > 0 && !graph)) return MSFM_STORE_ERR_ARG;
    File fp(join(fold, "graph_matching.txt"), "wb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    for (int32_t i = 0; i < num_imgs; ++i) {
        for (int32_t j = 0; j < num_imgs; ++j) fprintf(fp.f, "%d ", graph[(int64_t)i * num_imgs + j]);
        fputc('\n', fp.f);
    }
    return MSFM_STORE_OK;
}

// Graph::ReadinMatchingGraph, graph.cc:72-85.
int msfm_graph_read(const char *fold, int32_t num_imgs, int32_t *graph) {
    if (!fold || num_imgs < 0 || (num_imgs > 0 && !graph)) return MSFM_STORE_ERR_ARG;
    File fp(join(fold, "graph_matching.txt"), "rb");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    for (int64_t i = 0; i < (int64_t)num_imgs * num_imgs; ++i)
        if (fscanf(fp.f, "%d", &graph[i]) != 1) return MSFM_STORE_ERR_FORMAT;
    return MSFM_STORE_OK;
}

// FineMatchingGraph::RecoverMatchingGraph, fine_matching_graph.cc:294-330.
int msfm_graph_recover(const char *fold, int32_t num_imgs, const int32_t *existing, int32_t n_existing, int32_t *graph) {
    if (!fold || num_imgs < 0 || n_existing < 0 || (n_existing > 0 && !existing) || (num_imgs > 0 && !graph)) return MSFM_STORE_ERR_ARG;
    std::fill(graph, graph + (int64_t)num_imgs * num_imgs, 0);
    for (int32_t e = 0; e < n_existing; ++e) {
        const int32_t idx = existing[e];
        if (idx < 0 || idx >= num_imgs) return MSFM_STORE_ERR_ARG;
        File fp(join(fold, std::to_string(idx) + "_match"), "rb");
        if (!fp.f) continue;
        int32_t id = 0, num = 0;
        while (fp.rd(&id, 4)) {
            if (!fp.rd(&num, 4) || num < 0 || fseek(fp.f, (long)num * 8, SEEK_CUR) != 0) return MSFM_STORE_ERR_FORMAT;
            if (id >= 0 && id < num_imgs) graph[(int64_t)idx * num_imgs + id] = num;
        }
    }
    return MSFM_STORE_OK;
}

// ------------------------------------------------------------------------------------------------ pair lists
// matching_type == "all", initial_matching_graph.cc:55-64.
int msfm_pairs_all(int32_t num_imgs, int64_t *offsets, int32_t *list) {
    if (num_imgs < 0 || !offsets || (num_imgs > 1 && !list)) return MSFM_STORE_ERR_ARG;
    int64_t n = 0;
    for (int32_t i = 0; i < num_imgs; ++i) {
        offsets[i] = n;
        for (int32_t j = 0; j < num_imgs; ++j)
            if (j != i) list[n++] = j;
    }
    offsets[num_imgs] = n;
    return MSFM_STORE_OK;
}

// match_graph_priori_xy, initial_matching_graph.cc:114-162.
int msfm_pairs_priori_xy(int32_t num_imgs, const double *xy, int32_t knn, int64_t *offsets, int32_t *list) {
    if (num_imgs < 0 || knn < 0 || !offsets || (num_imgs > 0 && !xy)) return MSFM_STORE_ERR_ARG;
    const double th_dis = 1.0;
    auto by_dist_then_id = [](const std::pair<int, double> &l, const std::pair<int, double> &r) {
        return l.second < r.second || (l.second == r.second && l.first < r.first);
    };
    // images whose x + y lies within th_dis of the previously kept one count as redundant
    std::vector<std::pair<int, double>> id_dis((size_t)num_imgs);
    for (int32_t i = 0; i < num_imgs; ++i) id_dis[i] = {i, xy[2 * i] + xy[2 * i + 1]};
    std::sort(id_dis.begin(), id_dis.end(), by_dist_then_id);
    std::vector<char> redundant((size_t)num_imgs, 0);
    if (num_imgs > 0) {
        double dis_pre = id_dis[0].second - 100.0;
        for (const auto &e : id_dis) {
            // the reference calls abs() on a double here; with <cmath> in scope that is the floating-point overload
            if (std::fabs(e.second - dis_pre) < th_dis) redundant[e.first] = 1;
            else dis_pre = e.second;
        }
    }
    const int32_t k = std::min(knn, num_imgs / 10);
    int64_t n = 0;
    std::vector<std::pair<int, double>> info;
    for (int32_t i = 0; i < num_imgs; ++i) {
        offsets[i] = n;
        if (redundant[i]) continue;
        info.clear();
        for (int32_t j = 0; j < num_imgs; ++j)
            if (j != i && !redundant[j]) info.push_back({j, std::fabs(xy[2 * i] - xy[2 * j]) + std::fabs(xy[2 * i + 1] - xy[2 * j + 1])});
        std::sort(info.begin(), info.end(), by_dist_then_id);
        const int32_t t = std::min<int32_t>((int32_t)info.size(), k);
        for (int32_t j = 0; j < t; ++j) list[n++] = info[j].first;
    }
    offsets[num_imgs] = n;
    return MSFM_STORE_OK;
}

// ---- BoW retrieval route -------------------------------------------------------------------------------------------
namespace {
// math::keep_unique_vector (utils/basic_funcs.h:126-151) as it behaves: after sorting, a value is kept when its run has
// length one, except that the first run is never "unique" (the loop starts by comparing data[0] with itself) and the last
// run is never flushed.
std::vector<int32_t> unique_words_of_image(const int32_t *w, int64_t n) {
    std::vector<int32_t> data(w, w + n), kept;
    if (data.empty()) return kept;
    std::sort(data.begin(), data.end());
    int32_t v = data[0];
    bool is_unique = true;
    for (int32_t x : data) {
        if (x != v) {
            if (is_unique) kept.push_back(v);
            v = x;
            is_unique = true;
        } else {
            is_unique = false;
        }
    }
    return kept;
}

// math::keep_unique_idx_vector (utils/basic_funcs.cc:380-406) + the pt_word_map insertion of
// initial_matching_graph.cc:194-201: (word, keypoint index) sorted by word; whenever the value changes and the run that
// just ended was a singleton (and not the first run), the FIRST element of the new run is recorded.
std::vector<std::pair<int32_t, int32_t>> pt_word_map(const int32_t *w, int32_t n) {
    std::vector<std::pair<int32_t, int32_t>> d((size_t)n), out;  // (index, word)
    for (int32_t i = 0; i < n; ++i) d[i] = {i, w[i]};
    std::stable_sort(d.begin(), d.end(), [](const std::pair<int32_t, int32_t> &l, const std::pair<int32_t, int32_t> &r) { return l.second < r.second; });
    if (d.empty()) return out;
    bool is_unique = true;
    int32_t v = d[0].second;
    for (const auto &e : d) {
        if (e.second != v) {
            if (is_unique) out.push_back({e.second, e.first});  // map key = the new run's word, value = its first keypoint
            v = e.second;
            is_unique = true;
        } else {
            is_unique = false;
        }
    }
    return out;  // ascending word id, each word at most once
}
}  // namespace

int msfm_similarity_invfile(int32_t num_imgs, const int64_t *word_offsets, const int32_t *word_ids, int32_t num_words, float *similarity) {
    if (num_imgs < 0 || num_words < 0 || !word_offsets || !similarity || (word_offsets[num_imgs] > 0 && !word_ids)) return MSFM_STORE_ERR_ARG;
    std::vector<std::vector<int32_t>> inverted((size_t)num_words);
    for (int32_t i = 0; i < num_imgs; ++i) {
        const int64_t a = word_offsets[i], n = word_offsets[i + 1] - a;
        if (n <= 0) continue;
        for (int32_t id : unique_words_of_image(word_ids + a, n)) {
            if (id < 0 || id >= num_words) return MSFM_STORE_ERR_ARG;
            inverted[id].push_back(i);
        }
    }
    const int32_t th_bin_size = num_words / 100;  // similarity_graph.cc:108
    std::fill(similarity, similarity + (size_t)num_imgs * num_imgs, 0.0f);
    for (const std::vector<int32_t> &bin : inverted) {
        if (bin.empty() || (int64_t)bin.size() > th_bin_size) continue;
        for (size_t m = 0; m + 1 < bin.size(); ++m)
            for (size_t n = m + 1; n < bin.size(); ++n) {
                similarity[(size_t)bin[m] * num_imgs + bin[n]] += 1.0f;
                similarity[(size_t)bin[n] * num_imgs + bin[m]] += 1.0f;
            }
    }
    return MSFM_STORE_OK;
}

int msfm_pairs_similarity_topk(int32_t num_imgs, const float *similarity, int32_t th_num_match, int64_t *offsets, int32_t *list) {
    if (num_imgs < 0 || !offsets || (num_imgs > 0 && (!similarity || !list))) return MSFM_STORE_ERR_ARG;
    if (th_num_match <= 0) {
        th_num_match = std::min(std::max(200, num_imgs / 10), num_imgs - 1);  // initial_matching_graph.cc:166-168
        if (th_num_match > 500) th_num_match = 500;
    }
    int64_t n = 0;
    std::vector<std::pair<int32_t, float>> sim_sort;
    for (int32_t i = 0; i < num_imgs; ++i) {
        offsets[i] = n;
        sim_sort.clear();
        for (int32_t j = 0; j < num_imgs; ++j)
            if (j != i && !(similarity[(size_t)i * num_imgs + j] < 0)) sim_sort.push_back({j, similarity[(size_t)i * num_imgs + j]});
        std::stable_sort(sim_sort.begin(), sim_sort.end(), [](const std::pair<int32_t, float> &l, const std::pair<int32_t, float> &r) { return l.second > r.second; });
        const int32_t t = std::min<int32_t>((int32_t)sim_sort.size(), std::max(th_num_match, 0));
        for (int32_t j = 0; j < t; ++j) list[n++] = sim_sort[j].first;
    }
    offsets[num_imgs] = n;
    return MSFM_STORE_OK;
}

int msfm_word_matches(const int32_t *words1, int32_t n1, const int32_t *words2, int32_t n2, int32_t (*matches)[2], int32_t cap) {
    if (n1 < 0 || n2 < 0 || (n1 > 0 && !words1) || (n2 > 0 && !words2) || cap < 0 || (cap > 0 && !matches)) return MSFM_STORE_ERR_ARG;
    const auto m1 = pt_word_map(words1, n1), m2 = pt_word_map(words2, n2);
    int32_t n = 0;
    size_t b = 0;
    for (const auto &e : m1) {  // both ascending by word: a merge instead of the reference's std::map lookups
        while (b < m2.size() && m2[b].first < e.first) ++b;
        if (b < m2.size() && m2[b].first == e.first) {
            if (n < cap) { matches[n][0] = e.second; matches[n][1] = m2[b].second; }
            ++n;
        }
    }
    return n;
}

// WriteOutInitMatchGraph, initial_matching_graph.cc:324-344.
int msfm_init_graph_write(const char *fold, int32_t num_imgs, int32_t id_last, const int64_t *offsets, const int32_t *list) {
    if (!fold || num_imgs < 0 || !offsets) return MSFM_STORE_ERR_ARG;
    File fp(join(fold, "init_match_graph.txt"), "w");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    fprintf(fp.f, "%d\n%d\n", num_imgs, id_last);
    for (int32_t i = 0; i < num_imgs; ++i) {
        fprintf(fp.f, "%lld ", (long long)(offsets[i + 1] - offsets[i]));
        for (int64_t j = offsets[i]; j < offsets[i + 1]; ++j) fprintf(fp.f, "%d ", list[j]);
        fputc('\n', fp.f);
    }
    return MSFM_STORE_OK;
}

// ReadinInitMatchGraph, initial_matching_graph.cc:296-322.
int msfm_init_graph_read(const char *fold, int32_t *num_imgs, int32_t *id_last, int64_t *offsets, int32_t offsets_cap,
                         int32_t *list, int64_t list_cap, int64_t *n_list) {
    if (!fold || !num_imgs || !id_last || !n_list) return MSFM_STORE_ERR_ARG;
    File fp(join(fold, "init_match_graph.txt"), "r");
    if (!fp.f) return MSFM_STORE_ERR_OPEN;
    if (fscanf(fp.f, "%d", num_imgs) != 1 || fscanf(fp.f, "%d", id_last) != 1 || *num_imgs < 0) return MSFM_STORE_ERR_FORMAT;
    bool fits = offsets && offsets_cap >= *num_imgs + 1;
    *n_list = 0;
    for (int32_t i = 0; i < *num_imgs; ++i) {
        int cnt = 0;
        if (fscanf(fp.f, "%d", &cnt) != 1 || cnt < 0) return MSFM_STORE_ERR_FORMAT;
        if (fits) offsets[i] = *n_list;
        for (int j = 0; j < cnt; ++j) {
            int v;
            if (fscanf(fp.f, "%d", &v) != 1) return MSFM_STORE_ERR_FORMAT;
            if (fits && list && *n_list < list_cap) list[*n_list] = v;
            else fits = false;
            *n_list += 1;
        }
    }
    if (fits) offsets[*num_imgs] = *n_list;
    return fits ? MSFM_STORE_OK : MSFM_STORE_ERR_CAPACITY;
}

}  // extern "C"
