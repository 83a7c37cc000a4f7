// msfm_sched.cc — the pair scheduler's host logic (include/msfm_sched.h): cost-balanced partition of a candidate pair list
// over the GPUs of one box and the stitching of the per-GPU match lists.  No CUDA in here.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/msfm_sched.h"

extern "C" {

int msfm_sched_shard(const msfm_pair *pairs, int64_t n_pairs, const int32_t *rows_per_image, int32_t n_images,
                     int32_t n_workers, int32_t *worker_of_pair, int64_t *cost_per_worker) {
    if (n_pairs < 0 || n_workers < 1 || n_images < 0 || (n_pairs > 0 && (!pairs || !rows_per_image || !worker_of_pair))) return -1;
    for (int64_t i = 0; i < n_pairs; ++i)
        if (pairs[i].ref < 0 || pairs[i].ref >= n_images || pairs[i].query < 0 || pairs[i].query >= n_images) return -1;
    if (cost_per_worker) std::fill(cost_per_worker, cost_per_worker + n_workers, (int64_t)0);
    if (n_pairs == 0) return 0;
    if (n_workers == 1) {
        std::fill(worker_of_pair, worker_of_pair + n_pairs, 0);
        if (cost_per_worker)
            for (int64_t i = 0; i < n_pairs; ++i) cost_per_worker[0] += (int64_t)rows_per_image[pairs[i].ref] * rows_per_image[pairs[i].query];
        return 0;
    }
    // cost of a pair = M x N (x 2 x 128 int8 ops); doubles: 1e6 x 1e6 rows x 1e6 pairs still fits
    std::vector<double> cost((size_t)n_pairs);
    double total = 0.0, largest = 0.0;
    for (int64_t i = 0; i < n_pairs; ++i) {
        cost[i] = (double)rows_per_image[pairs[i].ref] * (double)rows_per_image[pairs[i].query];
        total += cost[i];
        largest = std::max(largest, cost[i]);
    }
    const double cap = std::max(total / (4.0 * n_workers), largest);
    // runs of equal reference id (the reference's own iteration order keeps them adjacent, fine_matching_graph.cc:58-64),
    // cut when a run grows past `cap` so that one popular image cannot unbalance the shards
    struct Run { int64_t first, last; double cost; };
    std::vector<Run> runs;
    int64_t start = 0;
    double acc = 0.0;
    for (int64_t i = 0; i < n_pairs; ++i) {
        const bool new_ref = i > start && pairs[i].ref != pairs[i - 1].ref;
        if (i > start && (new_ref || acc + cost[i] > cap)) {
            runs.push_back({start, i, acc});
            start = i;
            acc = 0.0;
        }
        acc += cost[i];
    }
    runs.push_back({start, n_pairs, acc});
    std::vector<int64_t> order(runs.size());
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return runs[a].cost > runs[b].cost; });
    std::vector<double> load((size_t)n_workers, 0.0);
    for (int64_t g : order) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[w] += runs[g].cost;
        for (int64_t i = runs[g].first; i < runs[g].last; ++i) worker_of_pair[i] = w;
    }
    if (cost_per_worker)
        for (int64_t i = 0; i < n_pairs; ++i)
            cost_per_worker[worker_of_pair[i]] += (int64_t)rows_per_image[pairs[i].ref] * rows_per_image[pairs[i].query];
    return 0;
}

int msfm_sched_image_owner(int32_t n_images, int32_t n_workers, int32_t *owner) {
    if (n_images < 0 || n_workers < 1 || (n_images > 0 && !owner)) return -1;
    const int32_t per = (n_images + n_workers - 1) / n_workers;
    for (int32_t i = 0; i < n_images; ++i) owner[i] = std::min(i / std::max(per, 1), n_workers - 1);
    return 0;
}

int64_t msfm_sched_offsets(const int64_t *counts, int64_t n_pairs, int64_t *offsets) {
    if (n_pairs < 0 || !offsets || (n_pairs > 0 && !counts)) return -1;
    offsets[0] = 0;
    for (int64_t p = 0; p < n_pairs; ++p) offsets[p + 1] = offsets[p] + counts[p];
    return offsets[n_pairs];
}

int msfm_sched_scatter(const int64_t *pair_index, int64_t n_local, const int64_t *local_offsets, const int32_t (*local_matches)[2],
                       const uint8_t *local_good, const int64_t *global_offsets, int32_t (*matches)[2], uint8_t *good) {
    if (n_local < 0 || (n_local > 0 && (!pair_index || !local_offsets || !global_offsets))) return -1;
    for (int64_t k = 0; k < n_local; ++k) {
        const int64_t a = local_offsets[k], n = local_offsets[k + 1] - a;
        if (n <= 0) continue;
        if (!local_matches || !matches) return -1;
        const int64_t dst = global_offsets[pair_index[k]];
        memcpy(matches + dst, local_matches + a, (size_t)n * sizeof(int32_t[2]));
        if (good) {
            if (local_good) memcpy(good + dst, local_good + a, (size_t)n);
            else memset(good + dst, 0, (size_t)n);
        }
    }
    return 0;
}

}  // extern "C"
