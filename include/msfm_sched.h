/*
 * msfm_sched.h — C ABI of the pair scheduler (north_star subsystem 4): shards a candidate pair list over the GPUs of
 * one box and stitches the per-GPU match lists back into the caller's pair order.  Pure host logic, no CUDA: exported by
 * libmsfm_match.so (where msfm_multi.h builds on it) and by the host-only libmsfm_sched.so (CPU tests, other launchers).
 *
 * Reference analogue: the work distribution of FineMatchingGraph::BuildMatchGraph — the serial idx1 loop with an OpenMP
 * team over the partners of one idx1 (/root/reference/SfM/src/graph/fine_matching_graph.cc:58-100).  Pairs are
 * independent units there too; here the unit of distribution is a run of pairs that share the reference image (they
 * re-use the same L2-resident reference tiles), balanced over the devices by cost M x N.
 */
#ifndef MSFM_SCHED_H_
#define MSFM_SCHED_H_

#include <stddef.h>
#include <stdint.h>

#include "msfm_match.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Longest-processing-time partition.  Runs of consecutive pairs with the same `ref` stay together; a run is cut when its
 * cost (sum of rows[ref] * rows[query]) would exceed max(total / (4 * n_workers), largest single pair); runs are handed,
 * most expensive first, to the currently lightest worker (ties: lowest run index / lowest worker).
 * worker_of_pair[n_pairs] receives the worker of every pair; within a worker the caller keeps the original pair order.
 * cost_per_worker[n_workers] (optional) receives the summed cost.  Returns 0, or -1 on a bad argument (null pointer,
 * n_workers < 1, an image id outside [0, n_images)). */
int msfm_sched_shard(const msfm_pair *pairs, int64_t n_pairs, const int32_t *rows_per_image, int32_t n_images,
                     int32_t n_workers, int32_t *worker_of_pair, int64_t *cost_per_worker);

/* Which worker stages which image before replication: contiguous blocks of ceil(n_images / n_workers) ids
 * (one arena range per worker, so that replication is one collective per block). */
int msfm_sched_image_owner(int32_t n_images, int32_t n_workers, int32_t *owner);

/* offsets[0] = 0, offsets[p + 1] = offsets[p] + counts[p]; returns the total. */
int64_t msfm_sched_offsets(const int64_t *counts, int64_t n_pairs, int64_t *offsets);

/* Copy one worker's concatenated match lists to their places in the global result.  The worker matched the pairs
 * pair_index[0 .. n_local) (positions in the caller's list) and holds their lists back to back:
 * local_matches[local_offsets[k] .. local_offsets[k + 1]) belongs to pair pair_index[k] and goes to
 * matches[global_offsets[pair_index[k]] ..).  good / local_good may be NULL. */
int msfm_sched_scatter(const int64_t *pair_index, int64_t n_local, const int64_t *local_offsets,
                       const int32_t (*local_matches)[2], const uint8_t *local_good, const int64_t *global_offsets,
                       int32_t (*matches)[2], uint8_t *good);

#ifdef __cplusplus
}
#endif
#endif /* MSFM_SCHED_H_ */
