/*
 * msfm_graph.h — C ABI of the fine-matching-graph driver: the caller of the matching hot path, rebuilt on top of
 * msfm_match.h (GPU matcher) and msfm_store.h (the reference's on-disk formats).  libmsfm_graph.so.
 *
 * Replaces FineMatchingGraph::BuildMatchGraph (/root/reference/SfM/src/graph/fine_matching_graph.cc:40-194):
 *   resume     CheckMissingMatchingFile / RecoverMatchingGraph (:49-54, :209-244, :294-330)
 *   matching   per idx1: FLANN index + OpenMP partner loop + ratio 0.6 "good" / 0.85 "all" (:58-133)
 *                -> every image's descriptors are read from its <idx>_feature file and staged in HBM once, the whole
 *                   candidate pair list of the missing images goes through msfm_match_pairs in one batch
 *   verify     GeoVerification::GeoVerificationFundamental on the good set, F-filter on the all set (:137-153)
 *                -> msfm_graph_options.geo_verify: msfm_geo_verify on the GPU for the whole batch; or the msfm_verify_fn
 *                   callback seam; with neither, every "all" match of a pair with >= `min_good` good matches is kept
 *   output     WriteOutMatches per accepted pair, match_index.txt line per finished idx1, graph_matching.txt (:181-193)
 * The files written are byte-identical to what the reference writes for the same match lists.
 */
#ifndef MSFM_GRAPH_H_
#define MSFM_GRAPH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msfm_graph_options {
    int32_t device;          /* CUDA ordinal */
    float th_good;           /* 0.6  fine_matching_graph.cc:42 */
    float th_all;            /* 0.85 fine_matching_graph.cc:43 */
    int32_t mutual;          /* mutual cross-check (SiftMatchGPU semantics); the CPU reference has none */
    int32_t min_keypoints;   /* pairs with fewer keypoints on either side are skipped; 0 = the FLANN path has no gate */
    int32_t min_good;        /* without a verifier: pairs with fewer good matches are dropped (GeoVerificationFundamental
                                needs >= 30 points, utils/geo_verification.cc:30-58); 0 = keep all */
    float descriptor_scale;  /* float descriptors: quantisation scale (1 for 512-scaled VLSIFT rows, 512 for unit-norm) */
    float rescore_band;      /* float descriptors: fp32 re-scoring band (msfm_params.rescore_band); 0 = off */
    int32_t geo_verify;      /* 1 (and no callback): the reference's GeoVerificationFundamental stages on the GPU for the
                                whole batch (msfm_geo_verify: >= 30 good matches, RANSAC-F 3 px, >= 30 inliers, then the
                                F-filter of the "all" set); rejected pairs leave no record, like the reference */
    uint32_t geo_seed;       /* RANSAC stream of msfm_geo_verify */
    int64_t max_batch_rows;  /* the missing images are processed in chunks of consecutive idx1 whose pair lists need at most
                                this many match slots (sum of query rows) on the host; match_index.txt advances per chunk.
                                0 = 32 Mi.  Results do not depend on the chunking. */
    int32_t n_devices;       /* > 1: the pair list is matched on devices 0 .. n_devices-1 by the single-process multi-GPU
                                engine (msfm_multi.h); `device` is then ignored.  0 / 1 = one GPU (`device`) */
    int32_t reserved;        /* must be zero */
} msfm_graph_options;

/* Geo-verification seam.  xy1/xy2: centred keypoints of both images as stored in the feature files; matches: the
 * pair's "all" list (point id in idx1, point id in idx2) ascending id2 with its good flags.  The callback writes the
 * indices (into `matches`) of the matches to keep, ascending, and returns 1 to accept the pair, 0 to drop it. */
typedef int (*msfm_verify_fn)(void *user, int32_t idx1, int32_t idx2, const float *xy1, int32_t n1, const float *xy2, int32_t n2,
                              const int32_t (*matches)[2], const uint8_t *good, int32_t n_matches, int32_t *keep, int32_t *n_keep);

/* Partners of image i are list[offsets[i] .. offsets[i+1]) (match_graph_init).  Returns 0 on success; on failure a
 * negative code and a message in err. */
int msfm_build_match_graph(const char *fold, int32_t num_imgs, const int64_t *offsets, const int32_t *list,
                           const msfm_graph_options *opt, msfm_verify_fn verify, void *user, char *err, size_t err_cap);

#ifdef __cplusplus
}
#endif
#endif /* MSFM_GRAPH_H_ */
