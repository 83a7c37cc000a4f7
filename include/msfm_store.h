/*
 * msfm_store.h — C ABI of the on-disk formats either side of the matching hot path:
 * the per-image `<idx>_feature` store the descriptors come from, the candidate pair lists that drive the matcher, and
 * the `<idx1>_match` / `match_index.txt` / `graph_matching.txt` files its results go to.  Byte-compatible with the
 * reference so that the unchanged downstream stages (Graph::QueryMatch, IncrementalSfM) read what this writes.
 * Host-only (no CUDA): libmsfm_store.so.  Every call returns 0 on success, a negative msfm_store_status otherwise.
 *
 * Reference interfaces replaced (paths relative to /root/reference/SfM):
 *   Database::ReadinImageFeatures / WriteoutImageFeature          src/database.cc:352-423, 490-541
 *   FineMatchingGraph::WriteOutMatches                            src/graph/fine_matching_graph.cc:247-272
 *   Graph::QueryMatch (reader of <idx>_match)                     src/graph.cc:92-137
 *   FineMatchingGraph::CheckMissingMatchingFile (match_index.txt) src/graph/fine_matching_graph.cc:209-244
 *   FineMatchingGraph::WriteOutMatchGraph / RecoverMatchingGraph  src/graph/fine_matching_graph.cc:275-330
 *   Graph::ReadinMatchingGraph                                    src/graph.cc:72-85
 *   InitialMatchingGraph "all" and "priori xy" pair lists         src/graph/initial_matching_graph.cc:55-64, 114-162
 *   InitialMatchingGraph::ReadinInitMatchGraph / WriteOut...      src/graph/initial_matching_graph.cc:296-344
 * All files live in one output folder; paths are built as `fold + "//" + name` exactly like the reference.
 */
#ifndef MSFM_STORE_H_
#define MSFM_STORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum msfm_store_status {
    MSFM_STORE_OK = 0,
    MSFM_STORE_ERR_ARG = -1,      /* null pointer / negative size */
    MSFM_STORE_ERR_OPEN = -2,     /* file cannot be opened */
    MSFM_STORE_ERR_FORMAT = -3,   /* truncated or inconsistent file */
    MSFM_STORE_ERR_CAPACITY = -4  /* caller buffer too small (the needed size is still reported) */
} msfm_store_status;

/* ---- <idx>_feature --------------------------------------------------------------------------------------------
 * Layout (database.cc:500-536): int rows, cols; float zoom_ratio, f_mm, f_pixel, gps_latitude, gps_longitude;
 * int len + maker bytes; int len + model bytes; int num_pts; float xy[2*num_pts] (centred: x - cols/2, y - rows/2);
 * int desc_rows, desc_cols, desc_type (OpenCV type code: 5 = CV_32FC1, 0 = CV_8UC1); raw descriptor bytes. */
typedef struct msfm_feature_info {
    int32_t rows, cols; /* image size */
    float zoom_ratio, f_mm, f_pixel, gps_latitude, gps_longitude;
    int32_t maker_len, model_len;
    int32_t num_pts;
    int32_t desc_rows, desc_cols, desc_type;
    int32_t desc_elem_size;   /* bytes per element derived from desc_type (4 for CV_32F, 1 for CV_8U) */
    int64_t keypoints_offset; /* file offsets of the two payload blocks */
    int64_t desc_offset;
} msfm_feature_info;

int msfm_feature_path(const char *fold, int32_t idx, char *out, size_t cap);
/* Parse the header only. */
int msfm_feature_stat(const char *path, msfm_feature_info *info);
/* Read the pieces a caller wants (any pointer may be NULL).  maker/model: NUL-terminated, capacity >= len + 1.
 * xy: 2*num_pts floats as stored (centred).  desc: desc_rows rows of desc_cols*elem_size bytes written with the given
 * row stride (>= row bytes), e.g. straight into a pinned staging buffer for msfm_upload_f32 / msfm_upload_u8. */
int msfm_feature_read(const char *path, const msfm_feature_info *info, char *maker, char *model, float *xy, void *desc,
                      int64_t desc_row_stride_bytes);
/* Write a feature file.  xy_pixel are image coordinates; they are centred on write like the reference does
 * (x - cols/2.0, y - rows/2.0 in double, stored as float; database.cc:522-527). */
int msfm_feature_write(const char *path, const msfm_feature_info *info, const char *maker, const char *model,
                       const float *xy_pixel, const void *desc, int64_t desc_row_stride_bytes);

/* ---- <idx1>_match -----------------------------------------------------------------------------------------------
 * Appended records {int idx2; int n; int pairs[2n]} with pairs = (point id in idx1, point id in idx2);
 * nothing is written for n == 0 (fine_matching_graph.cc:250-253). */
int msfm_match_path(const char *fold, int32_t idx1, char *out, size_t cap);
int msfm_match_append(const char *fold, int32_t idx1, int32_t idx2, const int32_t (*pairs)[2], int32_t n);
/* Read every record of <idx1>_match.  n_records / n_pairs always receive the totals in the file; data is stored only
 * while it fits (MSFM_STORE_ERR_CAPACITY otherwise).  offsets[r] .. offsets[r+1] index `pairs` for record r. */
int msfm_match_read(const char *fold, int32_t idx1, int32_t *idx2, int64_t *offsets, int32_t record_cap,
                    int32_t (*pairs)[2], int64_t pair_cap, int32_t *n_records, int64_t *n_pairs);

/* ---- match_index.txt (resume) --------------------------------------------------------------------------------- */
/* Images whose matching has not been recorded as finished (all of them when the file is absent). */
int msfm_match_index_missing(const char *fold, int32_t num_imgs, int32_t *missing, int32_t *n_missing);
int msfm_match_index_append(const char *fold, int32_t idx1);

/* ---- graph_matching.txt ------------------------------------------------------------------------------------------
 * num_imgs lines of num_imgs match counts, each followed by a blank (fine_matching_graph.cc:283-288). */
int msfm_graph_write(const char *fold, int32_t num_imgs, const int32_t *graph);
int msfm_graph_read(const char *fold, int32_t num_imgs, int32_t *graph);
/* RecoverMatchingGraph: zero the graph, then graph[idx][idx2] = n for every record of the listed <idx>_match files. */
int msfm_graph_recover(const char *fold, int32_t num_imgs, const int32_t *existing, int32_t n_existing, int32_t *graph);

/* ---- candidate pair lists (the matcher's input) ----------------------------------------------------------------
 * Adjacency form like match_graph_init: partners of image i are list[offsets[i] .. offsets[i+1]). */
/* "all": every j != i, ascending (initial_matching_graph.cc:55-64).  list capacity num_imgs*(num_imgs-1). */
int msfm_pairs_all(int32_t num_imgs, int64_t *offsets, int32_t *list);
/* "priori xy": redundancy filter on x+y (th 1.0) then the k = min(knn, num_imgs/10) nearest images in L1 distance
 * (initial_matching_graph.cc:114-162).  Ties in distance are broken by the lower image index (the reference's
 * std::sort leaves them unspecified).  list capacity num_imgs*k. */
int msfm_pairs_priori_xy(int32_t num_imgs, const double *xy, int32_t knn, int64_t *offsets, int32_t *list);

/* "feature" (BoW retrieval) route, InitialMatchingGraph::match_graph_feature (initial_matching_graph.cc:164-294) on top of
 * SimilarityGraph::SimilarityGraphInvFile (similarity_graph.cc:47-117).  Visual-word ids per image come from the
 * reference's <idx>_words files (fbow, out of scope here) in adjacency form: words of image i are
 * word_ids[word_offsets[i] .. word_offsets[i+1]), one per keypoint.
 *
 * msfm_similarity_invfile: inverted file over each image's "unique" words as math::keep_unique_vector leaves them
 *   (utils/basic_funcs.h:126-151: the ids that occur exactly once in the image, without the image's smallest and largest
 *   id), words listed by more than num_words/100 images dropped (:107-116), similarity[i][j] = number of shared words. */
int msfm_similarity_invfile(int32_t num_imgs, const int64_t *word_offsets, const int32_t *word_ids, int32_t num_words,
                            float *similarity /* [num_imgs * num_imgs] */);
/* Hypotheses of image i: the th_num_match most similar images, similarity descending (ties: lower image index; the
 * reference's std::sort leaves them unspecified).  th_num_match = 0 takes the reference's rule
 * min(max(200, num_imgs/10), num_imgs-1), capped at 500 (:166-168).  list capacity num_imgs * th_num_match. */
int msfm_pairs_similarity_topk(int32_t num_imgs, const float *similarity, int32_t th_num_match, int64_t *offsets,
                               int32_t *list);
/* Word-collision matches of two images (:239-251): pt_word_map of each image from math::keep_unique_idx_vector
 * (utils/basic_funcs.cc:380-406: word -> keypoint index, for the word groups that FOLLOW a singleton group in sorted
 * order; first element of the group, the lower keypoint index among equals), then (pt1, pt2) for every word in both maps,
 * ascending word id.  Returns the number of matches (written up to `cap`), or a negative error.  The reference keeps a
 * hypothesis when it has >= 30 such matches and F-RANSAC leaves > 20 inliers (msfm_geo_ransac on the GPU). */
int msfm_word_matches(const int32_t *words1, int32_t n1, const int32_t *words2, int32_t n2, int32_t (*matches)[2],
                      int32_t cap);
/* init_match_graph.txt */
int msfm_init_graph_write(const char *fold, int32_t num_imgs, int32_t id_last, const int64_t *offsets, const int32_t *list);
int msfm_init_graph_read(const char *fold, int32_t *num_imgs, int32_t *id_last, int64_t *offsets, int32_t offsets_cap,
                         int32_t *list, int64_t list_cap, int64_t *n_list);

#ifdef __cplusplus
}
#endif
#endif /* MSFM_STORE_H_ */
