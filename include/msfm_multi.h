/*
 * msfm_multi.h — C ABI of the multi-GPU pair scheduler (north_star subsystem 4): ONE host process drives all GPUs of a
 * box.  Exported by libmsfm_match.so next to msfm_match.h.
 *
 *   stage      every image is packed on ONE device (its owner: contiguous blocks of the ids of an upload call) from the
 *              caller's host rows, then forwarded to the other devices by NCCL broadcast over NVLink/NVSwitch straight
 *              into their descriptor tables (ncclCommInitAll, one communicator per device, grouped calls) — the
 *              "descriptor sets are broadcast once" step; nothing waits on the host
 *   match      the candidate pair list is sharded by cost (msfm_sched_shard: LPT over runs of pairs sharing the reference
 *              image), one host thread per device runs the single-GPU matcher on its shard — no data-path collective —
 *              and every launch waits, on the device, only for the transfers of the images it touches
 *   gather     the per-device match lists come back into page-locked blocks and are stitched into the caller's single
 *              msfm_result in the caller's pair order (msfm_sched_scatter)
 *
 * Reference analogue: FineMatchingGraph::BuildMatchGraph's loops (/root/reference/SfM/src/graph/fine_matching_graph.cc:
 * 58-133), which use one process, one CPU socket and OpenMP threads over the partners of an image.  A MetricSfM build
 * keeps its single C++ process and swaps msfm_create/msfm_upload_x/msfm_match_pairs for the msfm_multi_* calls below to use every GPU
 * of the box.
 */
#ifndef MSFM_MULTI_H_
#define MSFM_MULTI_H_

#include <stddef.h>
#include <stdint.h>

#include "msfm_match.h"
#include "msfm_sched.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msfm_multi msfm_multi;

typedef struct msfm_multi_config {
    int32_t n_devices;      /* >= 1 */
    const int32_t *devices; /* CUDA ordinals; NULL = 0 .. n_devices-1 */
    int32_t max_images;     /* image ids are 0 .. max_images-1 (same table layout on every device) */
    int64_t arena_rows;     /* rows of each device's replica of the packed table */
    int32_t reserved[4];    /* must be zero */
} msfm_multi_config;

/* Fails with MSFM_ERR_UNSUPPORTED when a device is not sm_100 or (n_devices > 1) NCCL cannot be loaded. */
msfm_status msfm_multi_create(const msfm_multi_config *cfg, msfm_multi **out);
msfm_status msfm_multi_destroy(msfm_multi *mm);
const char *msfm_multi_last_error(const msfm_multi *mm);
int32_t msfm_multi_device_count(const msfm_multi *mm);
/* The single-GPU context of device slot k (0 .. n_devices-1), e.g. for msfm_last_timing or msfm_download_packed. */
msfm_ctx *msfm_multi_context(msfm_multi *mm, int32_t k);

/* Stage a group of images (dense rows: 128 bytes / 128 floats per row; page-locked memory lets the copies run
 * asynchronously).  No host wait: the rows must stay valid until msfm_multi_sync() or a match call that uses them has
 * returned.  Groups staged by successive calls pipeline against matching: a match launch waits only for the images it
 * touches.  float rows are quantised like msfm_upload_f32 (the retained-float regime is single-GPU only). */
msfm_status msfm_multi_upload_u8(msfm_multi *mm, int32_t n, const int32_t *image_ids, const uint8_t *const *descs,
                                 const int32_t *rows);
msfm_status msfm_multi_upload_f32(msfm_multi *mm, int32_t n, const int32_t *image_ids, const float *const *descs,
                                  const int32_t *rows, float scale);
msfm_status msfm_multi_release_all(msfm_multi *mm);
msfm_status msfm_multi_sync(msfm_multi *mm);

/* msfm_match_pairs over all devices; `out` is filled exactly like the single-GPU call (global pair order). */
msfm_status msfm_multi_match_pairs(msfm_multi *mm, const msfm_pair *pairs, int64_t n_pairs, const msfm_params *params,
                                   msfm_result *out);

typedef struct msfm_multi_timing {
    float wall_ms;          /* host wall clock of the last msfm_multi_match_pairs call */
    float device_ms_max;    /* max over devices of msfm_timing.total_ms */
    float stitch_ms;        /* host time spent copying the lists into the caller's buffers */
    int32_t n_devices;
    int64_t int8_ops;       /* sum over devices */
    int64_t bytes_broadcast; /* bytes every device received over NCCL since creation / the last release_all */
} msfm_multi_timing;
/* per_device: optional array of n_devices msfm_timing (the last match call of each device's context). */
msfm_status msfm_multi_last_timing(const msfm_multi *mm, msfm_multi_timing *out, msfm_timing *per_device);

#ifdef __cplusplus
}
#endif
#endif /* MSFM_MULTI_H_ */
