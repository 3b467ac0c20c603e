/*
 * frame_builder.h -- our process_frame: the callback the frame cache invokes to build one virtual
 * DNG's pixel data (reference main.c:908-1005), re-shaped around the GPU library:
 *
 *   resolve "<clip>.MLV/<clip>_NNNNNN.dng" -> cached headers (mlv_index) -> pread the VIDF payload into
 *   pinned memory -> mlvb_process_frame (all pixel stages on the GPU, one H2D + one D2H) -> DNG header.
 *
 * The option block is snapshotted once per frame (the web GUI may change it concurrently,
 * webgui.c:298-336).  The DNG header writer (dng.c:612-789) is outside the hot path (SURVEY 8(f) rank
 * 2); the front-end plugs its own in through frame_builder_set_header_writer.
 */
#ifndef MLVB_HOST_FRAME_BUILDER_H
#define MLVB_HOST_FRAME_BUILDER_H

#include "mlvfs_b200.h"
#include "resource_manager.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the processing fields of the reference's `struct mlvfs` (mlvfs.h:32-48) plus the MLV directory */
struct frame_builder_config {
    const char *mlv_path;          /* --mlv_dir */
    mlvb_options options;
    double fps;                    /* --fps override handed to the header writer */
};

/* size_t dng_get_header_data(hdrs, out, offset, max_size, fps_override, mlv_basename)  (dng.h:29) */
typedef size_t (*dng_header_writer)(struct frame_headers *, uint8_t *, off_t, size_t, double, char *);

void frame_builder_configure(const struct frame_builder_config *cfg);   /* copies; call again to change options */
void frame_builder_set_context(mlvb_context *ctx);                      /* one caller-owned context (default: mlvb_default_context()) */
void frame_builder_set_header_writer(dng_header_writer fn);             /* default: dng_get_header_data below; NULL: header left zeroed */

/* Multi-GPU: create one context per GPU (ngpus <= 0: every visible device) with `slots` frames in flight each, and
 * deal a clip's frames to them in chunks of `chunk` consecutive frames: frame n is built on GPU (n / chunk) mod G.
 * Use the same `chunk` for resource_manager_set_batch_builder so that a look-ahead chunk is one device batch on one
 * GPU.  Returns the number of contexts (0 on failure).  With $MLVB_SHARE_DEVICES set, more contexts than devices are
 * allowed (context i on device i mod #devices): the dispatcher can then be exercised on a one-GPU machine. */
int  frame_builder_use_gpus(int ngpus, int slots, int chunk);
int  frame_builder_gpu_count(void);
mlvb_context *frame_builder_context(int i);
void frame_builder_shutdown(void);                                      /* destroys the contexts frame_builder_use_gpus made */
/* Forget which clips have been primed and drop the per-clip state of every context (call this instead of
 * mlvb_reset_clip_state / free_focus_pixel_maps when the frame builder is in use). */
void frame_builder_reset_clip_state(void);

/* where the builder threads' time went, summed over threads: reading payloads into pinned memory, inside the GPU
 * library (copies + kernels + waiting), writing DNG headers, priming clips */
struct frame_builder_stats { uint64_t read_ns, gpu_ns, header_ns, prime_ns, frames, gpu_calls; };
void frame_builder_get_stats(struct frame_builder_stats *out);
void frame_builder_reset_stats(void);

/* The callback for get_or_create_image_buffer (same type as the reference's process_frame). */
int process_frame(struct image_buffer *image_buffer);
/* The batch builder for resource_manager_set_batch_builder: a chunk of look-ahead frames as one device batch. */
int process_frame_batch(struct image_buffer **image_buffers, int n);

/* reference dng.h:29-30 (dng.c:612-803): the CinemaDNG header of a frame, byte-identical to the reference's */
size_t dng_get_header_data(struct frame_headers *frame_headers, uint8_t *output_buffer, off_t offset, size_t max_size,
                           double fps_override, char *mlv_basename);
size_t dng_get_header_size(void);
/* frame count of the clip a virtual DNG path belongs to (prefetch limit) */
int frame_builder_frame_limit(const char *dng_path);

#define MLVB_DNG_HEADER_SIZE 65536    /* dng.c HEADER_SIZE (dng.c:800-803) */

#ifdef __cplusplus
}
#endif
#endif
