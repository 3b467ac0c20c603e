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
void frame_builder_set_context(mlvb_context *ctx);                      /* default: mlvb_default_context() */
void frame_builder_set_header_writer(dng_header_writer fn);             /* default: none (header zeroed) */

/* The callback for get_or_create_image_buffer (same type as the reference's process_frame). */
int process_frame(struct image_buffer *image_buffer);
/* frame count of the clip a virtual DNG path belongs to (prefetch limit) */
int frame_builder_frame_limit(const char *dng_path);

#define MLVB_DNG_HEADER_SIZE 65536    /* dng.c HEADER_SIZE (dng.c:800-803) */

#ifdef __cplusplus
}
#endif
#endif
