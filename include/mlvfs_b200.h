/*
 * mlvfs_b200.h -- C ABI of libmlvfs_b200.so, the B200 (sm_100a) implementation of MLVFS's
 * per-frame raw path (MLV video frame -> DNG pixel data).
 *
 * Two groups of entry points:
 *
 *  (1) DROP-IN SYMBOLS.  The exact C signatures the reference's front-ends link against, so that
 *      main.c (FUSE), gif.c and win/mlvfs-pfm.cpp compile and link unchanged when dng.o / cs.o /
 *      stripes.o / hdr.o / patternnoise.o are replaced by this library.  They take HOST buffers,
 *      work in place, and keep the reference's error behaviour (sizes return 0 on failure, void
 *      functions return silently).  Each declaration cites the reference interface it replaces.
 *
 *  (2) mlvb_* ENTRY POINTS.  The fused per-frame call the host-side frame builder
 *      (mlvfs_b200/host/frame_builder.c, our process_frame) uses: raw VIDF payload in, finished
 *      16-bit frame out, options passed per call, per-clip state owned by the library.  A
 *      submit/wait pair feeds the --prefetch queue; a device-resident batch form is what
 *      bench.py times for the roofline.
 *
 * The library never falls back to the CPU: without a CUDA device every entry point fails
 * (mlvb_* return MLVB_ERR_CUDA, drop-in functions return 0 / leave the buffer untouched and
 * print to stderr).
 */
#ifndef MLVFS_B200_H
#define MLVFS_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>

#include "mlvb_mlv_format.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status codes ---------- */
#define MLVB_OK                 0
#define MLVB_ERR_CUDA          (-1)   /* no device / CUDA runtime error (message on stderr)        */
#define MLVB_ERR_ARG           (-2)   /* bad argument (NULL, black level > 16384, ...)             */
#define MLVB_ERR_UNSUPPORTED   (-3)   /* payload kind not handled on the GPU path (e.g. LZMA)       */
#define MLVB_ERR_NOMEM         (-4)
#define MLVB_ERR_NOT_DUAL_ISO  (-5)

/* ------------------------------------------------------------------ (2) mlvb_* API -------- */

typedef struct mlvb_context mlvb_context;   /* one per GPU: streams, frame slots, LUTs, clip state */

/* Immutable per-call snapshot of the processing fields of `struct mlvfs` (reference mlvfs.h:37-46).
 * The web GUI mutates that struct while frames are in flight (webgui.c:298-336), so the frame
 * builder copies it once at entry and passes the copy down. */
typedef struct mlvb_options {
    int chroma_smooth;             /* 0, 2, 3, 5          --cs2x2 / --cs3x3 / --cs5x5              */
    int fix_bad_pixels;            /* 0, 1, 2             --bad-pix / --really-bad-pix             */
    int fix_stripes;               /* 0, 1                --stripes                                */
    int dual_iso;                  /* 0, 1 preview, 2 full --dual-iso-preview / --dual-iso          */
    int hdr_interpolation_method;  /* 0 AMaZE+edge, 1 mean23                                        */
    int hdr_no_fullres;
    int hdr_no_alias_map;
    int fix_pattern_noise;         /* --fix-pattern-noise                                          */
    int deflicker;                 /* --deflicker=<target>                                         */
} mlvb_options;

/* What process_frame needs back to (re)build the DNG header (main.c:944, 961-965). */
typedef struct mlvb_frame_result {
    int     status;                /* MLVB_OK or MLVB_ERR_*                                        */
    int     is_dual_iso;           /* 1: frame was converted, black/white below are already x4      */
    int32_t black_level;
    int32_t white_level;
    int32_t exposure_bias[2];      /* deflicker result (main.c:895-906), else copied from input     */
} mlvb_frame_result;

typedef int64_t mlvb_ticket;       /* >= 0: in-flight frame; < 0: MLVB_ERR_*                        */

/* Create a context on CUDA device `device` with `slots` frames in flight (0 = default 4). */
int  mlvb_context_create(int device, int slots, mlvb_context **out);
void mlvb_context_destroy(mlvb_context *ctx);
/* Process-wide context used by the drop-in symbols; device from $MLVB_DEVICE (default 0). */
mlvb_context *mlvb_default_context(void);
int  mlvb_device_count(void);

/* Pinned host memory for image_buffer->data / payload staging (replaces malloc/free at
 * main.c:931 and resource_manager.c:143-146; pageable buffers also work, through a staging copy). */
void *mlvb_host_alloc(size_t bytes);
void  mlvb_host_free(void *p);
/* mlvb_host_free keeps blocks for reuse (a pool of up to $MLVB_PIN_POOL_MB MiB, default 4096: page-locking is a
 * millisecond-scale driver call and the frame cache needs a frame-sized buffer per frame); this returns the idle
 * blocks to the system. */
void  mlvb_host_pool_trim(void);

/* Build one frame.  `payload` is the VIDF payload exactly as stored in the MLV (packed bits, or
 * uint32 size + LJ92 stream when file_hdr.videoClass has MLVB_VIDEO_CLASS_FLAG_LJ92);
 * `mlv_filename` identifies the clip for per-clip state (stripe coefficients; the bad-pixel map is
 * keyed by fileGuid + aggressiveness like cs.c:233-254).  `dst` receives xRes*yRes uint16. */
int mlvb_process_frame(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                       const mlvb_options *opts, const char *mlv_filename, uint16_t *dst, mlvb_frame_result *res);

/* Asynchronous form used by the prefetch queue: submit returns at once (it blocks only while all
 * slots are busy); payload and dst must stay valid until mlvb_wait returns. */
mlvb_ticket mlvb_submit(mlvb_context *ctx, const struct frame_headers *hdr, const void *payload, size_t payload_bytes,
                        const mlvb_options *opts, const char *mlv_filename, uint16_t *dst);
int mlvb_wait(mlvb_context *ctx, mlvb_ticket ticket, mlvb_frame_result *res);

/* Host batch -- what the --prefetch queue calls with its look-ahead frames: `nframes` frames of ONE clip, payloads
 * and destinations in host memory (pinned for asynchronous copies; pageable memory works but serialises).  Frames
 * whose results do not depend on per-frame statistics (no dual ISO, no deflicker) and that share one shape run as
 * one device batch -- one pass of the fused kernels over all of them, copies overlapped with other batches in
 * flight; anything else is pipelined frame by frame over the context's slots.  Blocks until every frame is in its
 * destination.  results may be NULL.  Returns MLVB_OK or the last error (per-frame status in results[]). */
int mlvb_process_frames(mlvb_context *ctx, int nframes, const struct frame_headers *hdrs, const void *const *payloads,
                        const size_t *payload_bytes, const mlvb_options *opts, const char *mlv_filename, uint16_t *const *dsts,
                        mlvb_frame_result *results);

/* Device-resident batch: `nframes` payloads of one clip already in HBM (payload_stride bytes apart,
 * 16-byte aligned) -> `nframes` finished frames in HBM (out_stride_px uint16 apart).  Enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = the context's own stream); does not synchronise unless the
 * clip's per-clip state has to be created from frame 0 of the batch. */
int mlvb_process_batch_device(mlvb_context *ctx, const struct frame_headers *hdr, const mlvb_options *opts,
                              const char *mlv_filename, const void *d_payload, size_t payload_stride,
                              size_t payload_bytes, uint16_t *d_out, size_t out_stride_px, int nframes,
                              void *cuda_stream);

/* Drop all per-clip state (stripe coefficients, bad-pixel maps, focus maps). */
void mlvb_reset_clip_state(mlvb_context *ctx);
/* Seed of the glibc-compatible rand() stream used for the stripes dither (a fresh process = 1). */
void mlvb_seed_dither(mlvb_context *ctx, unsigned seed);
/* Introspection for tests: stripe coefficients / bad-pixel list of a clip. Return count or <0. */
int mlvb_get_stripes(mlvb_context *ctx, const char *mlv_filename, int *needed, int coef[8]);
int mlvb_get_bad_pixels(mlvb_context *ctx, uint64_t file_guid, int aggressive, int *xy, int cap);
/* Number of kernels launched by this context so far (bench.py's gpu_launches). */
uint64_t mlvb_launch_count(mlvb_context *ctx);
/* Introspection for tests: how often a given kernel path was taken.  which: 0 = fused single-ISO strip kernel
 * (per-frame and small batches), 1 = fused single-ISO wide kernel (large batches), 2 = host batches that ran as one
 * device batch (mlvb_process_frames). */
uint64_t mlvb_path_count(mlvb_context *ctx, int which);
/* Per-stage device timing for the roofline report: between begin and end every stage of the
 * batch / per-frame pipeline is bracketed by CUDA events on its own stream.  Stage ids:
 * 0 unpack, 1 bad/focus-pixel fix, 2 chroma smoothing (+fused stripes), 3 stripes apply,
 * 4 pattern noise, 5 dual ISO, 6 LJ92.  Single caller only. */
void mlvb_profile_begin(mlvb_context *ctx);
int  mlvb_profile_end(mlvb_context *ctx, float *ms_per_stage, int *spans_per_stage, int nstages);

/* ------------------------------------------------------------------ (1) drop-in symbols --- */

/* reference dng.h:31-32 (dng.c:854-872, 879-882) */
size_t dng_get_image_data(struct frame_headers *frame_headers, uint16_t *packed_bits, uint8_t *output_buffer,
                          off_t offset, size_t max_size);
size_t dng_get_image_size(struct frame_headers *frame_headers);

/* reference mlvfs.h:80 (main.c:569-706): read the VIDF payload from `file` and decode it */
size_t get_image_data(struct frame_headers *frame_headers, FILE *file, uint8_t *output_buffer, off_t offset,
                      size_t max_size);

/* reference mlvfs.h:90-92 (main.c:128-196): host copies of the EV tables */
double *get_raw2evf(int black);
int    *get_raw2ev(int black);
int    *get_ev2raw(void);

/* reference cs.h:27-30 (cs.c:49-84, 220-331, 440-503, 404-419) */
void chroma_smooth(struct frame_headers *frame_headers, uint16_t *image_data, int method);
void fix_bad_pixels(struct frame_headers *frame_headers, uint16_t *image_data, int aggressive, int dual_iso);
void fix_focus_pixels(struct frame_headers *frame_headers, uint16_t *image_data, int dual_iso);
void free_focus_pixel_maps(void);

/* reference stripes.h:30-43 (stripes.c:29-266) */
struct stripes_correction {
    struct stripes_correction *next;
    char *mlv_filename;
    int correction_needed;
    int coeffficients[8];          /* sic: the reference spells it with three f */
};
struct stripes_correction *stripes_get_correction(const char *mlv_filename);
struct stripes_correction *stripes_new_correction(const char *mlv_filename);
void stripes_free_corrections(void);
void stripes_compute_correction(struct frame_headers *frame_headers, struct stripes_correction *correction,
                                uint16_t *image_data, off_t offset, size_t size);
void stripes_apply_correction(struct frame_headers *frame_headers, struct stripes_correction *correction,
                              uint16_t *image_data, off_t offset, size_t size);

/* reference patternnoise.h:16 (patternnoise.c:357-380) */
void fix_pattern_noise(int16_t *raw, int w, int h, int white, int debug_flags);

/* reference hdr.h:27-28 (hdr.c:40-227, 1932-1957); return 1 = converted (and black/white x4 in
 * frame_headers), 0 = not dual ISO / failed */
int hdr_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, off_t offset, size_t max_size);
int cr2hdr20_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, int interp_method, int fullres,
                          int use_alias_map, int chroma_smooth, int fix_bad_pixels_mode);

/* reference amaze_demosaic_RT.c:113-120 (declared inside hdr.c:1028-1035, called at hdr.c:1040): AMaZE demosaic
 * of a float RGGB mosaic.  rawData / red / green / blue are arrays of `winh` row pointers, every row
 * winw + 16 floats (hdr.c:967-975).  Only winx = winy = 0 (the reference's own call) and winw % 4 == 0. */
void amaze_demosaic_RT(float **rawData, float **red, float **green, float **blue, int winx, int winy, int winw,
                       int winh);

#ifdef __cplusplus
}
#endif
#endif
