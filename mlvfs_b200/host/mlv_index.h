/*
 * mlv_index.h -- per-clip block index + frame-header cache (host side, C99).
 *
 * Replaces the reference's per-frame header walk: mlv_get_frame_headers (main.c:429-558) re-opens
 * every chunk (resource_manager.c:285, index.c:368), re-reads <clip>.IDX (index.c:458) and walks all
 * xref entries up to frame n for EVERY frame -- O(n) freads per frame, which caps the host at ~80 fps
 * (SURVEY.md 8(f) rank 1).  Here each clip is scanned once; every frame's `struct frame_headers` is
 * materialised up front, payloads are fetched with pread() on descriptors that stay open.
 *
 * Semantics kept from the reference: blocks are ordered by timestamp with a stable sort, MLVI counts
 * as time 0, NULL blocks are ignored (index.c:216-341); frame n is the n-th VIDF in that order and
 * carries the latest MLVI/RTCI/IDNT/RAWI/EXPO/LENS/WBAL seen before it (main.c:457-542).
 */
#ifndef MLVB_HOST_MLV_INDEX_H
#define MLVB_HOST_MLV_INDEX_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#include "mlvfs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

struct mlv_clip;   /* opaque, cached by path for the life of the process */

/* Open (or fetch from the cache) the clip at `mlv_path` ("x.MLV" plus chunks "x.M00".."x.M98"). */
struct mlv_clip *mlv_clip_open(const char *mlv_path);
int  mlv_clip_frame_count(const struct mlv_clip *clip);
/* Copy frame `index`'s headers; returns 1 on success, 0 if out of range / no RAWI (main.c:544-557). */
int  mlv_clip_frame_headers(const struct mlv_clip *clip, int index, struct frame_headers *out);
/* Size in bytes of the frame's VIDF payload (blockSize - header - frameSpace). */
size_t mlv_clip_payload_size(const struct frame_headers *hdr);
/* pread() the VIDF payload into dst (thread-safe); returns bytes read or -1. */
ssize_t mlv_clip_read_payload(const struct mlv_clip *clip, const struct frame_headers *hdr, void *dst, size_t cap);
void mlv_clip_close_all(void);

/* reference-named entry points (mlvfs.h:78-79) on top of the cache */
int mlv_get_frame_headers(const char *mlv_filename, int index, struct frame_headers *frame_headers);
int mlv_get_frame_count(const char *real_path);

#ifdef __cplusplus
}
#endif
#endif
