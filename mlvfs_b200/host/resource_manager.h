/*
 * resource_manager.h -- frame cache + --prefetch queue (host side, C99).
 *
 * Keeps the reference's frame-request API (resource_manager.h:33-51): the FUSE read handler calls
 *     get_or_create_image_buffer(path, &process_frame, &was_created)          main.c:1460
 * and release_image_buffer_by_path(path) on release (main.c:1634-1644), exactly as before.
 * New: the --prefetch queue documented in the reference README (README.md:42) but absent from its
 * sources.  After a frame of a clip is requested, the next `depth` frames are built ahead of time by
 * worker threads through the same callback, so that their GPU work (mlvb_submit slots) overlaps the
 * reader's memcpy of the current frame.
 */
#ifndef MLVB_HOST_RESOURCE_MANAGER_H
#define MLVB_HOST_RESOURCE_MANAGER_H

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct image_buffer {                 /* reference resource_manager.h:33-43, same field order */
    struct image_buffer *next;
    char *dng_filename;
    size_t header_size;
    size_t size;
    uint8_t *header;
    uint16_t *data;
    pthread_mutex_t mutex;
    int in_use;
    int prefetched;                   /* appended (front-ends never allocate this struct): built by the prefetch queue */
    int pins;                         /* appended: lookups in progress; a pinned buffer is never evicted */
};

typedef int (*image_buffer_cbr)(struct image_buffer *);
/* builds `n` frames of one clip at once (consecutive frame numbers, every buffer locked by the caller) */
typedef int (*image_buffer_batch_cbr)(struct image_buffer **, int n);

struct image_buffer *get_or_create_image_buffer(const char *path, image_buffer_cbr new_buffer_cbr, int *was_created);
void release_image_buffer_by_path(const char *path);
void release_image_buffer(struct image_buffer *image_buffer);
void free_all_image_buffers(void);
int  get_image_buffer_count(void);

/* `data` may come from mlvb_host_alloc (pinned); the cache frees it with this hook (default free). */
void resource_manager_set_data_free(void (*free_fn)(void *));

/* --prefetch=N: build up to `depth` following frames ahead with `workers` threads (0 disables).
 * `frame_limit(path)` returns the clip's frame count for a virtual DNG path (or <= 0 if unknown). */
void resource_manager_set_prefetch(int depth, int workers, int (*frame_limit)(const char *dng_path));
/* Look-ahead frames are requested in chunks of `batch` consecutive frames and handed to `batch_cbr` as one call
 * (one device batch on one GPU; frame_builder.c: process_frame_batch).  Call before resource_manager_set_prefetch.
 * batch <= 1 or batch_cbr == NULL: frame-by-frame look-ahead through the frame callback. */
void resource_manager_set_batch_builder(image_buffer_batch_cbr batch_cbr, int batch);
void resource_manager_shutdown(void);
/* statistics for tests / the frame server: frames built by prefetch workers, cache hits on them */
void resource_manager_prefetch_stats(uint64_t *built, uint64_t *hits);
uint64_t resource_manager_prefetch_batches(void);

#ifdef __cplusplus
}
#endif
#endif
