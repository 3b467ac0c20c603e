/*
 * resource_manager.c -- frame cache + prefetch queue.  See resource_manager.h.
 * Cache behaviour follows reference resource_manager.c:39-121,195-227: a list of image_buffers in
 * creation order, one mutex per buffer held while the callback builds the frame, eviction of the
 * oldest idle buffers above a soft limit.  The limits grow with the prefetch depth so that frames
 * built ahead are not evicted before they are read.
 *
 * Lifetime rule (the reference has none and can free a buffer under a reader): a buffer is never
 * freed while it is pinned (`pins`, taken under the list lock by every lookup -- readers and prefetch
 * workers alike -- and dropped after the build) or in use (`in_use`, from the request until
 * release_image_buffer*).
 *
 * Prefetch: look-ahead frames are requested in CHUNKS of `batch` consecutive frames.  A worker
 * creates the chunk's missing buffers, holds their mutexes (a reader that wants one of them waits,
 * exactly as it would on a frame being built), and hands the whole chunk to the batch builder -- one
 * device batch on one GPU (frame_builder.c) -- or, without a batch builder, builds them one by one
 * through the frame callback.
 */
#define _GNU_SOURCE
#include "resource_manager.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BASE_MAX_UNUSED 4      /* resource_manager.c:39 */
#define BASE_MAX_TOTAL 16      /* resource_manager.c:40 */
#define MAX_BATCH 64

static pthread_mutex_t g_list_mu = PTHREAD_MUTEX_INITIALIZER;
static struct image_buffer *g_head = NULL;
static int g_count = 0;
static void (*g_data_free)(void *) = free;

/* ---- prefetch queue: chunk requests ---- */
struct prefetch_req { char *path; int first, count; image_buffer_cbr cbr; };   /* path = any frame of the clip */
#define PF_QUEUE_CAP 256
static struct prefetch_req g_q[PF_QUEUE_CAP];
static int g_q_head = 0, g_q_len = 0;
static pthread_mutex_t g_q_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_q_cv = PTHREAD_COND_INITIALIZER;
static pthread_t g_workers[32];
static int g_nworkers = 0, g_depth = 0, g_stop = 0, g_batch = 1;
static int (*g_frame_limit)(const char *) = NULL;
static image_buffer_batch_cbr g_batch_cbr = NULL;
static uint64_t g_pf_built = 0, g_pf_hits = 0, g_pf_batches = 0;

void resource_manager_set_data_free(void (*free_fn)(void *)) { g_data_free = free_fn ? free_fn : free; }

static int max_unused(void) { return BASE_MAX_UNUSED + 2 * (g_depth + g_batch); }

static struct image_buffer *find_locked(const char *path)
{
    for (struct image_buffer *b = g_head; b; b = b->next)
        if (!strcmp(b->dng_filename, path)) return b;
    return NULL;
}

static void unlink_and_free_locked(struct image_buffer *victim)
{
    struct image_buffer **pp = &g_head;
    while (*pp && *pp != victim) pp = &(*pp)->next;
    if (*pp) *pp = victim->next;
    pthread_mutex_destroy(&victim->mutex);
    free(victim->dng_filename);
    if (victim->data) g_data_free(victim->data);
    free(victim->header);
    free(victim);
    g_count--;
}

/* resource_manager.c:195-227: drop the oldest idle buffers above the soft limit.  Pinned or in-use buffers are
 * never victims; `pins` and `in_use` are only written under g_list_mu, which the caller holds. */
static void cleanup_locked(void)
{
    while (g_count > max_unused()) {
        struct image_buffer *victim = NULL;
        for (struct image_buffer *b = g_head; b; b = b->next)
            if (!b->pins && !b->in_use) { victim = b; break; }
        if (!victim) break;
        unlink_and_free_locked(victim);
    }
}

static struct image_buffer *new_buffer_locked(const char *path)
{
    cleanup_locked();
    struct image_buffer *b = calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->dng_filename = strdup(path);
    if (!b->dng_filename) { free(b); return NULL; }
    pthread_mutex_init(&b->mutex, NULL);
    struct image_buffer **pp = &g_head;
    while (*pp) pp = &(*pp)->next;
    *pp = b;
    g_count++;
    return b;
}

static void unpin(struct image_buffer *b)
{
    pthread_mutex_lock(&g_list_mu);
    b->pins--;
    pthread_mutex_unlock(&g_list_mu);
}

static struct image_buffer *get_or_create(const char *path, image_buffer_cbr cbr, int *was_created, int claim)
{
    struct image_buffer *b;
    int created = 0;
    pthread_mutex_lock(&g_list_mu);
    b = find_locked(path);
    if (!b) {
        b = new_buffer_locked(path);
        created = 1;
    } else if (b->prefetched) {
        g_pf_hits++;
        b->prefetched = 0;
    }
    if (b) {
        b->pins++;                                           /* cannot be evicted between here and the unpin below */
        if (claim) b->in_use = 1;
    }
    pthread_mutex_unlock(&g_list_mu);
    if (was_created) *was_created = created;
    if (!b) return NULL;

    pthread_mutex_lock(&b->mutex);                           /* resource_manager.c:111-118 */
    if (!b->data) cbr(b);
    pthread_mutex_unlock(&b->mutex);
    unpin(b);
    return b;
}

/* "<...>_NNNNNN.dng" -> frame number and the position of the digits (main.c:316-328) */
static int parse_frame_number(const char *path, size_t *digits_at)
{
    const char *dot = strrchr(path, '.');
    if (!dot || dot - path < 7 || strcmp(dot, ".dng")) return -1;
    for (int i = 1; i <= 6; i++) if (!isdigit((unsigned char)dot[-i])) return -1;
    *digits_at = (size_t)(dot - path) - 6;
    return atoi(dot - 6);
}

static char *path_of_frame(const char *any_frame_path, size_t digits_at, int n)
{
    char *p = strdup(any_frame_path);
    char num[16];
    if (!p) return NULL;
    snprintf(num, sizeof(num), "%06d", n);
    memcpy(p + digits_at, num, 6);
    return p;
}

/* queue the chunks that cover frames n+1 .. n+depth of the clip (chunk c = frames [c*batch, (c+1)*batch)) */
static void enqueue_following(const char *path, image_buffer_cbr cbr)
{
    if (g_depth <= 0 || g_nworkers <= 0) return;
    size_t at;
    int n = parse_frame_number(path, &at);
    if (n < 0) return;
    int limit = g_frame_limit ? g_frame_limit(path) : 0;
    int last = n + g_depth;
    if (limit > 0 && last >= limit) last = limit - 1;
    if (last <= n) return;
    pthread_mutex_lock(&g_q_mu);
    for (int c = (n + 1) / g_batch; c <= last / g_batch; c++) {
        int first = c * g_batch, count = g_batch;
        if (first <= n) { count -= n + 1 - first; first = n + 1; }           /* the rest of the requested frame's own chunk */
        if (limit > 0 && first + count > limit) count = limit - first;
        if (count <= 0) continue;
        /* skip chunks whose frames are all cached (or being built) already, and chunks already queued */
        int missing = 0;
        pthread_mutex_lock(&g_list_mu);
        for (int k = 0; k < count && !missing; k++) {
            char *p = path_of_frame(path, at, first + k);
            if (p) { missing = find_locked(p) == NULL; free(p); }
        }
        pthread_mutex_unlock(&g_list_mu);
        if (!missing) continue;
        int dup = 0;
        for (int i = 0; i < g_q_len && !dup; i++) {
            const struct prefetch_req *r = &g_q[(g_q_head + i) % PF_QUEUE_CAP];
            dup = r->first <= first && first + count <= r->first + r->count && !strncmp(r->path, path, at);
        }
        if (dup || g_q_len == PF_QUEUE_CAP) continue;
        struct prefetch_req *r = &g_q[(g_q_head + g_q_len) % PF_QUEUE_CAP];
        r->path = strdup(path);
        if (!r->path) continue;
        r->first = first; r->count = count; r->cbr = cbr;
        g_q_len++;
    }
    pthread_cond_broadcast(&g_q_cv);
    pthread_mutex_unlock(&g_q_mu);
}

/* Build the missing frames of one chunk.  New buffers are created, pinned and locked under the list lock (nobody
 * else can hold a new buffer's mutex yet), so a reader that asks for one of them meanwhile waits on its mutex. */
static void build_chunk(const struct prefetch_req *r)
{
    size_t at;
    if (parse_frame_number(r->path, &at) < 0) return;
    struct image_buffer *bufs[MAX_BATCH];
    int n = 0;
    pthread_mutex_lock(&g_list_mu);
    for (int k = 0; k < r->count && n < MAX_BATCH; k++) {
        char *p = path_of_frame(r->path, at, r->first + k);
        if (!p) continue;
        if (!find_locked(p)) {
            struct image_buffer *b = new_buffer_locked(p);
            if (b) {
                b->pins++;
                b->prefetched = 1;
                pthread_mutex_lock(&b->mutex);
                bufs[n++] = b;
            }
        }
        free(p);
    }
    pthread_mutex_unlock(&g_list_mu);
    if (!n) return;
    if (g_batch_cbr && n > 1) {
        g_batch_cbr(bufs, n);
        __sync_fetch_and_add(&g_pf_batches, 1);
    } else {
        for (int k = 0; k < n; k++) r->cbr(bufs[k]);
    }
    for (int k = 0; k < n; k++) {
        if (bufs[k]->data) __sync_fetch_and_add(&g_pf_built, 1);
        pthread_mutex_unlock(&bufs[k]->mutex);
    }
    pthread_mutex_lock(&g_list_mu);
    for (int k = 0; k < n; k++) bufs[k]->pins--;
    pthread_mutex_unlock(&g_list_mu);
}

static void *prefetch_worker(void *arg)
{
    (void)arg;
    for (;;) {
        pthread_mutex_lock(&g_q_mu);
        while (!g_stop && g_q_len == 0) pthread_cond_wait(&g_q_cv, &g_q_mu);
        if (g_stop) { pthread_mutex_unlock(&g_q_mu); return NULL; }
        struct prefetch_req r = g_q[g_q_head];
        g_q_head = (g_q_head + 1) % PF_QUEUE_CAP;
        g_q_len--;
        pthread_mutex_unlock(&g_q_mu);
        build_chunk(&r);
        free(r.path);
    }
}

struct image_buffer *get_or_create_image_buffer(const char *path, image_buffer_cbr new_buffer_cbr, int *was_created)
{
    enqueue_following(path, new_buffer_cbr);                 /* start the look-ahead before we block on this frame */
    return get_or_create(path, new_buffer_cbr, was_created, 1);
}

void release_image_buffer(struct image_buffer *image_buffer)
{
    pthread_mutex_lock(&g_list_mu);
    image_buffer->in_use = 0;
    pthread_mutex_unlock(&g_list_mu);
}

void release_image_buffer_by_path(const char *path)
{
    pthread_mutex_lock(&g_list_mu);
    struct image_buffer *b = find_locked(path);
    if (b) b->in_use = 0;
    pthread_mutex_unlock(&g_list_mu);
}

void free_all_image_buffers(void)
{
    pthread_mutex_lock(&g_list_mu);
    while (g_head) unlink_and_free_locked(g_head);
    pthread_mutex_unlock(&g_list_mu);
}

int get_image_buffer_count(void) { return g_count; }

void resource_manager_set_prefetch(int depth, int workers, int (*frame_limit)(const char *))
{
    resource_manager_shutdown();
    g_frame_limit = frame_limit;
    g_depth = depth < 0 ? 0 : (depth > 256 ? 256 : depth);
    if (g_depth == 0) return;
    if (workers <= 0) {
        workers = (g_depth + g_batch - 1) / g_batch;         /* one worker per chunk in flight */
        if (g_batch == 1 && workers > 8) workers = 8;
        if (workers < 2 && g_batch > 1) workers = 2;
    }
    if (workers > 32) workers = 32;
    g_stop = 0;
    for (g_nworkers = 0; g_nworkers < workers; g_nworkers++)
        if (pthread_create(&g_workers[g_nworkers], NULL, prefetch_worker, NULL)) break;
}

void resource_manager_set_batch_builder(image_buffer_batch_cbr batch_cbr, int batch)
{
    g_batch_cbr = batch_cbr;
    g_batch = batch < 1 ? 1 : (batch > MAX_BATCH ? MAX_BATCH : batch);
}

void resource_manager_shutdown(void)
{
    pthread_mutex_lock(&g_q_mu);
    g_stop = 1;
    pthread_cond_broadcast(&g_q_cv);
    pthread_mutex_unlock(&g_q_mu);
    for (int i = 0; i < g_nworkers; i++) pthread_join(g_workers[i], NULL);
    g_nworkers = 0;
    pthread_mutex_lock(&g_q_mu);
    while (g_q_len) { free(g_q[g_q_head].path); g_q_head = (g_q_head + 1) % PF_QUEUE_CAP; g_q_len--; }
    pthread_mutex_unlock(&g_q_mu);
    g_depth = 0;
}

void resource_manager_prefetch_stats(uint64_t *built, uint64_t *hits)
{
    if (built) *built = g_pf_built;
    if (hits) *hits = g_pf_hits;
}

uint64_t resource_manager_prefetch_batches(void) { return g_pf_batches; }
