/*
 * resource_manager.c -- frame cache + prefetch queue.  See resource_manager.h.
 * Cache behaviour follows reference resource_manager.c:39-121,195-227: a list of image_buffers in
 * creation order, one mutex per buffer held while the callback builds the frame, eviction of the
 * oldest unused buffers above a soft limit and of the oldest buffers, used or not, above a hard limit.
 * The limits grow with the prefetch depth so that frames built ahead are not evicted before they
 * are read.
 */
#define _GNU_SOURCE
#include "resource_manager.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BASE_MAX_UNUSED 4      /* resource_manager.c:39 */
#define BASE_MAX_TOTAL 16      /* resource_manager.c:40 */

static pthread_mutex_t g_list_mu = PTHREAD_MUTEX_INITIALIZER;
static struct image_buffer *g_head = NULL;
static int g_count = 0;
static void (*g_data_free)(void *) = free;

/* ---- prefetch queue ---- */
struct prefetch_req { char *path; image_buffer_cbr cbr; };
#define PF_QUEUE_CAP 256
static struct prefetch_req g_q[PF_QUEUE_CAP];
static int g_q_head = 0, g_q_len = 0;
static pthread_mutex_t g_q_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_q_cv = PTHREAD_COND_INITIALIZER;
static pthread_t g_workers[32];
static int g_nworkers = 0, g_depth = 0, g_stop = 0;
static int (*g_frame_limit)(const char *) = NULL;
static uint64_t g_pf_built = 0, g_pf_hits = 0;

void resource_manager_set_data_free(void (*free_fn)(void *)) { g_data_free = free_fn ? free_fn : free; }

static int max_unused(void) { return BASE_MAX_UNUSED + 2 * g_depth; }
static int max_total(void) { return BASE_MAX_TOTAL + 4 * g_depth; }

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

/* resource_manager.c:195-227.  A buffer whose mutex is held is being built right now: skip it. */
static void cleanup_locked(void)
{
    while (g_count > max_unused()) {
        struct image_buffer *victim = NULL;
        for (struct image_buffer *b = g_head; b; b = b->next) {
            if (pthread_mutex_trylock(&b->mutex) != 0) continue;
            int idle = !b->in_use && b->data != NULL;
            pthread_mutex_unlock(&b->mutex);
            if (idle) { victim = b; break; }
        }
        if (!victim) break;
        unlink_and_free_locked(victim);
    }
    while (g_count > max_total() && g_head) {
        struct image_buffer *victim = NULL;
        for (struct image_buffer *b = g_head; b; b = b->next) {
            if (pthread_mutex_trylock(&b->mutex) != 0) continue;
            pthread_mutex_unlock(&b->mutex);
            victim = b;
            break;
        }
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

static struct image_buffer *get_or_create(const char *path, image_buffer_cbr cbr, int *was_created, int claim, int prefetch)
{
    struct image_buffer *b;
    int created = 0;
    pthread_mutex_lock(&g_list_mu);
    b = find_locked(path);
    if (!b) {
        b = new_buffer_locked(path);
        created = 1;
    } else if (!prefetch && b->prefetched) {
        g_pf_hits++;
        b->prefetched = 0;
    }
    if (b && claim) b->in_use = 1;
    if (b && created && prefetch) b->prefetched = 1;
    pthread_mutex_unlock(&g_list_mu);
    if (was_created) *was_created = created;
    if (!b) return NULL;

    pthread_mutex_lock(&b->mutex);                           /* resource_manager.c:111-118 */
    if (!b->data) {
        cbr(b);
        if (prefetch) { __sync_fetch_and_add(&g_pf_built, 1); }
    }
    pthread_mutex_unlock(&b->mutex);
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

static void enqueue_following(const char *path, image_buffer_cbr cbr)
{
    if (g_depth <= 0 || g_nworkers <= 0) return;
    size_t at;
    int n = parse_frame_number(path, &at);
    if (n < 0) return;
    int limit = g_frame_limit ? g_frame_limit(path) : 0;
    pthread_mutex_lock(&g_q_mu);
    for (int k = 1; k <= g_depth; k++) {
        if (limit > 0 && n + k >= limit) break;
        if (g_q_len == PF_QUEUE_CAP) break;
        char *p = strdup(path);
        char num[16];
        snprintf(num, sizeof(num), "%06d", n + k);
        memcpy(p + at, num, 6);
        int dup = 0;
        for (int i = 0; i < g_q_len && !dup; i++) dup = !strcmp(g_q[(g_q_head + i) % PF_QUEUE_CAP].path, p);
        if (dup) { free(p); continue; }
        struct prefetch_req *r = &g_q[(g_q_head + g_q_len) % PF_QUEUE_CAP];
        r->path = p; r->cbr = cbr;
        g_q_len++;
    }
    pthread_cond_broadcast(&g_q_cv);
    pthread_mutex_unlock(&g_q_mu);
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
        get_or_create(r.path, r.cbr, NULL, 0, 1);
        free(r.path);
    }
}

struct image_buffer *get_or_create_image_buffer(const char *path, image_buffer_cbr new_buffer_cbr, int *was_created)
{
    enqueue_following(path, new_buffer_cbr);                 /* start the look-ahead before we block on this frame */
    return get_or_create(path, new_buffer_cbr, was_created, 1, 0);
}

void release_image_buffer(struct image_buffer *image_buffer)
{
    pthread_mutex_lock(&image_buffer->mutex);
    image_buffer->in_use = 0;
    pthread_mutex_unlock(&image_buffer->mutex);
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
    g_depth = depth < 0 ? 0 : (depth > 64 ? 64 : depth);
    if (g_depth == 0) return;
    if (workers <= 0) workers = g_depth < 8 ? g_depth : 8;
    if (workers > 32) workers = 32;
    g_stop = 0;
    for (g_nworkers = 0; g_nworkers < workers; g_nworkers++)
        if (pthread_create(&g_workers[g_nworkers], NULL, prefetch_worker, NULL)) break;
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
