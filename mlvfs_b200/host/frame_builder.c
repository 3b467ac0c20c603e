/* frame_builder.c -- our process_frame.  See frame_builder.h; stage order and header handling follow
 * reference main.c:908-1005, the pixel stages themselves run inside libmlvfs_b200.so.
 *
 * Multi-GPU: one mlvb_context per GPU (frame_builder_use_gpus).  A clip's frames are dealt to the GPUs in
 * chunks of `batch` consecutive frames -- frame n goes to GPU (n / batch) mod G -- so that a look-ahead chunk of
 * the prefetch queue is one device batch on one GPU.  Per-clip state (stripe coefficients, bad-pixel map, dual-ISO
 * table white level) is derived from frame 0: the first time a clip is seen, frame 0 is pushed through EVERY GPU's
 * context, in the same clip order on all of them, so that all contexts hold identical state (same input, same
 * dither stream) and the output does not depend on which GPU builds a frame or on request order. */
#define _GNU_SOURCE
#include "frame_builder.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <time.h>

#include "mlv_index.h"

#define MAX_GPUS 16
#define MAX_BATCH 64

static pthread_mutex_t g_cfg_mu = PTHREAD_MUTEX_INITIALIZER;
static struct frame_builder_config g_cfg;
static char g_mlv_dir[4096];
static mlvb_context *g_ctx[MAX_GPUS];
static int g_nctx = 0, g_owns_ctx = 0, g_chunk = 1;
static dng_header_writer g_header_writer = dng_get_header_data;

/* where a frame's host time goes (summed over all threads; frame_builder_get_stats) */
static uint64_t g_ns_read = 0, g_ns_gpu = 0, g_ns_header = 0, g_ns_prime = 0, g_n_frames = 0, g_n_calls = 0;

static uint64_t now_ns(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

void frame_builder_reset_stats(void) { g_ns_read = g_ns_gpu = g_ns_header = g_ns_prime = g_n_frames = g_n_calls = 0; }

void frame_builder_get_stats(struct frame_builder_stats *out)
{
    out->read_ns = g_ns_read; out->gpu_ns = g_ns_gpu; out->header_ns = g_ns_header; out->prime_ns = g_ns_prime;
    out->frames = g_n_frames; out->gpu_calls = g_n_calls;
}

void frame_builder_configure(const struct frame_builder_config *cfg)
{
    pthread_mutex_lock(&g_cfg_mu);
    g_cfg = *cfg;
    snprintf(g_mlv_dir, sizeof(g_mlv_dir), "%s", cfg->mlv_path ? cfg->mlv_path : ".");
    g_cfg.mlv_path = g_mlv_dir;
    pthread_mutex_unlock(&g_cfg_mu);
}

static void drop_contexts(void)
{
    if (g_owns_ctx)
        for (int i = 0; i < g_nctx; i++) mlvb_context_destroy(g_ctx[i]);
    g_nctx = 0;
    g_owns_ctx = 0;
}

void frame_builder_set_context(mlvb_context *ctx)
{
    drop_contexts();
    if (ctx) { g_ctx[0] = ctx; g_nctx = 1; }
}

void frame_builder_set_header_writer(dng_header_writer fn) { g_header_writer = fn; }

int frame_builder_use_gpus(int ngpus, int slots, int chunk)
{
    drop_contexts();
    const int have = mlvb_device_count();
    if (have <= 0) {
        fprintf(stderr, "frame_builder: no CUDA device -- libmlvfs_b200 has no CPU path\n");
        return 0;
    }
    if (ngpus <= 0) ngpus = have;
    if (ngpus > MAX_GPUS) ngpus = MAX_GPUS;
    /* more contexts than devices only when asked for explicitly (tests of the dispatcher on a one-GPU box) */
    if (ngpus > have && !getenv("MLVB_SHARE_DEVICES")) ngpus = have;
    for (int i = 0; i < ngpus; i++) {
        if (mlvb_context_create(i % have, slots, &g_ctx[i]) != MLVB_OK) {
            g_nctx = i;
            g_owns_ctx = 1;
            drop_contexts();
            return 0;
        }
    }
    g_nctx = ngpus;
    g_owns_ctx = 1;
    g_chunk = chunk < 1 ? 1 : (chunk > MAX_BATCH ? MAX_BATCH : chunk);
    return g_nctx;
}

void frame_builder_shutdown(void) { drop_contexts(); }
int frame_builder_gpu_count(void) { return g_nctx; }
mlvb_context *frame_builder_context(int i) { return (i >= 0 && i < g_nctx) ? g_ctx[i] : NULL; }

static int contexts_ready(void)
{
    if (g_nctx) return 1;
    mlvb_context *ctx = mlvb_default_context();
    if (!ctx) {
        fprintf(stderr, "frame_builder: no CUDA context -- libmlvfs_b200 has no CPU path\n");
        return 0;
    }
    g_ctx[0] = ctx;
    g_nctx = 1;
    return 1;
}

static mlvb_context *context_for_frame(int frame) { return g_ctx[(frame / g_chunk) % g_nctx]; }

/* "/<sub/dirs/>clip.MLV/clip_000123.dng" -> real MLV path + frame number (main.c:800-872, 316-328;
 * plain naming scheme only) */
static int resolve(const char *dng_path, const char *mlv_dir, char *mlv_file, size_t cap, int *frame)
{
    const char *slash = strrchr(dng_path, '/');
    const char *dot = strrchr(dng_path, '.');
    if (!slash || !dot || dot - slash < 8 || strcmp(dot, ".dng")) return 0;
    *frame = atoi(dot - 6);
    size_t dir_len = (size_t)(slash - dng_path);
    while (*dng_path == '/') { dng_path++; dir_len--; }
    if ((size_t)snprintf(mlv_file, cap, "%s/%.*s", mlv_dir, (int)dir_len, dng_path) >= cap) return 0;
    size_t n = strlen(mlv_file);
    return n > 4 && (!strcmp(mlv_file + n - 4, ".MLV") || !strcmp(mlv_file + n - 4, ".mlv"));
}

int frame_builder_frame_limit(const char *dng_path)
{
    char mlv_file[4096], dir[4096];
    int frame;
    pthread_mutex_lock(&g_cfg_mu);
    snprintf(dir, sizeof(dir), "%s", g_mlv_dir);
    pthread_mutex_unlock(&g_cfg_mu);
    if (!resolve(dng_path, dir, mlv_file, sizeof(mlv_file), &frame)) return 0;
    return mlv_get_frame_count(mlv_file);
}

/* ---- priming: per-clip state comes from frame 0, on every GPU, in one clip order ----
 * The reference leaves "first frame" to thread timing; we pin it to frame 0 (SURVEY.md section 7 hard part 3).
 * The key holds the clip and the options the per-clip state depends on (the web GUI may change options between
 * frames: --bad-pix <-> --really-bad-pix is another map, --stripes may be switched on later). */
static pthread_mutex_t g_primed_mu = PTHREAD_MUTEX_INITIALIZER;
static char **g_primed = NULL;
static int g_nprimed = 0;

static void primed_key(char *key, size_t cap, const char *mlv_file, const mlvb_options *o)
{
    snprintf(key, cap, "%s|%d.%d.%d", mlv_file, o->fix_bad_pixels, o->fix_stripes, o->dual_iso);
}

void frame_builder_reset_clip_state(void)
{
    pthread_mutex_lock(&g_primed_mu);
    for (int i = 0; i < g_nprimed; i++) free(g_primed[i]);
    free(g_primed);
    g_primed = NULL;
    g_nprimed = 0;
    for (int i = 0; i < g_nctx; i++) mlvb_reset_clip_state(g_ctx[i]);
    pthread_mutex_unlock(&g_primed_mu);
}

static int build_one(mlvb_context *ctx, struct mlv_clip *clip, const char *mlv_file, int frame, const mlvb_options *opts,
                     uint16_t *data, struct frame_headers *hdrs, mlvb_frame_result *res);

/* makes sure every context has seen frame 0 of this clip under these options; returns 0 on failure */
static int prime_clip(struct mlv_clip *clip, const char *mlv_file, const mlvb_options *opts)
{
    char key[4300];
    primed_key(key, sizeof(key), mlv_file, opts);
    int ok = 1;
    const uint64_t t_prime = now_ns();
    pthread_mutex_lock(&g_primed_mu);                    /* held across the priming: clips are primed one at a time */
    int hit = 0;
    for (int i = 0; i < g_nprimed && !hit; i++) hit = !strcmp(g_primed[i], key);
    if (!hit) {
        struct frame_headers h0;
        if (mlv_clip_frame_headers(clip, 0, &h0)) {
            uint16_t *scratch = mlvb_host_alloc(dng_get_image_size(&h0));
            for (int g = 0; g < g_nctx && ok && scratch; g++) {
                mlvb_frame_result r;
                ok = build_one(g_ctx[g], clip, mlv_file, 0, opts, scratch, &h0, &r);
            }
            ok = ok && scratch;
            mlvb_host_free(scratch);
        } else ok = 0;
        if (ok) {
            char **grown = realloc(g_primed, sizeof(char *) * (size_t)(g_nprimed + 1));
            char *k = strdup(key);
            if (grown && k) { g_primed = grown; g_primed[g_nprimed++] = k; }
            else { free(k); if (grown) g_primed = grown; }
        }
    }
    pthread_mutex_unlock(&g_primed_mu);
    __sync_fetch_and_add(&g_ns_prime, now_ns() - t_prime);
    return ok;
}

/* read one frame's payload into pinned memory and run it through one context */
static int build_one(mlvb_context *ctx, struct mlv_clip *clip, const char *mlv_file, int frame, const mlvb_options *opts,
                     uint16_t *data, struct frame_headers *hdrs, mlvb_frame_result *res)
{
    (void)frame;
    const size_t payload_bytes = mlv_clip_payload_size(hdrs);
    uint8_t *payload = mlvb_host_alloc(payload_bytes + 16);
    memset(res, 0, sizeof(*res));
    const uint64_t t0 = now_ns();
    int ok = payload && mlv_clip_read_payload(clip, hdrs, payload, payload_bytes) == (ssize_t)payload_bytes;
    const uint64_t t1 = now_ns();
    if (ok) ok = mlvb_process_frame(ctx, hdrs, payload, payload_bytes, opts, mlv_file, data, res) == MLVB_OK;
    __sync_fetch_and_add(&g_ns_read, t1 - t0);
    __sync_fetch_and_add(&g_ns_gpu, now_ns() - t1);
    __sync_fetch_and_add(&g_n_frames, 1);
    __sync_fetch_and_add(&g_n_calls, 1);
    mlvb_host_free(payload);
    return ok;
}

/* the header is written after the pixel stages: dual ISO changes black/white (main.c:961-965), deflicker sets
 * exposure_bias (main.c:895-906) */
static uint8_t *finish_header(struct frame_headers *hdrs, const mlvb_frame_result *res, const char *dng_filename, double fps)
{
    const uint64_t t0 = now_ns();
    uint8_t *header = calloc(1, MLVB_DNG_HEADER_SIZE);
    if (!header) return NULL;
    hdrs->rawi_hdr.raw_info.black_level = res->black_level;
    hdrs->rawi_hdr.raw_info.white_level = res->white_level;
    hdrs->rawi_hdr.raw_info.exposure_bias[0] = res->exposure_bias[0];
    hdrs->rawi_hdr.raw_info.exposure_bias[1] = res->exposure_bias[1];
    if (g_header_writer) {
        char *base = strdup(dng_filename);                       /* main.c:935-940: the virtual path minus the file name */
        char *sep = base ? strrchr(base, '/') : NULL;
        if (sep) *sep = 0;
        g_header_writer(hdrs, header, 0, MLVB_DNG_HEADER_SIZE, fps, base);
        free(base);
    }
    __sync_fetch_and_add(&g_ns_header, now_ns() - t0);
    return header;
}

static void snapshot_config(struct frame_builder_config *cfg, char *dir, size_t cap)
{
    pthread_mutex_lock(&g_cfg_mu);                       /* snapshot: options may change between frames */
    *cfg = g_cfg;
    snprintf(dir, cap, "%s", g_mlv_dir);
    pthread_mutex_unlock(&g_cfg_mu);
}

int process_frame(struct image_buffer *image_buffer)
{
    struct frame_builder_config cfg;
    char dir[4096], mlv_file[4096];
    int frame;
    snapshot_config(&cfg, dir, sizeof(dir));
    if (!resolve(image_buffer->dng_filename, dir, mlv_file, sizeof(mlv_file), &frame)) return 1;
    struct mlv_clip *clip = mlv_clip_open(mlv_file);
    struct frame_headers hdrs;
    if (!clip || !mlv_clip_frame_headers(clip, frame, &hdrs)) return 1;
    if (!contexts_ready() || !prime_clip(clip, mlv_file, &cfg.options)) return 0;

    const size_t size = dng_get_image_size(&hdrs);
    uint16_t *data = mlvb_host_alloc(size);                      /* pinned, pooled; freed through resource_manager_set_data_free */
    mlvb_frame_result res;
    int ok = data && build_one(context_for_frame(frame), clip, mlv_file, frame, &cfg.options, data, &hdrs, &res);
    uint8_t *header = ok ? finish_header(&hdrs, &res, image_buffer->dng_filename, cfg.fps) : NULL;
    if (!ok || !header) {                                        /* read handler maps "no data" to a 0-byte read */
        mlvb_host_free(data);
        free(header);
        return 0;
    }
    image_buffer->size = size;
    image_buffer->data = data;
    image_buffer->header_size = MLVB_DNG_HEADER_SIZE;
    image_buffer->header = header;
    return 1;
}

/* A chunk of consecutive frames of one clip (the prefetch queue's look-ahead): one mlvb_process_frames call on the
 * chunk's GPU -- one H2D stream of payloads, one pass of the fused kernels over all frames, D2H into each buffer. */
int process_frame_batch(struct image_buffer **bufs, int n)
{
    if (n <= 0) return 1;
    if (n > MAX_BATCH) n = MAX_BATCH;
    struct frame_builder_config cfg;
    char dir[4096], mlv_file[4096], other[4096];
    int frame0;
    snapshot_config(&cfg, dir, sizeof(dir));
    if (!resolve(bufs[0]->dng_filename, dir, mlv_file, sizeof(mlv_file), &frame0)) return 1;
    struct mlv_clip *clip = mlv_clip_open(mlv_file);
    if (!clip || !contexts_ready() || !prime_clip(clip, mlv_file, &cfg.options)) return 0;

    struct frame_headers *hdrs = calloc((size_t)n, sizeof(*hdrs));
    mlvb_frame_result *res = calloc((size_t)n, sizeof(*res));
    const void **payloads = calloc((size_t)n, sizeof(void *));
    uint16_t **dsts = calloc((size_t)n, sizeof(uint16_t *));
    size_t *bytes = calloc((size_t)n, sizeof(size_t));
    int *slot = calloc((size_t)n, sizeof(int));                  /* batch position -> bufs index */
    int m = 0, ok = hdrs && res && payloads && dsts && bytes && slot;
    const uint64_t t_read = now_ns();
    for (int k = 0; k < n && ok; k++) {
        int frame;
        if (!resolve(bufs[k]->dng_filename, dir, other, sizeof(other), &frame) || strcmp(other, mlv_file)) continue;
        if (!mlv_clip_frame_headers(clip, frame, &hdrs[m])) continue;
        bytes[m] = mlv_clip_payload_size(&hdrs[m]);
        uint8_t *p = mlvb_host_alloc(bytes[m] + 16);
        uint16_t *d = mlvb_host_alloc(dng_get_image_size(&hdrs[m]));
        if (!p || !d || mlv_clip_read_payload(clip, &hdrs[m], p, bytes[m]) != (ssize_t)bytes[m]) {
            mlvb_host_free(p);
            mlvb_host_free(d);
            continue;
        }
        payloads[m] = p; dsts[m] = d; slot[m] = k;
        m++;
    }
    if (ok && m) {
        /* the whole chunk goes to the GPU that owns its first frame */
        const uint64_t t_gpu = now_ns();
        mlvb_process_frames(context_for_frame(frame0), m, hdrs, payloads, bytes, &cfg.options, mlv_file, dsts, res);
        __sync_fetch_and_add(&g_ns_read, t_gpu - t_read);
        __sync_fetch_and_add(&g_ns_gpu, now_ns() - t_gpu);
        __sync_fetch_and_add(&g_n_frames, (uint64_t)m);
        __sync_fetch_and_add(&g_n_calls, 1);
        for (int j = 0; j < m; j++) {
            struct image_buffer *ib = bufs[slot[j]];
            uint8_t *header = res[j].status == MLVB_OK ? finish_header(&hdrs[j], &res[j], ib->dng_filename, cfg.fps) : NULL;
            mlvb_host_free((void *)payloads[j]);
            if (!header) { mlvb_host_free(dsts[j]); continue; }
            ib->size = dng_get_image_size(&hdrs[j]);
            ib->data = dsts[j];
            ib->header_size = MLVB_DNG_HEADER_SIZE;
            ib->header = header;
        }
    }
    free(hdrs); free(res); free(payloads); free(dsts); free(bytes); free(slot);
    return ok;
}
