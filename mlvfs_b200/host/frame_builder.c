/* frame_builder.c -- our process_frame.  See frame_builder.h; stage order and header handling follow
 * reference main.c:908-1005, the pixel stages themselves run inside libmlvfs_b200.so. */
#define _GNU_SOURCE
#include "frame_builder.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mlv_index.h"

static pthread_mutex_t g_cfg_mu = PTHREAD_MUTEX_INITIALIZER;
static struct frame_builder_config g_cfg;
static char g_mlv_dir[4096];
static mlvb_context *g_ctx = NULL;
static dng_header_writer g_header_writer = NULL;

void frame_builder_configure(const struct frame_builder_config *cfg)
{
    pthread_mutex_lock(&g_cfg_mu);
    g_cfg = *cfg;
    snprintf(g_mlv_dir, sizeof(g_mlv_dir), "%s", cfg->mlv_path ? cfg->mlv_path : ".");
    g_cfg.mlv_path = g_mlv_dir;
    pthread_mutex_unlock(&g_cfg_mu);
}

void frame_builder_set_context(mlvb_context *ctx) { g_ctx = ctx; }
void frame_builder_set_header_writer(dng_header_writer fn) { g_header_writer = fn; }

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

/* Per-clip state (stripe coefficients, bad-pixel map, dual-ISO LUT white) is created by the first
 * frame the library sees for a clip.  The reference leaves "first" to thread timing; we pin it to
 * frame 0 (SURVEY.md section 7 hard part 3): a clip's first request for any other frame builds frame 0
 * first, so sequential and prefetching / multi-reader runs give identical output. */
static pthread_mutex_t g_primed_mu = PTHREAD_MUTEX_INITIALIZER;
static char **g_primed = NULL;
static int g_nprimed = 0;

static int clip_is_primed(const char *mlv_file)
{
    int hit = 0;
    pthread_mutex_lock(&g_primed_mu);
    for (int i = 0; i < g_nprimed && !hit; i++) hit = !strcmp(g_primed[i], mlv_file);
    pthread_mutex_unlock(&g_primed_mu);
    return hit;
}

static void clip_mark_primed(const char *mlv_file)
{
    pthread_mutex_lock(&g_primed_mu);
    int hit = 0;
    for (int i = 0; i < g_nprimed && !hit; i++) hit = !strcmp(g_primed[i], mlv_file);
    if (!hit) {
        g_primed = realloc(g_primed, sizeof(char *) * (size_t)(g_nprimed + 1));
        g_primed[g_nprimed++] = strdup(mlv_file);
    }
    pthread_mutex_unlock(&g_primed_mu);
}

int process_frame(struct image_buffer *image_buffer)
{
    struct frame_builder_config cfg;
    char dir[4096], mlv_file[4096];
    int frame;
    pthread_mutex_lock(&g_cfg_mu);                       /* snapshot: options may change between frames */
    cfg = g_cfg;
    snprintf(dir, sizeof(dir), "%s", g_mlv_dir);
    pthread_mutex_unlock(&g_cfg_mu);

    if (!resolve(image_buffer->dng_filename, dir, mlv_file, sizeof(mlv_file), &frame)) return 1;
    struct mlv_clip *clip = mlv_clip_open(mlv_file);
    struct frame_headers hdrs;
    if (!clip || !mlv_clip_frame_headers(clip, frame, &hdrs)) return 1;

    if (frame != 0 && !clip_is_primed(mlv_file)) {
        struct image_buffer first;
        memset(&first, 0, sizeof(first));
        size_t n = strlen(image_buffer->dng_filename);
        first.dng_filename = strdup(image_buffer->dng_filename);
        memcpy(first.dng_filename + n - 10, "000000", 6);         /* "..._NNNNNN.dng" */
        process_frame(&first);
        if (first.data) mlvb_host_free(first.data);
        free(first.header);
        free(first.dng_filename);
    }

    mlvb_context *ctx = g_ctx ? g_ctx : mlvb_default_context();
    if (!ctx) {
        fprintf(stderr, "frame_builder: no CUDA context -- libmlvfs_b200 has no CPU path\n");
        return 0;
    }
    const size_t size = dng_get_image_size(&hdrs);
    const size_t payload_bytes = mlv_clip_payload_size(&hdrs);
    uint16_t *data = mlvb_host_alloc(size);                      /* pinned; freed through resource_manager_set_data_free */
    uint8_t *payload = mlvb_host_alloc(payload_bytes + 16);
    uint8_t *header = calloc(1, MLVB_DNG_HEADER_SIZE);
    int ok = data && payload && header &&
             mlv_clip_read_payload(clip, &hdrs, payload, payload_bytes) == (ssize_t)payload_bytes;
    mlvb_frame_result res;
    memset(&res, 0, sizeof(res));
    if (ok) ok = mlvb_process_frame(ctx, &hdrs, payload, payload_bytes, &cfg.options, mlv_file, data, &res) == MLVB_OK;
    mlvb_host_free(payload);
    if (!ok) {                                                   /* read handler maps "no data" to a 0-byte read */
        mlvb_host_free(data);
        free(header);
        return 0;
    }
    /* the header is written after the pixel stages: dual ISO changes black/white (main.c:961-965),
       deflicker sets exposure_bias (main.c:895-906) */
    hdrs.rawi_hdr.raw_info.black_level = res.black_level;
    hdrs.rawi_hdr.raw_info.white_level = res.white_level;
    hdrs.rawi_hdr.raw_info.exposure_bias[0] = res.exposure_bias[0];
    hdrs.rawi_hdr.raw_info.exposure_bias[1] = res.exposure_bias[1];
    if (g_header_writer) {
        char *base = strdup(image_buffer->dng_filename);
        char *sep = base ? strrchr(base, '/') : NULL;
        if (sep) *sep = 0;
        g_header_writer(&hdrs, header, 0, MLVB_DNG_HEADER_SIZE, cfg.fps, base);
        free(base);
    }
    clip_mark_primed(mlv_file);
    image_buffer->size = size;
    image_buffer->data = data;
    image_buffer->header_size = MLVB_DNG_HEADER_SIZE;
    image_buffer->header = header;
    return 1;
}
