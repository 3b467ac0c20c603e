/*
 * frame_server.c -- command-line driver of the host frame-request path, without FUSE:
 *   mlvb_frames <mlv_dir> <clip.MLV> [--cs3x3 --bad-pix --stripes ...] [--prefetch=N] [--readers=T]
 *               [--gpus=G] [--batch=B] [--slots=S] [--repeat=R] [--frames=K] [--dump=out.raw] [--dump-headers=out.bin] [--hash]
 * --gpus: one context per GPU, frames dealt in chunks of B (--batch, default 8 with --prefetch, else 1): frame n on
 * GPU (n / B) mod G; look-ahead chunks of the prefetch queue run as one device batch.  --repeat: request the clip R
 * times (cache emptied in between) and report the sustained rate of all passes after the first.
 * Requests every frame of the clip the way the FUSE read handler does (main.c:1460):
 * get_or_create_image_buffer(path, &process_frame, &was_created), touches the data, releases it.
 * Prints frames/s and an FNV-1a hash per run (and optionally dumps the frames) so tests can compare
 * against the oracle.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "frame_builder.h"
#include "mlv_index.h"

static char g_clip[1024];
static int g_nframes, g_next = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static uint64_t *g_hash;
static FILE *g_dump;
static int g_failed = 0;
static FILE *g_dump_headers;
static int g_hashing = 0;
static uint64_t g_copy_ns = 0, g_copy_frames = 0;

/* FNV-1a over 64-bit words (the byte-wise form costs ~5 ms per 1080p frame and would be the slowest stage) */
static uint64_t fnv1a_words(const void *p, size_t n)
{
    const uint64_t *w = p;
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n / 8; i++) { h ^= w[i]; h *= 1099511628211ull; }
    for (size_t i = n / 8 * 8; i < n; i++) { h ^= ((const uint8_t *)p)[i]; h *= 1099511628211ull; }
    return h;
}

static void *reader(void *arg)
{
    (void)arg;
    char base[1024];
    snprintf(base, sizeof(base), "%s", g_clip);
    char *dot = strrchr(base, '.');
    if (dot) *dot = 0;
    uint8_t *sink = NULL;                                /* what a FUSE read() does with a frame: copy it out */
    size_t sink_cap = 0;
    for (;;) {
        pthread_mutex_lock(&g_mu);
        int i = g_next < g_nframes ? g_next++ : -1;
        pthread_mutex_unlock(&g_mu);
        if (i < 0) { free(sink); return NULL; }
        char path[2200];
        snprintf(path, sizeof(path), "/%s/%s_%06d.dng", g_clip, base, i);
        int created = 0;
        struct image_buffer *ib = get_or_create_image_buffer(path, &process_frame, &created);
        if (!ib || !ib->data) { __sync_fetch_and_add(&g_failed, 1); continue; }
        if (ib->size > sink_cap) { free(sink); sink = malloc(ib->size); sink_cap = sink ? ib->size : 0; }
        if (sink) {                                          /* main.c:1463-1483 copies the requested range to the FUSE buffer */
            struct timespec c0, c1;
            clock_gettime(CLOCK_MONOTONIC, &c0);
            memcpy(sink, ib->data, ib->size);
            clock_gettime(CLOCK_MONOTONIC, &c1);
            __sync_fetch_and_add(&g_copy_ns, (uint64_t)((c1.tv_sec - c0.tv_sec) * 1000000000ll + (c1.tv_nsec - c0.tv_nsec)));
            __sync_fetch_and_add(&g_copy_frames, 1);
        }
        g_hash[i] = g_hashing ? fnv1a_words(sink ? sink : (const uint8_t *)ib->data, ib->size) : ib->data[ib->size / 4];
        if (g_dump) {
            pthread_mutex_lock(&g_mu);
            fseeko(g_dump, (off_t)i * (off_t)ib->size, SEEK_SET);
            fwrite(ib->data, 1, ib->size, g_dump);
            pthread_mutex_unlock(&g_mu);
        }
        if (g_dump_headers && ib->header) {
            pthread_mutex_lock(&g_mu);
            fseeko(g_dump_headers, (off_t)i * (off_t)ib->header_size, SEEK_SET);
            fwrite(ib->header, 1, ib->header_size, g_dump_headers);
            pthread_mutex_unlock(&g_mu);
        }
        release_image_buffer_by_path(path);
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <mlv_dir> <clip.MLV> [options]\n", argv[0]);
        return 2;
    }
    struct frame_builder_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.mlv_path = argv[1];
    snprintf(g_clip, sizeof(g_clip), "%s", argv[2]);
    int prefetch = 0, readers = 1, limit = -1, gpus = 1, batch = 0, slots = 0, repeat = 1, workers = 0;
    const char *dump = NULL, *dump_headers = NULL;
    for (int i = 3; i < argc; i++) {                       /* option names of main.c:1853-1882 */
        const char *a = argv[i];
        if (!strcmp(a, "--cs2x2")) cfg.options.chroma_smooth = 2;
        else if (!strcmp(a, "--cs3x3")) cfg.options.chroma_smooth = 3;
        else if (!strcmp(a, "--cs5x5")) cfg.options.chroma_smooth = 5;
        else if (!strcmp(a, "--bad-pix")) cfg.options.fix_bad_pixels = 1;
        else if (!strcmp(a, "--really-bad-pix")) cfg.options.fix_bad_pixels = 2;
        else if (!strcmp(a, "--fix-pattern-noise")) cfg.options.fix_pattern_noise = 1;
        else if (!strcmp(a, "--stripes")) cfg.options.fix_stripes = 1;
        else if (!strncmp(a, "--deflicker=", 12)) cfg.options.deflicker = atoi(a + 12);
        else if (!strcmp(a, "--dual-iso-preview")) cfg.options.dual_iso = 1;
        else if (!strcmp(a, "--dual-iso")) cfg.options.dual_iso = 2;
        else if (!strcmp(a, "--amaze-edge")) cfg.options.hdr_interpolation_method = 0;
        else if (!strcmp(a, "--mean23")) cfg.options.hdr_interpolation_method = 1;
        else if (!strcmp(a, "--no-alias-map")) cfg.options.hdr_no_alias_map = 1;
        else if (!strcmp(a, "--alias-map")) cfg.options.hdr_no_alias_map = 0;
        else if (!strncmp(a, "--prefetch=", 11)) prefetch = atoi(a + 11);
        else if (!strncmp(a, "--readers=", 10)) readers = atoi(a + 10);
        else if (!strncmp(a, "--frames=", 9)) limit = atoi(a + 9);
        else if (!strncmp(a, "--dump=", 7)) dump = a + 7;
        else if (!strncmp(a, "--dump-headers=", 15)) dump_headers = a + 15;
        else if (!strncmp(a, "--gpus=", 7)) gpus = atoi(a + 7);
        else if (!strncmp(a, "--batch=", 8)) batch = atoi(a + 8);
        else if (!strncmp(a, "--slots=", 8)) slots = atoi(a + 8);
        else if (!strncmp(a, "--repeat=", 9)) repeat = atoi(a + 9);
        else if (!strncmp(a, "--workers=", 10)) workers = atoi(a + 10);
        else if (!strcmp(a, "--no-fullres")) cfg.options.hdr_no_fullres = 1;
        else if (!strcmp(a, "--hash")) g_hashing = 1;
        else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    }
    if (batch <= 0) batch = prefetch > 0 ? 8 : 1;
    frame_builder_configure(&cfg);
    resource_manager_set_data_free(mlvb_host_free);
    resource_manager_set_batch_builder(batch > 1 ? process_frame_batch : NULL, batch);
    resource_manager_set_prefetch(prefetch, workers, frame_builder_frame_limit);

    char mlv_file[4096];
    snprintf(mlv_file, sizeof(mlv_file), "%s/%s", argv[1], g_clip);
    g_nframes = mlv_get_frame_count(mlv_file);
    if (limit >= 0 && limit < g_nframes) g_nframes = limit;
    if (g_nframes <= 0) { fprintf(stderr, "no frames in %s\n", mlv_file); return 1; }
    gpus = frame_builder_use_gpus(gpus, slots, batch);
    if (gpus <= 0) { fprintf(stderr, "no CUDA device: no CPU path\n"); return 3; }
    g_hash = calloc((size_t)g_nframes, sizeof(uint64_t));
    if (dump) g_dump = fopen(dump, "wb");
    if (dump_headers) g_dump_headers = fopen(dump_headers, "wb");

    if (readers < 1) readers = 1;
    if (readers > 64) readers = 64;
    if (repeat < 1) repeat = 1;
    pthread_t th[64];
    struct timespec t0, t1;
    double dt = 0, dt_sustained = 0;
    for (int pass = 0; pass < repeat; pass++) {
        g_next = 0;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < readers; i++) pthread_create(&th[i], NULL, reader, NULL);
        for (int i = 0; i < readers; i++) pthread_join(th[i], NULL);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double d = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        if (pass == 0) dt = d; else dt_sustained += d;
        if (pass == 0 && repeat > 1) { frame_builder_reset_stats(); g_copy_ns = g_copy_frames = 0; }   /* the breakdown covers the warm passes */
        if (pass + 1 < repeat) {                               /* every pass builds every frame again */
            resource_manager_set_prefetch(0, 0, NULL);
            free_all_image_buffers();
            resource_manager_set_prefetch(prefetch, workers, frame_builder_frame_limit);
        }
    }
    uint64_t all = 1469598103934665603ull, built = 0, hits = 0;
    for (int i = 0; i < g_nframes; i++) { all ^= g_hash[i]; all *= 1099511628211ull; }
    resource_manager_prefetch_stats(&built, &hits);
    struct frame_builder_stats fs;
    frame_builder_get_stats(&fs);
    const double per = fs.frames ? 1e-3 / (double)fs.frames : 0.0;      /* -> microseconds of builder-thread time per frame */
    uint64_t device_batches = 0;
    for (int i = 0; i < gpus; i++) device_batches += mlvb_path_count(frame_builder_context(i), 2);
    printf("{\"frames\": %d, \"failed\": %d, \"seconds\": %.6f, \"fps\": %.2f, \"sustained_fps\": %.2f, \"passes\": %d, "
           "\"readers\": %d, \"prefetch\": %d, \"gpus\": %d, \"batch\": %d, "
           "\"prefetch_built\": %llu, \"prefetch_hits\": %llu, \"prefetch_batches\": %llu, \"device_batches\": %llu, "
           "\"builder_us_per_frame\": {\"read\": %.1f, \"gpu_call\": %.1f, \"header\": %.1f, \"prime\": %.1f}, \"reader_copy_us_per_frame\": %.1f, "
           "\"hash\": \"%016llx\", \"frame0_hash\": \"%016llx\"}\n",
           g_nframes, g_failed, dt, g_nframes / dt, repeat > 1 ? g_nframes * (repeat - 1) / dt_sustained : g_nframes / dt, repeat,
           readers, prefetch, gpus, batch, (unsigned long long)built, (unsigned long long)hits,
           (unsigned long long)resource_manager_prefetch_batches(), (unsigned long long)device_batches,
           fs.read_ns * per, fs.gpu_ns * per, fs.header_ns * per, fs.prime_ns * per,
           g_copy_frames ? 1e-3 * (double)g_copy_ns / (double)g_copy_frames : 0.0,
           (unsigned long long)all, (unsigned long long)g_hash[0]);
    if (g_dump) fclose(g_dump);
    if (g_dump_headers) fclose(g_dump_headers);
    resource_manager_shutdown();
    free_all_image_buffers();
    frame_builder_shutdown();
    mlvb_host_pool_trim();
    mlv_clip_close_all();
    return g_failed ? 1 : 0;
}
