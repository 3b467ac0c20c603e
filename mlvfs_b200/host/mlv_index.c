/*
 * mlv_index.c -- per-clip block index + frame-header cache.  See mlv_index.h.
 * Replaces reference main.c:429-558 (mlv_get_frame_headers), index.c:216-341 (make_index) and the
 * per-frame chunk re-opening of resource_manager.c:285-317 / index.c:368-423.
 */
#define _GNU_SOURCE
#include "mlv_index.h"

#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#define MAX_CHUNKS 100

struct block_ref {
    uint64_t time;      /* sort key: block timestamp, 0 for MLVI (index.c:283-289) */
    uint64_t offset;
    uint16_t chunk;
    uint8_t  kind;      /* 1 VIDF, 2 AUDF, 0 other */
    char     type[4];
    uint32_t size;
};

struct mlv_clip {
    struct mlv_clip *next;
    char *path;
    int nchunks;
    int fd[MAX_CHUNKS];
    int nframes;
    struct frame_headers *frames;   /* one fully resolved header bundle per video frame */
    uint8_t *has_rawi;
    off_t size0;                    /* size and mtime of the first chunk when the index was built: a clip that was */
    struct timespec mtime0;         /* re-recorded or is still growing is indexed again (the reference re-reads per request) */
    int retired;                    /* replaced by a newer index; kept (callers may still hold it) until close_all */
};

static pthread_mutex_t g_clips_mu = PTHREAD_MUTEX_INITIALIZER;
static struct mlv_clip *g_clips = NULL;

static int open_chunks(const char *path, int fd[MAX_CHUNKS])
{
    int n = 0;
    fd[0] = open(path, O_RDONLY);
    if (fd[0] < 0) {
        fprintf(stderr, "mlv_index: open('%s'): %s\n", path, strerror(errno));
        return 0;
    }
    n = 1;
    size_t len = strlen(path);
    if (len < 2) return n;
    char *name = strdup(path);
    for (int seq = 0; seq < 99 && n < MAX_CHUNKS; seq++) {       /* x.M00, x.M01, ... (index.c:389-413) */
        snprintf(name + len - 2, 3, "%02d", seq);
        int f = open(name, O_RDONLY);
        if (f < 0) break;
        fd[n++] = f;
    }
    free(name);
    return n;
}

/* stable merge sort by time (the reference bubble-sorts, which is stable: index.c:77-97) */
static void sort_blocks(struct block_ref *a, struct block_ref *tmp, size_t n)
{
    if (n < 2) return;
    size_t h = n / 2;
    sort_blocks(a, tmp, h);
    sort_blocks(a + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].time < a[i].time) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(*a));
}

static void read_into(int fd, uint64_t off, void *dst, size_t want, uint32_t block_size)
{
    size_t n = want < block_size ? want : block_size;            /* MIN(sizeof(hdr), blockSize), main.c:478 */
    if (pread(fd, dst, n, (off_t)off) != (ssize_t)n)
        fprintf(stderr, "mlv_index: short header read at %llu\n", (unsigned long long)off);
}

static void free_clip(struct mlv_clip *c)
{
    if (!c) return;
    for (int i = 0; i < c->nchunks; i++) close(c->fd[i]);
    free(c->frames);
    free(c->has_rawi);
    free(c->path);
    free(c);
}

static struct mlv_clip *build_clip(const char *path)
{
    struct mlv_clip *c = calloc(1, sizeof(*c));
    if (!c) return NULL;
    struct block_ref *blk = NULL, *tmp = NULL;
    c->nchunks = open_chunks(path, c->fd);
    if (!c->nchunks) goto fail;
    c->path = strdup(path);
    if (!c->path) goto fail;
    struct stat st;
    if (fstat(c->fd[0], &st) == 0) { c->size0 = st.st_size; c->mtime0 = st.st_mtim; }

    size_t nblk = 0, cap = 0;
    uint64_t guid = 0;
    for (int ch = 0; ch < c->nchunks; ch++) {
        uint64_t pos = 0;
        for (;;) {
            mlv_hdr_t h;
            if (pread(c->fd[ch], &h, sizeof(h), (off_t)pos) != (ssize_t)sizeof(h)) break;
            if (h.blockSize < sizeof(mlv_hdr_t) || h.blockSize > 1024u * 1024u * 1024u) {
                fprintf(stderr, "mlv_index: invalid block size %u at 0x%llx\n", h.blockSize, (unsigned long long)pos);
                break;
            }
            uint64_t t = h.timestamp;
            if (!memcmp(h.blockType, "MLVI", 4)) {
                mlv_file_hdr_t fh;
                memset(&fh, 0, sizeof(fh));
                read_into(c->fd[ch], pos, &fh, sizeof(fh), h.blockSize);
                if (fh.fileNum == 0) guid = fh.fileGuid;
                else if (guid != fh.fileGuid) break;                 /* foreign chunk (index.c:271-279) */
                t = 0;
            }
            if (memcmp(h.blockType, "NULL", 4)) {
                if (nblk == cap) {
                    cap = cap ? cap * 2 : 1024;
                    struct block_ref *grown = realloc(blk, cap * sizeof(*blk));
                    if (!grown) goto fail;
                    blk = grown;
                }
                struct block_ref *b = &blk[nblk++];
                b->time = t; b->offset = pos; b->chunk = (uint16_t)ch; b->size = h.blockSize;
                memcpy(b->type, h.blockType, 4);
                b->kind = !memcmp(h.blockType, "VIDF", 4) ? 1 : (!memcmp(h.blockType, "AUDF", 4) ? 2 : 0);
                if (b->kind == 1) c->nframes++;
            }
            pos += h.blockSize;
        }
    }
    tmp = malloc((nblk ? nblk : 1) * sizeof(*tmp));
    if (!tmp) goto fail;
    sort_blocks(blk, tmp, nblk);
    free(tmp);
    tmp = NULL;

    c->frames = calloc(c->nframes ? c->nframes : 1, sizeof(struct frame_headers));
    c->has_rawi = calloc(c->nframes ? c->nframes : 1, 1);
    if (!c->frames || !c->has_rawi) goto fail;
    struct frame_headers cur;
    memset(&cur, 0, sizeof(cur));
    int rawi = 0, f = 0;
    for (size_t i = 0; i < nblk; i++) {                               /* main.c:457-542 */
        const struct block_ref *b = &blk[i];
        int fd = c->fd[b->chunk];
        if (b->kind == 1) {
            struct frame_headers *fh = &c->frames[f];
            *fh = cur;
            fh->fileNumber = b->chunk;
            fh->position = b->offset;
            read_into(fd, b->offset, &fh->vidf_hdr, sizeof(fh->vidf_hdr), b->size);
            c->has_rawi[f] = (uint8_t)rawi;
            f++;
        } else if (b->kind == 0) {
            if (!memcmp(b->type, "MLVI", 4)) read_into(fd, b->offset, &cur.file_hdr, sizeof(cur.file_hdr), b->size);
            else if (!memcmp(b->type, "RTCI", 4)) read_into(fd, b->offset, &cur.rtci_hdr, sizeof(cur.rtci_hdr), b->size);
            else if (!memcmp(b->type, "IDNT", 4)) read_into(fd, b->offset, &cur.idnt_hdr, sizeof(cur.idnt_hdr), b->size);
            else if (!memcmp(b->type, "RAWI", 4)) { read_into(fd, b->offset, &cur.rawi_hdr, sizeof(cur.rawi_hdr), b->size); rawi = 1; }
            else if (!memcmp(b->type, "EXPO", 4)) read_into(fd, b->offset, &cur.expo_hdr, sizeof(cur.expo_hdr), b->size);
            else if (!memcmp(b->type, "LENS", 4)) read_into(fd, b->offset, &cur.lens_hdr, sizeof(cur.lens_hdr), b->size);
            else if (!memcmp(b->type, "WBAL", 4)) read_into(fd, b->offset, &cur.wbal_hdr, sizeof(cur.wbal_hdr), b->size);
        }
    }
    free(blk);
    return c;
fail:
    fprintf(stderr, "mlv_index: cannot index '%s'\n", path);
    free(blk);
    free(tmp);
    free_clip(c);
    return NULL;
}

static int clip_is_stale(const struct mlv_clip *c)
{
    struct stat st;
    if (stat(c->path, &st) != 0) return 0;                           /* vanished: keep serving the open descriptors */
    return st.st_size != c->size0 || st.st_mtim.tv_sec != c->mtime0.tv_sec || st.st_mtim.tv_nsec != c->mtime0.tv_nsec;
}

struct mlv_clip *mlv_clip_open(const char *mlv_path)
{
    pthread_mutex_lock(&g_clips_mu);
    struct mlv_clip *c = g_clips;
    while (c && (c->retired || strcmp(c->path, mlv_path))) c = c->next;
    if (c && clip_is_stale(c)) {                                     /* re-recorded or still growing: index it again */
        c->retired = 1;
        c = NULL;
    }
    if (!c) {
        c = build_clip(mlv_path);
        if (c) { c->next = g_clips; g_clips = c; }
    }
    pthread_mutex_unlock(&g_clips_mu);
    return c;
}

int mlv_clip_frame_count(const struct mlv_clip *clip) { return clip ? clip->nframes : 0; }

int mlv_clip_frame_headers(const struct mlv_clip *clip, int index, struct frame_headers *out)
{
    if (!clip || index < 0 || index >= clip->nframes) {
        if (clip) fprintf(stderr, "%s: vidf block for frame %d was not found\n", clip->path, index);
        return 0;
    }
    *out = clip->frames[index];
    if (!clip->has_rawi[index]) {
        fprintf(stderr, "%s: no rawi block was found\n", clip->path);
        return 0;
    }
    return 1;
}

size_t mlv_clip_payload_size(const struct frame_headers *hdr)
{
    size_t head = sizeof(mlv_vidf_hdr_t) + hdr->vidf_hdr.frameSpace;          /* main.c:583 */
    return hdr->vidf_hdr.blockSize > head ? hdr->vidf_hdr.blockSize - head : 0;
}

ssize_t mlv_clip_read_payload(const struct mlv_clip *clip, const struct frame_headers *hdr, void *dst, size_t cap)
{
    if (!clip || hdr->fileNumber >= (uint32_t)clip->nchunks) return -1;
    size_t n = mlv_clip_payload_size(hdr);
    if (n > cap) n = cap;
    off_t off = (off_t)(hdr->position + sizeof(mlv_vidf_hdr_t) + hdr->vidf_hdr.frameSpace);
    size_t done = 0;
    while (done < n) {
        ssize_t r = pread(clip->fd[hdr->fileNumber], (uint8_t *)dst + done, n - done, off + (off_t)done);
        if (r <= 0) {
            if (r < 0 && errno == EINTR) continue;
            break;
        }
        done += (size_t)r;
    }
    return (ssize_t)done;
}

void mlv_clip_close_all(void)
{
    pthread_mutex_lock(&g_clips_mu);
    while (g_clips) {
        struct mlv_clip *n = g_clips->next;
        free_clip(g_clips);
        g_clips = n;
    }
    pthread_mutex_unlock(&g_clips_mu);
}

int mlv_get_frame_headers(const char *mlv_filename, int index, struct frame_headers *frame_headers)
{
    return mlv_clip_frame_headers(mlv_clip_open(mlv_filename), index, frame_headers);
}

int mlv_get_frame_count(const char *real_path)
{
    return mlv_clip_frame_count(mlv_clip_open(real_path));
}
