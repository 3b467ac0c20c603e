/* oracle/orc_patternnoise.c -- TEST INFRASTRUCTURE.  Row/column pattern-noise removal, restating
 * patternnoise.c:47-380.  Works on the Bayer frame reinterpreted as int16 (main.c:948). */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static int cmp_i16(const void *a, const void *b) { return (int)*(const int16_t *)a - (int)*(const int16_t *)b; }
static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* wirth.h:129-131: k-th smallest with k = n/2 (odd n) or n/2 - 1 (even n): the LOWER median */
static int lower_median_i16(const int16_t *v, int n)
{
    int16_t tmp[128];
    memcpy(tmp, v, (size_t)n * sizeof(int16_t));
    qsort(tmp, (size_t)n, sizeof(int16_t), cmp_i16);
    return tmp[(n & 1) ? n / 2 : n / 2 - 1];
}

static int lower_median_int(int *v, int n)
{
    qsort(v, (size_t)n, sizeof(int), cmp_int);
    return v[(n & 1) ? n / 2 : n / 2 - 1];
}

static int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* patternnoise.c:88-180: per pixel, the run of neighbours whose average green stays within `thr` of
   this pixel's (at most 25 to the left, 24 to the right); medians of G1, G2, R-G, B-G over that run */
static void edge_aware_blur(const int16_t *r, const int16_t *g1, const int16_t *g2, const int16_t *b,
                            int16_t *rs, int16_t *g1s, int16_t *g2s, int16_t *bs, int w, int h)
{
    const int reach = 50 / 2, thr = 500;
    size_t n = (size_t)w * h;
    int16_t *avg = malloc(n * 2), *drg = malloc(n * 2), *dbg = malloc(n * 2);
    for (size_t i = 0; i < n; i++) {
        avg[i] = (int16_t)(((int)g1[i] + (int)g2[i]) / 2);
        drg[i] = (int16_t)(r[i] - avg[i]);
        dbg[i] = (int16_t)(b[i] - avg[i]);
    }
    for (int y = 0; y < h; y++) {
        const int16_t *a = avg + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int p0 = a[x], xr = x + 1, xl = x - 1;
            int rmax = x + reach < w ? x + reach : w, lmin = x - reach > 0 ? x - reach : 0;
            while (xr < rmax && abs(a[xr] - p0) <= thr) xr++;
            while (xl >= lmin && abs(a[xl] - p0) <= thr) xl--;
            int num = xr - xl - 1;
            size_t o = (size_t)y * w + xl + 1, i = (size_t)y * w + x;
            int m1 = lower_median_i16(g1 + o, num), m2 = lower_median_i16(g2 + o, num);
            int mg = (m1 + m2) / 2;
            g1s[i] = (int16_t)m1;
            g2s[i] = (int16_t)m2;
            rs[i] = (int16_t)(lower_median_i16(drg + o, num) + mg);
            bs[i] = (int16_t)(lower_median_i16(dbg + o, num) + mg);
        }
    }
    free(avg); free(drg); free(dbg);
}

/* patternnoise.c:185-282 */
static void fix_column_offsets(int16_t *orig, const int16_t *den, int w, int h, int white)
{
    size_t n = (size_t)w * h;
    int *offs = malloc((size_t)w * sizeof(int)), *col = malloc((size_t)(w > h ? w : h) * sizeof(int));
    uint8_t *mask = malloc(n);
    for (size_t i = 0; i < n; i++) {
        int hg = (i >= 2 && i + 2 < n) ? (int16_t)(orig[i - 2] - orig[i + 2]) : 0;      /* flat index: wraps rows */
        mask[i] = (abs(hg) > 500) || (orig[i] >= white);
    }
    for (int x = 0; x < w; x++) {
        int cnt = 0;
        for (int y = 0; y < h; y++) {
            size_t i = (size_t)y * w + x;
            if (!mask[i]) col[cnt++] = (int16_t)(orig[i] - den[i]);
        }
        offs[x] = cnt < 10 ? 0 : -lower_median_int(col, cnt);
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * w + x;
            orig[i] = (int16_t)clampi((int)orig[i] + offs[x], -32767, 32767);
        }
    int mc = lower_median_int(offs, w);            /* sorts offs: done after they were applied (:266-268) */
    for (size_t i = 0; i < n; i++) orig[i] = (int16_t)clampi((int)orig[i] - mc, 0, 32760);
    free(offs); free(col); free(mask);
}

/* patternnoise.c:312-355 */
static void fix_columns_rggb(int16_t *raw, int w, int h, int white)
{
    int pw = w / 2, ph = h / 2;
    size_t n = (size_t)pw * ph;
    int16_t *p[4], *s[4];
    for (int c = 0; c < 4; c++) { p[c] = malloc(n * 2); s[c] = malloc(n * 2); }
    for (int c = 0; c < 4; c++)                       /* c: 0 = (0,0) R, 1 = (1,0) G1, 2 = (0,1) G2, 3 = (1,1) B */
        for (int y = c >> 1; y < h; y += 2)
            for (int x = c & 1; x < w; x += 2) p[c][(x / 2) + (size_t)(y / 2) * pw] = raw[x + (size_t)y * w];
    edge_aware_blur(p[0], p[1], p[2], p[3], s[0], s[1], s[2], s[3], pw, ph);
    for (int c = 0; c < 4; c++) fix_column_offsets(p[c], s[c], pw, ph, white);
    for (int c = 0; c < 4; c++)
        for (int y = c >> 1; y < h; y += 2)
            for (int x = c & 1; x < w; x += 2) raw[x + (size_t)y * w] = p[c][(x / 2) + (size_t)(y / 2) * pw];
    for (int c = 0; c < 4; c++) { free(p[c]); free(s[c]); }
}

/* patternnoise.c:357-380 with debug_flags == 0: columns, then the same on the transposed frame */
void orc_fix_pattern_noise(int16_t *raw, int w, int h, int white)
{
    fix_columns_rggb(raw, w, h, white);
    int16_t *t = malloc((size_t)w * h * 2);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) t[y + (size_t)x * h] = raw[x + (size_t)y * w];
    fix_columns_rggb(t, h, w, white);
    for (int y = 0; y < w; y++)
        for (int x = 0; x < h; x++) raw[y + (size_t)x * w] = t[x + (size_t)y * h];
    free(t);
}
