/* oracle/orc_pixfix.c -- TEST INFRASTRUCTURE. Bad- and focus-pixel repair, restating cs.c:87-331
 * and cs.c:440-503.  The lists are applied strictly in order on the live image: an entry may read
 * pixels an earlier entry already rewrote. */
#include <stdlib.h>
#include "oracle.h"

static inline int wsub(int a, int b) { return (int)((uint32_t)a - (uint32_t)b); }
static inline int wabs(int a) { return a > 0 ? a : (int)(0u - (uint32_t)a); }        /* ABS, mlvfs.h:85 */
static inline int wmul(int a, int b) { return (int)((uint32_t)a * (uint32_t)b); }
static inline int clamp_ev(int e)
{
    return e < 0 ? 0 : (e > 14 * ORC_EV_RES - 1 ? 14 * ORC_EV_RES - 1 : e);
}

/* |EV(a) - EV(b)| of the two pixels at i+o1, i+o2 */
static inline int grad(const uint16_t *im, int i, int o1, int o2, const int *raw2ev)
{
    return wabs(wsub(raw2ev[im[i + o1]], raw2ev[im[i + o2]]));
}

/* cs.c:87-109 (step = 1) and cs.c:111-133 (step = w): two-neighbour gradient-weighted EV mean */
static void interp_line(uint16_t *im, int i, int step, const int *raw2ev, const int *ev2raw, int black)
{
    int d1 = grad(im, i, 3 * step, step, raw2ev);
    int d2 = grad(im, i, -step, -3 * step, raw2ev);
    int sum = (int)((uint32_t)d1 + (uint32_t)d2);
    if (sum == 0) { im[i] = im[i + 2 * step]; return; }
    int c1 = ((sum - d1) << 8) / sum;
    int c2 = ((sum - d2) << 8) / sum;
    int ev = (wmul(raw2ev[im[i + 2 * step]], c1) >> 8) + (wmul(raw2ev[im[i - 2 * step]], c2) >> 8);
    im[i] = (uint16_t)(ev2raw[clamp_ev(ev)] + black);
}

/* cs.c:135-168: four-neighbour version; note the zero-gradient fallback copies the RIGHT neighbour */
static void interp_cross(uint16_t *im, int i, int w, const int *raw2ev, const int *ev2raw, int black)
{
    int dv1 = grad(im, i, 3 * w, w, raw2ev);
    int dv2 = grad(im, i, -w, -3 * w, raw2ev);
    int dh1 = grad(im, i, 3, 1, raw2ev);
    int dh2 = grad(im, i, -1, -3, raw2ev);
    int sum = (int)((uint32_t)dh1 + (uint32_t)dh2 + (uint32_t)dv1 + (uint32_t)dv2);
    if (sum == 0) { im[i] = im[i + 2]; return; }
    int den = wmul(3, sum);
    int cv1 = ((sum - dv1) << 8) / den;
    int cv2 = ((sum - dv2) << 8) / den;
    int ch1 = ((sum - dh1) << 8) / den;
    int ch2 = ((sum - dh2) << 8) / den;
    int ev = (wmul(raw2ev[im[i + 2 * w]], cv1) >> 8) + (wmul(raw2ev[im[i - 2 * w]], cv2) >> 8)
           + (wmul(raw2ev[im[i + 2]], ch1) >> 8) + (wmul(raw2ev[im[i - 2]], ch2) >> 8);
    im[i] = (uint16_t)(ev2raw[clamp_ev(ev)] + black);
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* cs.c:257-306.  The reference tracks the two largest of the eight same-colour neighbours at +-2
   (as negated values) and, in aggressive mode, selects the third largest with a Wirth select. */
size_t orc_badpix_detect(const uint16_t *img, int w, int h, int black, int aggressive,
                         int crop_x, int crop_y, orc_pixel *list, size_t cap)
{
    const int *raw2ev = orc_raw2ev(black);
    if (!raw2ev) return 0;
    const int dark_noise = 12;                         /* cs.c:257 */
    const int dark_min = black - dark_noise * 8, dark_max = black + dark_noise * 8;
    size_t n = 0;
    for (int y = 6; y < h - 6; y++)
        for (int x = 6; x < w - 6; x++) {
            int p = img[x + y * w];
            int nb[8], k = 0;
            for (int i = -2; i <= 2; i += 2)
                for (int j = -2; j <= 2; j += 2)
                    if (i || j) nb[k++] = img[(x + j) + (y + i) * w];
            qsort(nb, 8, sizeof(int), cmp_int);
            int second = nb[6], third = nb[5];
            int bad = 0;
            if (p < dark_min) bad = 1;                                                  /* cold, :289 */
            else if (wsub(raw2ev[p], raw2ev[second]) > 2 * ORC_EV_RES && p > dark_max) bad = 1; /* hot, :293 */
            else if (aggressive &&
                     (wsub(raw2ev[p], raw2ev[second]) > ORC_EV_RES || wsub(raw2ev[p], raw2ev[third]) > ORC_EV_RES) &&
                     p > dark_max) bad = 1;                                             /* :297-304 */
            if (bad) {
                if (n < cap) { list[n].x = x + crop_x; list[n].y = y + crop_y; }
                n++;
            }
        }
    return n;
}

/* cs.c:314-330 */
void orc_badpix_apply(uint16_t *img, int w, int h, int black, const orc_pixel *list, size_t n,
                      int crop_x, int crop_y, int dual_iso)
{
    const int *raw2ev = orc_raw2ev(black);
    const int *ev2raw = orc_ev2raw();
    if (!raw2ev) return;
    for (size_t m = 0; m < n; m++) {
        int x = list[m].x - crop_x, y = list[m].y - crop_y;
        if (x > 2 && x < w - 3 && y > 2 && y < h - 3) {
            if (dual_iso) interp_line(img, x + y * w, 1, raw2ev, ev2raw, black);
            else interp_cross(img, x + y * w, w, raw2ev, ev2raw, black);
        }
    }
}

/* cs.c:462-501: interior entries like bad pixels; border entries fall back to 1-D or copies */
void orc_focuspix_apply(uint16_t *img, int w, int h, int black, const orc_pixel *map, size_t n,
                        int crop_x, int crop_y, int dual_iso)
{
    const int *raw2ev = orc_raw2ev(black);
    const int *ev2raw = orc_ev2raw();
    if (!raw2ev) return;
    for (size_t m = 0; m < n; m++) {
        int x = map[m].x - crop_x, y = map[m].y - crop_y;
        int i = x + y * w;
        if (x > 2 && x < w - 3 && y > 2 && y < h - 3) {
            if (dual_iso) interp_line(img, i, 1, raw2ev, ev2raw, black);
            else interp_cross(img, i, w, raw2ev, ev2raw, black);
        } else if (i > 0 && i < w * h) {
            int hedge = (x >= w - 3 && x < w) || (x >= 0 && x <= 3);
            int vedge = (y >= h - 3 && y < h) || (y >= 0 && y <= 3);
            if (hedge && !vedge && !dual_iso) interp_line(img, i, w, raw2ev, ev2raw, black);
            else if (vedge && !hedge) interp_line(img, i, 1, raw2ev, ev2raw, black);
            else if (x >= 0 && x <= 3) img[i] = img[i + 2];
            else if (x >= w - 3 && x < w) img[i] = img[i - 2];
        }
    }
}
