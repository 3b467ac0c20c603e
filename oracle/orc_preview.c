/* oracle/orc_preview.c -- TEST INFRASTRUCTURE.  The fast dual-ISO preview (hdr.c:40-227), the tiny
 * histogram it uses (histogram.c:33-84, uint16 bins that wrap) and deflicker (main.c:895-906). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct { uint16_t white; uint32_t count; uint16_t *data; } hist16;

static hist16 *hist_new(uint16_t white)
{
    hist16 *h = malloc(sizeof(*h));
    h->white = white; h->count = 0;
    h->data = calloc((size_t)white + 1, sizeof(uint16_t));
    return h;
}
static void hist_free(hist16 *h) { free(h->data); free(h); }
/* histogram.c:52-59: every (skip+1)-th sample, clamped at white, 16-bit counters */
static void hist_put(hist16 *h, const uint16_t *d, uint32_t size, uint16_t skip)
{
    for (uint32_t i = 0; i < size; i += (uint32_t)skip + 1) h->data[d[i] < h->white ? d[i] : h->white]++;
    h->count += size / ((uint32_t)skip + 1);
}
static uint16_t hist_med(const hist16 *h)
{
    uint32_t middle = h->count / 2, cur = 0;
    for (uint32_t i = 0; i <= h->white; i++) { cur += h->data[i]; if (cur > middle) return (uint16_t)i; }
    return 0;
}

/* main.c:895-906; bias[0] = log2((target-black)/(median-black)) * 10000, bias[1] = 10000 */
void orc_deflicker(const uint16_t *img, size_t bytes, int bpp, int black_level, int target, int bias[2])
{
    uint16_t black = (uint16_t)black_level, white = (uint16_t)((1 << bpp) + 1);
    hist16 *h = hist_new(white);
    hist_put(h, img + 1, (uint32_t)((bytes - 1) / 2), 1);
    uint16_t median = hist_med(h);
    double correction = log2((double)(target - black) / (median - black));
    bias[0] = (int)(correction * 10000);
    bias[1] = 10000;
    hist_free(h);
}

/* hdr.c:40-227.  Returns 1 and scales the frame to 16 bit (caller multiplies black/white by 4), or 0.
 * `focus`/`nfocus` is the clip's focus-pixel map (may be NULL), applied with the horizontal interpolator
 * after the row phase was detected (hdr.c:117). */
int orc_hdr_preview(uint16_t *img, int width, int height, int black_level, int white_level, size_t max_size,
                    const orc_pixel *focus, size_t nfocus, int crop_x, int crop_y)
{
    uint16_t black = (uint16_t)black_level, white = (uint16_t)white_level;
    int w = width, h = height;
    hist16 *hist[4];
    for (int i = 0; i < 4; i++) hist[i] = hist_new(white);
    for (int y = 4; y < h - 4; y += 5)
        hist_put(hist[y % 4], img + (size_t)y * w + (y + 1) % 2, (uint32_t)(w - (y + 1) % 2), 3);
    int m[4];
    for (int i = 0; i < 4; i++) m[i] = hist_med(hist[i]) - black;
    int start;
    hist16 *lo, *hi;
    if (m[2] > m[0] * 2 && m[2] > m[1] * 2 && m[3] > m[0] * 2 && m[3] > m[1] * 2) { start = 0; lo = hist[0]; hi = hist[2]; }
    else if (m[0] > m[1] * 2 && m[0] > m[2] * 2 && m[3] > m[1] * 2 && m[3] > m[2] * 2) { start = 1; lo = hist[1]; hi = hist[0]; }
    else if (m[0] > m[2] * 2 && m[0] > m[3] * 2 && m[1] > m[2] * 2 && m[1] > m[3] * 2) { start = 2; lo = hist[2]; hi = hist[0]; }
    else if (m[1] > m[0] * 2 && m[1] > m[3] * 2 && m[2] > m[0] * 2 && m[2] > m[3] * 2) { start = 3; lo = hist[0]; hi = hist[2]; }   /* sic */
    else { for (int i = 0; i < 4; i++) hist_free(hist[i]); return 0; }

    if (focus && nfocus) orc_focuspix_apply(img, w, h, black_level, focus, nfocus, crop_x, crop_y, 1);

    /* histogram matching: dark level as a function of bright level, one point every >100 samples */
    const int min_pix = 100;
    int cap = w * h / min_pix + 1, n = 0;
    int *dx = malloc(sizeof(int) * (size_t)cap), *dy = malloc(sizeof(int) * (size_t)cap);
    double *dw = malloc(sizeof(double) * (size_t)cap);
    int acc_lo = 0, acc_hi = 0, raw_lo = 0, prev = 0, total = (int)hist[0]->count;
    for (int raw_hi = 0; raw_hi < total; raw_hi++) {
        /* bins past `white` do not exist; the reference reads heap there and leaves the loop at once */
        if (raw_hi > white) break;
        acc_hi += hi->data[raw_hi];
        while (acc_lo < acc_hi && raw_lo <= white) { acc_lo += lo->data[raw_lo]; raw_lo++; }
        if (raw_lo >= white) break;
        if (acc_hi - prev > min_pix && acc_hi > total * 1 / 100 && acc_hi < total * 99.99 / 100) {
            dx[n] = raw_hi - black; dy[n] = raw_lo - black;
            dw[n] = (raw_hi - black + 100) > 0 ? (raw_hi - black + 100) : 0;
            n++;
            prev = acc_hi;
        }
    }
    double mx = 0, my = 0, mxy = 0, mx2 = 0, wsum = 0;
    for (int i = 0; i < n; i++) {
        mx += dx[i] * dw[i]; my += dy[i] * dw[i];
        mxy += (double)dx[i] * dy[i] * dw[i]; mx2 += (double)dx[i] * dx[i] * dw[i];
        wsum += dw[i];
    }
    mx /= wsum; my /= wsum; mxy /= wsum; mx2 /= wsum;
    double a = (mxy - mx * my) / (mx2 - mx * mx), b = my - a * mx;
    free(dx); free(dy); free(dw);
    for (int i = 0; i < 4; i++) hist_free(hist[i]);

    uint16_t shadow = (uint16_t)(black + 1 / (a * a) + b);
#define SCALE(v) ((white < ((v) - black) * a + black + b) ? (double)white : ((v) - black) * a + black + b)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * w + x;
            if (((y - start + 4) % 4) >= 2) {            /* bright row: rescale, patch clipped pixels from the dark rows */
                if (img[i] >= white)
                    img[i] = (uint16_t)(y > 2 ? (y < h - 2 ? (img[i - 2 * w] + img[i + 2 * w]) / 2 : img[i - 2 * w]) : img[i + 2 * w]);
                else
                    img[i] = (uint16_t)SCALE(img[i]);
            } else if (img[i] < shadow) {                /* dark row: deep shadows come from the bright rows */
                img[i] = (uint16_t)(y > 2 ? (y < h - 2 ? (img[i - 2 * w] + SCALE(img[i + 2 * w])) / 2 : img[i - 2 * w]) : SCALE(img[i + 2 * w]));
            }
        }
#undef SCALE
    size_t count = max_size / 2;
    for (size_t i = 0; i < count; i++) img[i] = (uint16_t)(img[i] << 2);
    return 1;
}
