/* oracle/orc_stripes.c -- TEST INFRASTRUCTURE. Vertical-stripe correction, restating
 * stripes.c:102-266, plus the glibc rand() the reference dithers with (stripes.c:129-130). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/*
 * glibc's default rand(): the TYPE_3 additive-feedback generator of random_r.c (degree 31,
 * separation 3).  Published algorithm: r[0] = seed; r[i] = 16807*r[i-1] mod (2^31-1) computed with
 * Schrage's method for i = 1..30; r[i] = r[i-31] for i = 31..33; then o[k] = r[k-31] + r[k-3]
 * (mod 2^32) with the first 310 outputs discarded, and rand() returns o >> 1.
 * An unseeded process starts from seed 1.  tests/test_oracle.py checks this against libc rand().
 */
void orc_rand_seed(orc_rand_t *st, unsigned seed)
{
    int32_t word = seed ? (int32_t)seed : 1;
    st->r[0] = (uint32_t)word;
    for (int i = 1; i < 31; i++) {
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        st->r[i] = (uint32_t)word;
    }
    st->f = 3; st->b = 0;
    for (int i = 0; i < 310; i++) {
        st->r[st->f] += st->r[st->b];
        st->f = (st->f + 1) % 31; st->b = (st->b + 1) % 31;
    }
    st->primed = 1;
}

int orc_rand_next(orc_rand_t *st)
{
    if (!st->primed) orc_rand_seed(st, 1);
    st->r[st->f] += st->r[st->b];
    int out = (int)(st->r[st->f] >> 1);
    st->f = (st->f + 1) % 31; st->b = (st->b + 1) % 31;
    return out;
}

#define HBINS 65536            /* FIXP_RANGE, stripes.c:103 */

/* stripes.c:108-140 */
static void add_pair(int *hist, int num[8], int g, int a, int b, int white, orc_rand_t *rng)
{
    int lo = a < b ? a : b, hi = a > b ? a : b;
    if (lo < 32) return;
    if (hi > white / 1.5) return;
    double af = a + (orc_rand_next(rng) % 1024) / 1024.0 - 0.5;
    double bf = b + (orc_rand_next(rng) % 1024) / 1024.0 - 0.5;
    double ev = log2(af / bf);
    int bin = (int)(HBINS / 2 + ev * HBINS / 2);       /* F2H, stripes.c:105 */
    bin = bin < 0 ? 0 : (bin > HBINS - 1 ? HBINS - 1 : bin);
    hist[g * HBINS + bin] += 1;
    num[g] += 1;
}

/* stripes.c:143-248.  For every 8-pixel block the columns 2..7 are compared against the reference
   columns 0/1 of this block (a, b) and of the next block (a2, b2), three-to-one weighted by distance. */
int orc_stripes_compute(const uint16_t *img, int w, int h, int black, int white, int frame_size,
                        orc_rand_t *rng, int coef[8])
{
    int *hist = calloc((size_t)8 * HBINS, sizeof(int));
    int num[8] = {0};
    static const int near_weight[8] = {0, 0, 3, 3, 2, 2, 1, 1};   /* repeats of (a|b, p) before (a2|b2, p) */
    for (int y = 0; y < h; y++)
        for (int x = y * w; x < y * w + w - 10; x += 8) {
            int p[10];
            for (int k = 0; k < 10; k++) p[k] = img[x + k] - black;
            for (int g = 2; g < 8; g++) {
                int near = p[g & 1], far = p[8 + (g & 1)];
                for (int r = 0; r < 4; r++)
                    add_pair(hist, num, g, r < near_weight[g] ? near : far, p[g], white, rng);
            }
        }
    for (int g = 0; g < 8; g++) {
        if (num[g] < frame_size / 128) continue;                    /* stripes.c:221 */
        int t = 0;
        for (int k = 0; k < HBINS; k++) {
            t += hist[g * HBINS + k];
            if (t >= num[g] / 2) {
                double ev = (double)(k - HBINS / 2) / (HBINS / 2);  /* H2F */
                coef[g] = (int)(pow(2, ev) * 65536);
                break;
            }
        }
    }
    coef[0] = coef[1] = 65536;
    int needed = 0;
    for (int g = 0; g < 8; g++) {
        double c = (double)coef[g] / 65536;
        if (c < 0.998 || c > 1.002) needed = 1;
    }
    free(hist);
    return needed;
}

/* stripes.c:250-266 (offset 0): double multiply, truncating cast, clamp at white */
void orc_stripes_apply(uint16_t *img, size_t n, int w, int black, int white, int needed, const int coef[8])
{
    if (!needed || w % 8 != 0) return;
    uint16_t blk = (uint16_t)black, wht = (uint16_t)white;
    for (size_t i = 0; i < n; i++) {
        double c = coef[i % 8];
        if (c && img[i] > blk + 64) {
            double v = (img[i] - blk) * c / 65536 + blk;
            img[i] = (uint16_t)(wht < v ? wht : v);
        }
    }
}
