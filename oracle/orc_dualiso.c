/* oracle/orc_dualiso.c -- TEST INFRASTRUCTURE.  Full dual-ISO conversion ("cr2hdr 20-bit"), restating
 * hdr.c:250-1957: both interpolation paths, mean23 (hdr.c:1231-1304) and AMaZE + edge-directed
 * (hdr.c:917-1229 on top of orc_amaze.c).
 *
 * The reference keeps its 20-bit EV tables and the full-res curve in function-static storage that is
 * rebuilt only when the black level changes (hdr.c:1080-1093, 1240-1253, 1575-1588, 1672-1685,
 * 890-898); orc_diso_state carries exactly that state so a test can model "a fresh process".
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define EV ORC_EV_RES
#define N20 (1 << 20)
#define ALIAS_MAP_MAX 15000                      /* hdr.c:248 */
static const double fullres_thr = 0.8;           /* hdr.c:245 */

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static int iclamp(int x, int lo, int hi) { return imax(imin(x, hi), lo); }
static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }
static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}
/* k-th smallest, k zero based (wirth.h:37-62 returns a[k] of the partially sorted array) */
static int kth_smallest(int *v, int n, int k)
{
    if (n <= 0 || k < 0) return 0;
    qsort(v, (size_t)n, sizeof(int), cmp_int);
    return v[k];
}
static int lower_median(int *v, int n) { return kth_smallest(v, n, (n & 1) ? n / 2 : n / 2 - 1); }

void orc_diso_state_init(orc_diso_state *S) { memset(S, 0, sizeof(*S)); S->lut_black = -1; S->curve_black = -1; }
void orc_diso_state_free(orc_diso_state *S)
{
    free(S->raw2ev); free(S->ev2raw_0); free(S->fullres_curve);
    orc_diso_state_init(S);
}

/* hdr.c:839-874 */
static void build_luts(orc_diso_state *S, int black, int white)
{
    if (!S->raw2ev) S->raw2ev = malloc(sizeof(int) * N20);
    if (!S->ev2raw_0) S->ev2raw_0 = malloc(sizeof(int) * 24 * EV);
    int *raw2ev = S->raw2ev, *ev2raw = S->ev2raw_0 + 10 * EV;
    for (int i = 0; i < N20; i++) {
        double signal = dmax(i / 64.0 - black / 64.0, -1023);
        raw2ev[i] = signal > 0 ? (int)round(log2(1 + signal) * EV) : -(int)round(log2(1 - signal) * EV);
    }
    for (int i = -10 * EV; i < 0; i++)
        ev2raw[i] = (int)dmax(dmin(black + 64 - round(64 * pow(2, ((double)-i / EV))), black), 0);
    for (int i = 0; i < 14 * EV; i++) {
        ev2raw[i] = (int)dmax(dmin(black - 64 + round(64 * pow(2, ((double)i / EV))), N20 - 1), black);
        if (i >= raw2ev[white]) ev2raw[i] = imax(ev2raw[i], white);
    }
    ev2raw[raw2ev[0]] = 0;
    S->lut_black = black;
    S->lut_white = white;
}

/* hdr.c:890-913 */
static const double *fullres_curve_for(orc_diso_state *S, int black)
{
    if (S->curve_black == black && S->fullres_curve) return S->fullres_curve;
    if (!S->fullres_curve) S->fullres_curve = malloc(sizeof(double) * N20);
    for (int i = 0; i < N20; i++) {
        double ev2 = log2(dmax(i / 64.0 - black / 64.0, 1));
        double c2 = -cos(dmax(dmin(ev2 - 4, 4), 0) * M_PI / 4);
        S->fullres_curve[i] = (c2 + 1) / 2;
    }
    S->curve_black = black;
    return S->fullres_curve;
}

/* hdr.c:407-439 */
int orc_hdr_check(const uint16_t *img, int w, int h, int black, int white)
{
    const double *raw2ev = orc_raw2evf(black);
    double avg = 0;
    int num = 0;
    for (int y = 2; y < h - 2; y++)
        for (int x = 2; x < w - 2; x++) {
            int p = img[x + y * w], p2 = img[x + (y + 2) * w];
            if ((p > black + 32 || p2 > black + 32) && p < white && p2 < white) {
                avg += fabs(raw2ev[p2] - raw2ev[p]);
                num++;
            }
        }
    avg /= num;
    return avg > 0.5;
}

/* hdr.c:441-495 */
static int identify_rggb(const uint16_t *img, int w, int h)
{
    int *hist = calloc(4 * 16384, sizeof(int));
    for (int y = 0; y < h / 4 * 4; y++)
        for (int x = 0; x < w; x++) hist[((y % 2) * 2 + (x % 2)) * 16384 + (img[x + y * w] & 16383)]++;
    double d_rggb = 0, d_gbrg = 0;
    int acc[4] = {0};
    for (int i = 0; i < 16384; i++) {
        for (int k = 0; k < 4; k++) acc[k] += hist[k * 16384 + i];
        d_rggb += abs(acc[1] - acc[2]);
        d_gbrg += abs(acc[0] - acc[3]);
    }
    free(hist);
    return d_rggb < d_gbrg;
}

/* hdr.c:497-636; y1 = active_area.y1 (1 after the GBRG row skip) */
static int identify_fields(const uint16_t *img, int w, int h, int black, int y1, int is_bright[4])
{
    const int white = 10000;
    int *hist = calloc(4 * 16384, sizeof(int));
    for (int y = (y1 + 3) & ~3; y < h / 4 * 4; y++)
        for (int x = 0; x < w; x++)
            if ((x % 2) != (y % 2)) hist[(y % 4) * 16384 + (img[x + y * w] & 16383)]++;
    int total = 0;
    for (int i = 0; i < 16384; i++) total += hist[i];
    int acc[4] = {0}, raw[4] = {0}, off[4] = {0};
    int ref_max = (int)(total * 0.998), ref_off = (int)(total * 0.05);
    for (int ref = 0; ref < ref_max; ref++) {
        for (int i = 0; i < 4; i++)
            while (acc[i] < ref) { acc[i] += hist[i * 16384 + raw[i]]; raw[i]++; }
        if (ref < ref_off && imax(imax(raw[0], raw[1]), imax(raw[2], raw[3])) < black + (white - black) / 4)
            memcpy(off, raw, sizeof(off));
        if (raw[0] >= white || raw[1] >= white || raw[2] >= white || raw[3] >= white) break;
    }
    free(hist);
    for (int i = 0; i < 4; i++) raw[i] -= off[i];
    int s[4];
    memcpy(s, raw, sizeof(s));
    qsort(s, 4, sizeof(int), cmp_int);
    double median_bright = (s[1] + s[2]) / 2;                   /* integer division, hdr.c:617 */
    for (int i = 0; i < 4; i++) is_bright[i] = raw[i] > median_bright;
    if (is_bright[0] + is_bright[1] + is_bright[2] + is_bright[3] != 2) return 0;
    if (is_bright[0] == is_bright[2] || is_bright[1] == is_bright[3]) return 0;
    return 1;
}

/* hdr.c:250-300 */
static void white_detect(const uint16_t *img, int w, int h, int y1, const int is_bright[4], int *white_dark, int *white_bright)
{
    int max_pix = w * h / 2 / 9;
    int *pix[2] = {malloc(sizeof(int) * (size_t)imax(max_pix, 1)), malloc(sizeof(int) * (size_t)imax(max_pix, 1))};
    int counts[2] = {0, 0};
    for (int y = y1; y < h; y += 3)
        for (int x = 0; x < w; x += 3) {
            int c = is_bright[y % 4];
            counts[c] = imin(counts[c], max_pix - 1);            /* full array: keep overwriting the last slot */
            pix[c][counts[c]] = -(int)img[x + y * w];
            counts[c]++;
        }
    int wd = -kth_smallest(pix[0], counts[0], 10) - 100;
    int wb = -kth_smallest(pix[1], counts[1], 50) - 1500;
    *white_dark = iclamp(wd, 10000, 16383);
    *white_bright = iclamp(wb, 5000, 16383);
    free(pix[0]); free(pix[1]);
}

/* hdr.c:638-823 */
static int match_exposures(uint32_t *raw32, int w, int h, int y1, int black20, int white_level20, const int is_bright[4],
                           double *corr_ev, int *white_darkened, double *out_a, double *out_b)
{
    int white20 = imin(white_level20, *white_darkened);
    int black = black20 / 16, white = white20 / 16;
    int clip0 = white - black, clip = (int)(clip0 * 0.95);
    int y0 = y1 + 2;
    size_t np = (size_t)w * h;
    int *dark = calloc(np, sizeof(int)), *bright = calloc(np, sizeof(int));
#define P16(x, y) ((int)((raw32[(x) + (size_t)(y) * w] >> 4) & 0xFFFF))
    for (int y = y0; y < h - 2; y += 3) {
        int *native = is_bright[y % 4] ? bright : dark, *interp = is_bright[y % 4] ? dark : bright;
        for (int x = 0; x < w; x += 3) {
            int pa = P16(x, y - 2) - black, pb = P16(x, y + 2) - black, pn = P16(x, y) - black;
            int pi = (pa + pb + 1) / 2;
            if (pa >= clip || pb >= clip) pi = clip0;
            if (pi >= clip) pn = clip0;
            interp[x + y * w] = pi;
            native[x + y * w] = pn;
        }
    }
#undef P16
    int nmax = (w + 2) * (h + 2) / 9;
    int *tmp = malloc(sizeof(int) * (size_t)imax(nmax, 1));
    int n = 0;
    for (int y = y0; y < h - 2; y += 3)
        for (int x = 0; x < w; x += 3)
            if (bright[x + y * w] < clip) tmp[n++] = bright[x + y * w];
    int bmed = lower_median(tmp, n);
    int b_lo = kth_smallest(tmp, n, n * 98 / 100);
    int b_hi = kth_smallest(tmp, n, (int)(n * 99.9 / 100));
    n = 0;
    for (int y = y0; y < h - 2; y += 3)
        for (int x = 0; x < w; x += 3)
            if (bright[x + y * w] < clip) tmp[n++] = dark[x + y * w];
    int dmed = lower_median(tmp, n);
    int hi_nmax = nmax / 50, hi_n = 0;
    int *hd = malloc(sizeof(int) * (size_t)(hi_nmax + h + 1)), *hb = malloc(sizeof(int) * (size_t)(hi_nmax + h + 1));
    for (int y = y0; y < h - 2; y += 3)
        for (int x = 0; x < w; x += 3) {
            int d = dark[x + y * w], b = bright[x + y * w];
            if (b >= b_hi || b <= b_lo) continue;
            hd[hi_n] = d; hb[hi_n] = b; hi_n++;
            if (hi_n >= hi_nmax) break;                         /* leaves only the row loop's inner level (hdr.c:744) */
        }
    double a = 0, b = 0;
    int best = 0;
    for (double ev = 0; ev < 6; ev += 0.002) {
        double ta = pow(2, -ev), tb = dmed - bmed * ta;
        int score = 0;
        for (int i = 0; i < hi_n; i++) {
            int e = (int)(hd[i] - (hb[i] * ta + tb));
            if (abs(e) < 50) score++;
        }
        if (score > best) { best = score; a = ta; b = tb; }
    }
    free(hd); free(hb); free(tmp); free(dark); free(bright);
    double b20 = b * 16;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int p = (int)raw32[x + (size_t)y * w];
            if (p == 0) continue;
            if (is_bright[y % 4]) p = (int)((p - black20) * a + black20 + b20 * a);
            else p = (int)(p - b20 + b20 * a);
            raw32[x + (size_t)y * w] = (uint32_t)iclamp(p, 0, 0xFFFFF);
        }
    *white_darkened = (int)((white20 - black20 + b20) * a + black20);
    *out_a = a; *out_b = b;
    double factor = 1 / a;
    if (factor < 1.2 || !isfinite(factor)) return 0;
    *corr_ev = log2(factor);
    return 1;
}

/* hdr.c:341-368 */
static int mean2(int a, int b, int white) { return (a >= white || b >= white) ? white : (a + b) / 2; }
static int mean3(int a, int b, int c, int white)
{
    int m = (a + b + c) / 3;
    return (a >= white || b >= white || c >= white) ? imax(m, white) : m;
}

/* hdr.c:1231-1304 */
static void mean23(const uint32_t *raw32, uint32_t *dark, uint32_t *bright, int w, int h, int white_level, int white_darkened,
                   const int is_bright[4], const int *raw2ev, const int *ev2raw)
{
#define R(x, y) ((int)raw32[(x) + (size_t)(y) * w])
    for (int y = 2; y < h - 2; y++) {
        int br = is_bright[y % 4];
        uint32_t *native = br ? bright : dark, *interp = br ? dark : bright;
        int white = !br ? white_darkened : white_level;
        int wev = raw2ev[white];
        int s = (is_bright[y % 4] == is_bright[(y + 1) % 4]) ? -1 : 1;
        for (int x = 2; x < w - 3; x += 2) {
            if (y % 2 == 0) {
                int ri = mean2(raw2ev[R(x, y - 2)], raw2ev[R(x, y + 2)], wev);
                int gi = mean3(raw2ev[R(x + 2, y + s)], raw2ev[R(x, y + s)], raw2ev[R(x + 1, y - 2 * s)], wev);
                interp[x + (size_t)y * w] = (uint32_t)ev2raw[ri];
                interp[x + 1 + (size_t)y * w] = (uint32_t)ev2raw[gi];
            } else {
                int bi = mean2(raw2ev[R(x + 1, y - 2)], raw2ev[R(x + 1, y + 2)], wev);
                int gi = mean3(raw2ev[R(x + 1, y + s)], raw2ev[R(x - 1, y + s)], raw2ev[R(x, y - 2 * s)], wev);
                interp[x + (size_t)y * w] = (uint32_t)ev2raw[gi];
                interp[x + 1 + (size_t)y * w] = (uint32_t)ev2raw[bi];
            }
            native[x + (size_t)y * w] = raw32[x + (size_t)y * w];
            native[x + 1 + (size_t)y * w] = raw32[x + 1 + (size_t)y * w];
        }
    }
#undef R
}

/* hdr.c:917-938: {ack, a, b, bck} offsets, y to be multiplied by s */
static const int edge_dirs[11][8] = {
    {-4, 2, -2, 1,  4, -2,  6, -3}, {-3, 2, -1, 1,  3, -2,  4, -3}, {-2, 2, -1, 1,  2, -2,  3, -3}, {-1, 2, -1, 1,  1, -2,  2, -3},
    {-1, 2,  0, 1,  1, -2,  1, -3}, { 0, 2,  0, 1,  0, -2,  0, -3}, { 1, 2,  0, 1, -1, -2, -1, -3}, { 1, 2,  1, 1, -1, -2, -2, -3},
    { 2, 2,  1, 1, -2, -2, -3, -3}, { 3, 2,  1, 1, -3, -2, -4, -3}, { 4, 2,  2, 1, -4, -2, -6, -3}};

/* hdr.c:940-952 */
static int edge_interp(const float *plane, int stride, const int *squeezed, const int *raw2ev, int dir, int x, int y, int s)
{
    int pa = iclamp((int)plane[(size_t)squeezed[y + edge_dirs[dir][3] * s] * stride + x + edge_dirs[dir][2]], 0, 0xFFFFF);
    int pb = iclamp((int)plane[(size_t)squeezed[y + edge_dirs[dir][5] * s] * stride + x + edge_dirs[dir][4]], 0, 0xFFFFF);
    return (raw2ev[pa] * 2 + raw2ev[pb]) / 3;
}

/* amaze_interpolate, hdr.c:954-1229 */
static void amaze_edge(const uint32_t *raw32, uint32_t *dark, uint32_t *bright, int w, int h, int black, int white_darkened,
                       const int is_bright[4], const int *raw2ev, const int *ev2raw, const double *curve, int fresh_tiles)
{
    const int ws = w + 16;
    size_t np = (size_t)w * h, nps = (size_t)ws * h;
    int *squeezed = calloc((size_t)h, sizeof(int));
    float *rawf = calloc(nps, sizeof(float)), *red = calloc(nps, sizeof(float)), *green = calloc(nps, sizeof(float)),
          *blue = calloc(nps, sizeof(float));
    /* squeeze: dark rows to the top, bright rows from h/4*2 on; greens halved (hdr.c:977-1026) */
    for (int pass = 0; pass < 2; pass++) {
        int yh = -1;
        for (int y = 0; y < h; y++) {
            if (is_bright[y % 4] != pass) continue;
            if (yh < 0) yh = pass ? h / 4 * 2 + y : y;
            for (int x = 0; x < w; x++) {
                int p = (int)raw32[x + (size_t)y * w];
                if (x % 2 != y % 2) p = (p - black) / 2 + black;
                rawf[(size_t)yh * ws + x] = p;
            }
            squeezed[y] = yh;
            yh++;
            if (pass && yh >= h) break;
        }
    }
    orc_amaze_demosaic(rawf, red, green, blue, ws, w, h, fresh_tiles);
    /* hdr.c:1045-1053 */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * ws + x;
            float g = (green[i] - black) * 2 + black;
            g = g < 0xFFFFF ? g : 0xFFFFF; green[i] = g > 0 ? g : 0;
            float r = red[i]; r = r < 0xFFFFF ? r : 0xFFFFF; red[i] = r > 0 ? r : 0;
            float b = blue[i]; b = b < 0xFFFFF ? b : 0xFFFFF; blue[i] = b > 0 ? b : 0;
        }
    uint32_t *gray = malloc(np * 4);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)squeezed[y] * ws + x;
            gray[x + (size_t)y * w] = (uint32_t)(green[i] / 2 + red[i] / 4 + blue[i] / 4);
        }
    uint8_t *edir = malloc(np);
    const int d0 = 5;
    memset(edir, d0, np);
    for (int y = 5; y < h - 5; y++) {
        int s = (is_bright[y % 4] == is_bright[(y + 1) % 4]) ? -1 : 1;
        for (int x = 5; x < w - 5; x++) {
            uint32_t p = raw32[x + (size_t)y * w];
            int search;
            if (!is_bright[y % 4]) search = !(curve[p] > fullres_thr);
            else search = !(p < (uint32_t)white_darkened);
            if (!search) continue;
            int e_best = 0x7FFFFFFF, d_best = d0;
            for (int d = 0; d <= 10; d++) {
                int e = 0;
                for (int j = -5; j <= 5; j++) {
                    int p1 = raw2ev[gray[x + edge_dirs[d][0] + j + (y + edge_dirs[d][1] * s) * w]];
                    int p2 = raw2ev[gray[x + edge_dirs[d][2] + j + (y + edge_dirs[d][3] * s) * w]];
                    int p3 = raw2ev[gray[x + edge_dirs[d][4] + j + (y + edge_dirs[d][5] * s) * w]];
                    int p4 = raw2ev[gray[x + edge_dirs[d][6] + j + (y + edge_dirs[d][7] * s) * w]];
                    e += abs(p1 - p2) + abs(p2 - p3) + abs(p3 - p4);
                }
                e += abs(d - d0) * EV / 8;
                if (e < e_best) { e_best = e; d_best = d; }
            }
            edir[x + (size_t)y * w] = (uint8_t)d_best;
        }
    }
    for (int y = 2; y < h - 2; y++) {
        int br = is_bright[y % 4];
        uint32_t *native = br ? bright : dark, *interp = br ? dark : bright;
        int is_rg = (y % 2 == 0);
        int s = (is_bright[y % 4] == is_bright[(y + 1) % 4]) ? -1 : 1;
        for (int x = 2; x < w - 2; x++) {
            const float *plane = is_rg ? (x % 2 == 0 ? red : green) : (x % 2 == 0 ? green : blue);
            int dir = edir[x + (size_t)y * w];
            int pi0 = edge_interp(plane, ws, squeezed, raw2ev, dir, x, y, s);
            int pip = edge_interp(plane, ws, squeezed, raw2ev, imin(dir + 1, 10), x, y, s);
            int pim = edge_interp(plane, ws, squeezed, raw2ev, imax(dir - 1, 0), x, y, s);
            interp[x + (size_t)y * w] = (uint32_t)ev2raw[(2 * pi0 + pip + pim) / 4];
            native[x + (size_t)y * w] = raw32[x + (size_t)y * w];
        }
    }
    free(edir); free(gray); free(squeezed); free(rawf); free(red); free(green); free(blue);
}

/* hdr.c:1306-1353 */
static void border(const uint32_t *raw32, uint32_t *dark, uint32_t *bright, int w, int h, const int is_bright[4])
{
#define SET(x, y, iv, nv) do { uint32_t *nat = is_bright[(y) % 4] ? bright : dark, *itp = is_bright[(y) % 4] ? dark : bright; \
                               itp[(x) + (size_t)(y) * w] = (iv); nat[(x) + (size_t)(y) * w] = (nv); } while (0)
#define R(x, y) raw32[(x) + (size_t)(y) * w]
    for (int y = 0; y < 3; y++)
        for (int x = 0; x < w; x++) SET(x, y, R(x, y + 2), R(x, y));
    for (int y = h - 4; y < h; y++)
        for (int x = 0; x < w; x++) SET(x, y, R(x, y - 2), R(x, y));
    for (int y = 2; y < h; y++) {
        for (int x = 0; x < 2; x++) SET(x, y, R(x, y - 2), R(x, y));
        for (int x = w - 3; x < w; x++) SET(x, y, R(x - 2, y - 2), R(x - 2, y));
    }
#undef R
#undef SET
}

/* hdr.c:1382-1486 */
static void alias_map_build(uint16_t *amap, const uint32_t *frs, const uint32_t *hrs, const uint32_t *bright, int w, int h,
                            int dark_noise, const double *curve, const int *raw2ev)
{
    size_t np = (size_t)w * h;
    uint16_t *aux = malloc(np * 2);
#define SKIP(i) (curve[bright[i]] > fullres_thr)
    for (size_t i = 0; i < np; i++) {
        if (SKIP(i)) continue;
        int f = (int)frs[i], hh = (int)hrs[i];
        int e_lin = imax(abs(f - hh) - dark_noise * 3 / 2, 0), e_log = abs(raw2ev[f] - raw2ev[hh]);
        amap[i] = (uint16_t)imin(imin(e_lin / 2, e_log / 16), 65530);
    }
    memcpy(aux, amap, np * 2);
    static const int ring[37][2] = {
        {-2,-6},{0,-6},{2,-6}, {-4,-4},{-2,-4},{0,-4},{2,-4},{4,-4},
        {-6,-2},{-4,-2},{-2,-2},{0,-2},{2,-2},{4,-2},{6,-2}, {-6,0},{-4,0},{-2,0},{0,0},{2,0},{4,0},{6,0},
        {-6,2},{-4,2},{-2,2},{0,2},{2,2},{4,2},{6,2}, {-4,4},{-2,4},{0,4},{2,4},{4,4}, {-2,6},{0,6},{2,6}};
    for (int y = 6; y < h - 6; y++)
        for (int x = 6; x < w - 6; x++) {
            size_t i = x + (size_t)y * w;
            if (SKIP(i)) continue;
            int nb[37];
            for (int k = 0; k < 37; k++) nb[k] = -(int)amap[(x + ring[k][0]) + (size_t)(y + ring[k][1]) * w];
            aux[i] = (uint16_t)(-kth_smallest(nb, 37, 5));       /* 6th largest */
        }
#define A(dx, dy) ((int)aux[(x + (dx)) + (size_t)(y + (dy)) * w])
    for (int y = 6; y < h - 6; y++)
        for (int x = 6; x < w - 6; x++) {
            size_t i = x + (size_t)y * w;
            if (SKIP(i)) continue;
            int cross = A(0,-2) + A(-2,0) + A(2,0) + A(0,2), diag = A(-2,-2) + A(2,-2) + A(-2,2) + A(2,2);
            int far = A(0,-6) + A(-6,0) + A(6,0) + A(0,6);
            int knight = A(-2,-6) + A(2,-6) + A(-6,-2) + A(6,-2) + A(-6,2) + A(6,2) + A(-2,6) + A(2,6);
            int c = A(0,0) + cross * 820 / 1024 + diag * 657 / 1024 + cross * 421 / 1024 + (2 * diag) * 337 / 1024 +
                    diag * 173 / 1024 + far * 139 / 1024 + knight * 111 / 1024 + knight * 57 / 1024;   /* sic, hdr.c:1451-1460 */
            amap[i] = (uint16_t)c;
        }
#undef A
#undef SKIP
    for (int y = 2; y < h - 2; y += 2)
        for (int x = 2; x < w - 2; x += 2) {
            uint16_t *p = amap + x + (size_t)y * w;
            int c = imin(imax(imax(p[0], p[1]), imax(p[w], p[w + 1])), ALIAS_MAP_MAX);
            p[0] = p[1] = p[w] = p[w + 1] = (uint16_t)c;
        }
    free(aux);
}

/* hdr_interpolate, hdr.c:1774-1930.  Returns 1 converted, 0 not dual ISO / failed, -1 unsupported. */
int orc_hdr_interpolate(uint16_t *image, int w, int h, int black14, int interp_method, int use_fullres, int use_alias_map,
                        int cs_method, orc_diso_state *S, orc_diso_info *info)
{
    if (w <= 0 || h <= 0) return 0;
    orc_diso_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof(*info));
    int rggb = identify_rggb(image, w, h), y1 = 0;
    info->rggb = rggb;
    if (!rggb) { image += w; h--; y1 = 1; }
    int is_bright[4];
    if (!identify_fields(image, w, h, black14, y1, is_bright)) return 0;
    memcpy(info->is_bright, is_bright, sizeof(is_bright));
    int black = black14 * 64, white, white_bright;
    white_detect(image, w, h, y1, is_bright, &white, &white_bright);
    white *= 64; white_bright *= 64;
    info->white_dark = white; info->white_bright = white_bright;
    const int dark_noise = 512;                      /* compute_noise sees an empty window: 8 DN * 64 */
    const double dark_noise_ev = 9;
    size_t np = (size_t)w * h;
    uint32_t *raw32 = malloc(np * 4);
    for (size_t i = 0; i < np; i++) raw32[i] = ((uint32_t)image[i] << 6) & 0xFFFFF;
    double corr_ev = 0, a = 0, b = 0;
    int white_darkened = white_bright, ret = 0;
    uint32_t *dark = calloc(np, 4), *bright = calloc(np, 4), *fullres = calloc(np, 4), *halfres = calloc(np, 4);
    uint32_t *frs = fullres, *hrs = halfres;
    uint16_t *over = calloc(np, 2), *amap = use_alias_map ? calloc(np, 2) : NULL;
    if (cs_method) {
        if (use_fullres) frs = malloc(np * 4);
        hrs = malloc(np * 4);
    }
    if (match_exposures(raw32, w, h, y1, black, white, is_bright, &corr_ev, &white_darkened, &a, &b)) {
        info->a = a; info->b = b; info->corr_ev = corr_ev; info->white_darkened = white_darkened;
        double lowiso_dr = log2(white - black) - dark_noise_ev;
        if (black != S->lut_black) build_luts(S, black, white);
        const int *raw2ev = S->raw2ev, *ev2raw = S->ev2raw_0 + 10 * EV;
        if (interp_method == 0)
            amaze_edge(raw32, dark, bright, w, h, black, white_darkened, is_bright, raw2ev, ev2raw, fullres_curve_for(S, black),
                       S->amaze_fresh_tiles);
        else
            mean23(raw32, dark, bright, w, h, white, white_darkened, is_bright, raw2ev, ev2raw);
        border(raw32, dark, bright, w, h, is_bright);
        if (use_fullres)                                              /* hdr.c:1355-1380 */
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    size_t i = x + (size_t)y * w;
                    if (is_bright[y % 4]) {
                        uint32_t f = bright[i];
                        fullres[i] = (int)f < white_darkened ? f : (f > dark[i] ? f : dark[i]);
                    } else fullres[i] = dark[i];
                }
        /* mix_images, hdr.c:1524-1661 */
        double overlap = lowiso_dr - corr_ev;
        overlap -= dmin(3, overlap - 3);
        info->overlap = overlap;
        if (!(overlap < 0.5)) {
            double max_ev = log2(white / 64 - black / 64);
            double *mix = malloc(sizeof(double) * N20);
            for (int i = 0; i < N20; i++) {
                double ev = log2(dmax(i / 64.0 - black / 64.0, 1)) + corr_ev;
                double c = -cos(dmax(dmin(ev - (max_ev - overlap), overlap), 0) * M_PI / overlap);
                mix[i] = (c + 1) / 2;
            }
            for (size_t i = 0; i < np; i++) {
                int bb = (int)bright[i], dd = (int)dark[i];
                double k = dmax(dmin(mix[bb & 0xFFFFF], 1), 0);
                int mixed = (int)(raw2ev[bb] * (1 - k) + raw2ev[dd] * k);
                halfres[i] = (uint32_t)ev2raw[mixed];
            }
            free(mix);
            if (cs_method) {
                memcpy(frs, fullres, np * 4);
                memcpy(hrs, halfres, np * 4);
                orc_chroma_smooth_u32(fullres, frs, w, h, cs_method, raw2ev, ev2raw);
                orc_chroma_smooth_u32(halfres, hrs, w, h, cs_method, raw2ev, ev2raw);
            }
            const double *curve = fullres_curve_for(S, black);
            if (amap) alias_map_build(amap, frs, hrs, bright, w, h, dark_noise, curve, raw2ev);
            for (size_t i = 0; i < np; i++) over[i] = ((int)bright[i] >= white_darkened || (int)dark[i] >= white) ? 100 : 0;
            uint16_t *oaux = malloc(np * 2);
            memcpy(oaux, over, np * 2);
#define O(dx, dy) ((int)oaux[(x + (dx)) + (size_t)(y + (dy)) * w])
            for (int y = 3; y < h - 3; y++)
                for (int x = 3; x < w - 3; x++)
                    over[x + (size_t)y * w] = (uint16_t)(O(0,0) + (O(0,-1) + O(-1,0) + O(1,0) + O(0,1)) * 820 / 1024 +
                                                         (O(-1,-1) + O(1,-1) + O(-1,1) + O(1,1)) * 657 / 1024);
#undef O
            free(oaux);
            /* final_blend, hdr.c:1663-1758 */
            for (size_t i = 0; i < np; i++) {
                int bb = (int)bright[i];
                int hrev = raw2ev[hrs[i]], frev = raw2ev[fullres[i]], frsev = raw2ev[frs[i]];
                double f = curve[bb & 0xFFFFF], c = 0;
                if (amap) c = dmax(dmin(amap[i] / (double)ALIAS_MAP_MAX, 1), 0);
                double ovf = dmax(dmin(over[i] / 200.0, 1), 0);
                c = dmax(c, ovf);
                double noo = dmax(ovf, 1 - f);
                f = dmax(f, c);
                double fev = noo * frsev + (1 - noo) * frev;
                int sig = (int)((dark[i] + bright[i]) / 2);
                f = dmax(0, dmin(f, (double)(sig - black) / (4 * dark_noise)));
                int out = (int)(hrev * (1 - f) + fev * f);
                out = iclamp(out, -10 * EV, 14 * EV - 1);
                raw32[i] = (uint32_t)ev2raw[out];
            }
            /* convert_20_to_16bit, hdr.c:1760-1772: the dither cache is never initialised (all zeros) */
            for (size_t i = 0; i < np; i++) image[i] = (uint16_t)iclamp((int)(raw32[i] / 16.0 + 0.0f + 0.5), 0, 0xFFFF);
            ret = 1;
        }
    } else {
        info->a = a; info->b = b;
    }
    free(dark); free(bright); free(fullres); free(halfres); free(over); free(raw32); free(amap);
    if (frs != fullres) free(frs);
    if (hrs != halfres) free(hrs);
    return ret;
}
