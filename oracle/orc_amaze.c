/* oracle/orc_amaze.c -- TEST INFRASTRUCTURE.  CPU restatement of the AMaZE demosaic as the reference
 * builds it on x86-64, i.e. the __SSE2__ variant of amaze_demosaic_RT.c:113-1487 (SSE2 is
 * always defined there, so the 4-lane code is the canonical behaviour, scalar twins are not).
 *
 * The vector loops are restated lane by lane with the reference's loop bounds, so that
 *   - lanes that run past the scalar bounds (cc += 4 / cc += 8 strides) write the same cells,
 *   - the in-place passes see the same mixture of already-updated and not-yet-updated neighbours
 *     (a vector iteration loads all its operands before it stores), and
 *   - work planes persist from tile to tile exactly like the reference's single calloc'ed block
 *     (amaze_demosaic_RT.c:244; only nyquist and rbint are cleared per tile, :294-295).
 * `fresh_tiles` != 0 instead clears the whole block before every tile; tests use it to show that the
 * output does not depend on the stale contents (which is what allows one CUDA block per tile).
 *
 * All arithmetic is IEEE binary32 in the reference's association order; the three places where the
 * reference's scalar code promotes to double (amaze_demosaic_RT.c:1297-1300, 1327, 1335) do so here too.
 * Build with -ffp-contract=off (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define TS  160                       /* amaze_demosaic_RT.c:137 */
#define TSH 80
enum { V1 = TS, V2 = 2 * TS, V3 = 3 * TS, P1 = -TS + 1, P2 = -2 * TS + 2, P3 = -3 * TS + 3,
       M1 = TS + 1, M2 = 2 * TS + 2, M3 = 3 * TS + 3 };                  /* :147 */

static const float eps = 1e-5f, epssq = 1e-10f, arthresh = 0.75f, nyqthresh = 0.5f;   /* :150-155 */
static const float clip_pt = 1.0f, clip_pt8 = 0.8f;                                     /* :133-134, initialGain 1 */
static const float gaussodd[4] = {0.14659727707323927f, 0.103592713382435f, 0.0732036125103057f, 0.0365543548389495f};
static const float gaussgrad[6] = {0.07384411893421103f, 0.06207511968171489f, 0.0521818194747806f,
                                   0.03687419286733595f, 0.03099732204057846f, 0.018413194161458882f};
static const float gausseven[2] = {0.13719494435797422f, 0.05640252782101291f};
static const float gquinc[4] = {0.169917f, 0.108947f, 0.069855f, 0.0287182f};           /* :158-165 */

static inline int fc(int r, int c) { return ((r & 1) == 0 && (c & 1) == 0) ? 0 : ((r & 1) && (c & 1)) ? 2 : 1; }   /* :41-49 */

/* exponent tricks, :88-100 (a zero stays zero, everything else has its exponent field decremented) */
static inline float expdec(float d, int n)
{
    union { float f; int32_t i; } u = { d };
    if (u.i & 0x7FFFFFFF) u.i -= n << 23;
    return u.f;
}
#define HALF(x) expdec((x), 1)
#define QUARTER(x) expdec((x), 2)

static inline float sq(float a) { return a * a; }
/* SSE semantics: minps/maxps return the second operand unless the strict comparison holds */
static inline float vmin(float a, float b) { return a < b ? a : b; }
static inline float vmax(float a, float b) { return a > b ? a : b; }
static inline float limv(float a, float b, float c) { return vmax(b, vmin(a, c)); }          /* sleefsseavx.c:1295 */
static inline float ulimv(float a, float b, float c) { return b < c ? limv(a, b, c) : limv(a, c, b); }   /* :1299 */
/* scalar macros, amaze_demosaic_RT.c:51-62,103-104 */
static inline float smin(float a, float b) { return a < b ? a : b; }
static inline float smax(float a, float b) { return a > b ? a : b; }
static inline float lims(float x, float lo, float hi) { return smax(smin(x, hi), lo); }
static inline float ulims(float a, float b, float c) { return b < c ? lims(a, b, c) : lims(a, c, b); }

/* the reference's block layout (:244-273): 64 spare bytes between consecutive planes */
typedef struct {
    char *block; size_t bytes;
    float *rgbgreen, *delhvsqsum, *dirwts0, *dirwts1, *vcd, *hcd, *vcdalt, *hcdalt, *cddiffsq, *hvwt, *Dgrb0, *Dgrb1,
          *delp, *delm, *rbint, *Dgrb2, *dgintv, *dginth, *Dgrbsq1m, *Dgrbsq1p, *cfa, *pmwt, *rbm, *rbp;
    char *nyquist;
} planes_t;

static void planes_alloc(planes_t *P)
{
    const size_t full = sizeof(float) * TS * TS, half = sizeof(float) * TS * TSH, gap = 64;
    P->bytes = 22 * full + TS * TSH + 23 * gap + 63;
    P->block = calloc(P->bytes, 1);
    char *p = (char *)(((uintptr_t)P->block + 63) / 64 * 64);
#define TAKE(name, sz) P->name = (float *)p; p += (sz) + gap
    TAKE(rgbgreen, full); TAKE(delhvsqsum, full); TAKE(dirwts0, full); TAKE(dirwts1, full); TAKE(vcd, full); TAKE(hcd, full);
    TAKE(vcdalt, full); TAKE(hcdalt, full); TAKE(cddiffsq, full); TAKE(hvwt, half);
    P->Dgrb0 = (float *)p; P->Dgrb1 = P->Dgrb0 + TS * TSH; p += full + gap;
    TAKE(delp, half); TAKE(delm, half); TAKE(rbint, half); TAKE(Dgrb2, full); TAKE(dgintv, full); TAKE(dginth, full);
    TAKE(Dgrbsq1m, half); TAKE(Dgrbsq1p, half); TAKE(cfa, full); TAKE(pmwt, half); TAKE(rbm, half); TAKE(rbp, half);
#undef TAKE
    P->nyquist = p;
}

/* one tile; raw/red/green/blue are row-major with `stride` floats per row (the reference's per-row
 * arrays of w+16 floats, hdr.c:967-975) */
static void amaze_tile(planes_t *P, const float *raw, float *red, float *green, float *blue, int stride,
                       int width, int height, int top, int left)
{
    float *rgbgreen = P->rgbgreen, *delhvsqsum = P->delhvsqsum, *dirwts0 = P->dirwts0, *dirwts1 = P->dirwts1, *vcd = P->vcd,
          *hcd = P->hcd, *vcdalt = P->vcdalt, *hcdalt = P->hcdalt, *cddiffsq = P->cddiffsq, *hvwt = P->hvwt, *delp = P->delp,
          *delm = P->delm, *rbint = P->rbint, *Dgrb2 = P->Dgrb2, *dgintv = P->dgintv, *dginth = P->dginth,
          *Dgrbsq1m = P->Dgrbsq1m, *Dgrbsq1p = P->Dgrbsq1p, *cfa = P->cfa, *pmwt = P->pmwt, *rbm = P->rbm, *rbp = P->rbp;
    float *Dgrb[2] = { P->Dgrb0, P->Dgrb1 };
    char *nyquist = P->nyquist;
#define RAW(r, c) raw[(size_t)(r) * stride + (c)]

    memset(nyquist, 0, TS * TSH);
    memset(rbint, 0, sizeof(float) * TS * TSH);
    const int bottom = top + TS < height + 16 ? top + TS : height + 16;
    const int right = left + TS < width + 16 ? left + TS : width + 16;
    const int rr1 = bottom - top, cc1 = right - left;
    const int rrmin = top < 0 ? 16 : 0, ccmin = left < 0 ? 16 : 0;
    const int rrmax = bottom > height ? height - top : rr1, ccmax = right > width ? width - left : cc1;
    int rr, cc, l;

    /* ---- tile load, value / 65535, :378-396 ---- */
    for (rr = rrmin; rr < rrmax; rr++) {
        for (cc = ccmin; cc < ccmax - 3; cc += 4)
            for (l = 0; l < 4; l++) cfa[rr * TS + cc + l] = rgbgreen[rr * TS + cc + l] = RAW(rr + top, cc + l + left) / 65535.0f;
        for (; cc < ccmax; cc++) {
            cfa[rr * TS + cc] = RAW(rr + top, cc + left) / 65535.0f;
            if (fc(rr, cc) == 1) rgbgreen[rr * TS + cc] = cfa[rr * TS + cc];
        }
    }
    /* ---- mirrored borders, :399-469 (vector variants copy 4 ascending source columns) ---- */
#define PUT_G(dst, val, r, c) do { cfa[dst] = (val); if (fc((r), (c)) == 1) rgbgreen[dst] = cfa[dst]; } while (0)
    if (rrmin > 0)
        for (rr = 0; rr < 16; rr++)
            for (cc = ccmin; cc < ccmax; cc++) PUT_G(rr * TS + cc, RAW(32 - rr + top, cc + left) / 65535.0f, rr, cc);
    if (rrmax < rr1)
        for (rr = 0; rr < 16; rr++)
            for (cc = ccmin; cc < ccmax; cc += 4)
                for (l = 0; l < 4; l++)
                    cfa[(rrmax + rr) * TS + cc + l] = rgbgreen[(rrmax + rr) * TS + cc + l] = RAW(height - rr - 2, left + cc + l) / 65535.0f;
    if (ccmin > 0)
        for (rr = rrmin; rr < rrmax; rr++)
            for (cc = 0; cc < 16; cc++) PUT_G(rr * TS + cc, RAW(rr + top, 32 - cc + left) / 65535.0f, rr, cc);
    if (ccmax < cc1)
        for (rr = rrmin; rr < rrmax; rr++)
            for (cc = 0; cc < 16; cc++) PUT_G(rr * TS + ccmax + cc, RAW(top + rr, width - cc - 2) / 65535.0f, rr, cc);
    if (rrmin > 0 && ccmin > 0)
        for (rr = 0; rr < 16; rr++)
            for (cc = 0; cc < 16; cc += 4)
                for (l = 0; l < 4; l++) cfa[rr * TS + cc + l] = rgbgreen[rr * TS + cc + l] = RAW(32 - rr, 32 - cc + l) / 65535.0f;
    if (rrmax < rr1 && ccmax < cc1)
        for (rr = 0; rr < 16; rr++)
            for (cc = 0; cc < 16; cc += 4)
                for (l = 0; l < 4; l++)
                    cfa[(rrmax + rr) * TS + ccmax + cc + l] = rgbgreen[(rrmax + rr) * TS + ccmax + cc + l] =
                        RAW(height - rr - 2, width - cc - 2 + l) / 65535.0f;
    if (rrmin > 0 && ccmax < cc1)
        for (rr = 0; rr < 16; rr++)
            for (cc = 0; cc < 16; cc++) PUT_G(rr * TS + ccmax + cc, RAW(32 - rr, width - cc - 2) / 65535.0f, rr, cc);
    if (rrmax < rr1 && ccmin > 0)
        for (rr = 0; rr < 16; rr++)
            for (cc = 0; cc < 16; cc++) PUT_G((rrmax + rr) * TS + cc, RAW(height - rr - 2, 32 - cc) / 65535.0f, rr, cc);
#undef PUT_G

    /* ---- horizontal / vertical gradients and directional weights, :553-567 ---- */
    for (rr = 2; rr < rr1 - 2; rr++)
        for (cc = 0; cc < cc1; cc += 4)
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + l;
                const float delh = fabsf(cfa[i + 1] - cfa[i - 1]), delv = fabsf(cfa[i + V1] - cfa[i - V1]);
                dirwts1[i] = eps + fabsf(cfa[i + 2] - cfa[i]) + fabsf(cfa[i] - cfa[i - 2]) + delh;
                dirwts0[i] = eps + fabsf(cfa[i + V2] - cfa[i]) + fabsf(cfa[i] - cfa[i - V2]) + delv;
                delhvsqsum[i] = delh * delh + delv * delv;
            }

    /* ---- diagonal gradients (at R/B sites) and squared diagonal green differences (at G sites), :581-606 ---- */
    for (rr = 6; rr < rr1 - 6; rr++) {
        const int o = (fc(rr, 2) & 1) ? 0 : 1;         /* offset of the green site inside the pair starting at even cc */
        for (cc = 6; cc < cc1 - 6; cc += 8)
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + 2 * l, g = i + o, c = i + (1 - o);
                const float t = cfa[g];
                const float dp = sq(t - cfa[g - P1]) + sq(t - cfa[g + P1]);
                delp[i >> 1] = fabsf(cfa[c + P1] - cfa[c - P1]);
                delm[i >> 1] = fabsf(cfa[c + M1] - cfa[c - M1]);
                Dgrbsq1m[i >> 1] = sq(t - cfa[g - M1]) + sq(t - cfa[g + M1]);
                Dgrbsq1p[i >> 1] = dp;
            }
    }

    /* ---- horizontal / vertical colour differences, adaptive-ratio vs Hamilton-Adams, :633-689 ---- */
    for (rr = 4; rr < rr1 - 4; rr++)
        for (cc = 4; cc < cc1 - 7; cc += 4)
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + l;
                const float sgn = ((rr + cc + l) & 1) ? -1.0f : 1.0f;        /* +1 on R/B sites, -1 on G sites */
                const float c0 = cfa[i];
                const float cru = cfa[i - V1] * (dirwts0[i - V2] + dirwts0[i]) / (dirwts0[i - V2] * (eps + c0) + dirwts0[i] * (eps + cfa[i - V2]));
                const float crd = cfa[i + V1] * (dirwts0[i + V2] + dirwts0[i]) / (dirwts0[i + V2] * (eps + c0) + dirwts0[i] * (eps + cfa[i + V2]));
                const float crl = cfa[i - 1] * (dirwts1[i - 2] + dirwts1[i]) / (dirwts1[i - 2] * (eps + c0) + dirwts1[i] * (eps + cfa[i - 2]));
                const float crr = cfa[i + 1] * (dirwts1[i + 2] + dirwts1[i]) / (dirwts1[i + 2] * (eps + c0) + dirwts1[i] * (eps + cfa[i + 2]));
                const float guha = cfa[i - V1] + 0.5f * (c0 - cfa[i - V2]), gdha = cfa[i + V1] + 0.5f * (c0 - cfa[i + V2]);
                const float glha = cfa[i - 1] + 0.5f * (c0 - cfa[i - 2]), grha = cfa[i + 1] + 0.5f * (c0 - cfa[i + 2]);
                float guar = fabsf(1.0f - cru) < arthresh ? c0 * cru : guha;
                float gdar = fabsf(1.0f - crd) < arthresh ? c0 * crd : gdha;
                float glar = fabsf(1.0f - crl) < arthresh ? c0 * crl : glha;
                float grar = fabsf(1.0f - crr) < arthresh ? c0 * crr : grha;
                const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
                const float vwt = dirwts0[i - V1] / (dirwts0[i + V1] + dirwts0[i - V1]);
                const float Ginthha = hwt * grha + (1.0f - hwt) * glha, Gintvha = vwt * gdha + (1.0f - vwt) * guha;
                const float ha = sgn * (Ginthha - c0), va = sgn * (Gintvha - c0);
                hcdalt[i] = ha; vcdalt[i] = va;
                const int clip = (c0 > clip_pt8) | (Gintvha > clip_pt8) | (Ginthha > clip_pt8);
                if (clip) { guar = guha; gdar = gdha; glar = glha; grar = grha; }
                vcd[i] = clip ? va : sgn * ((vwt * gdar + (1.0f - vwt) * guar) - c0);
                hcd[i] = clip ? ha : sgn * ((hwt * grar + (1.0f - hwt) * glar) - c0);
                dgintv[i] = vmin(sq(guha - gdha), sq(guar - gdar));
                dginth[i] = vmin(sq(glha - grha), sq(glar - grar));
            }

    /* ---- pick the smoother of the two estimates, bound it in saturated regions; in place, :748-803 ---- */
    for (rr = 4; rr < rr1 - 4; rr++)
        for (cc = 4; cc < cc1 - 4; cc += 4) {
            float hnew[4], vnew[4];
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + l;
                const float sgn = ((rr + cc + l) & 1) ? -1.0f : 1.0f, nsgn = -sgn, sgn3 = 3.0f * sgn;
                const float c0 = cfa[i];
                float h = hcd[i], v = vcd[i];
                const float hvar = 3.0f * (sq(hcd[i - 2]) + sq(h) + sq(hcd[i + 2])) - sq(hcd[i - 2] + h + hcd[i + 2]);
                const float ha = hcdalt[i];
                const float havar = 3.0f * (sq(hcdalt[i - 2]) + sq(ha) + sq(hcdalt[i + 2])) - sq(hcdalt[i - 2] + ha + hcdalt[i + 2]);
                const float vvar = 3.0f * (sq(vcd[i - V2]) + sq(v) + sq(vcd[i + V2])) - sq(vcd[i - V2] + v + vcd[i + V2]);
                const float va = vcdalt[i];
                const float vavar = 3.0f * (sq(vcdalt[i - V2]) + sq(va) + sq(vcdalt[i + V2])) - sq(vcdalt[i - V2] + va + vcdalt[i + V2]);
                if (havar < hvar) h = ha;
                if (vavar < vvar) v = va;

                const float Ginth = sgn * h + c0;
                float t2 = sgn3 * h;
                const float hw = 1.0f + t2 / (eps + Ginth + c0);
                const int hpos = nsgn * h > 0.0f;
                const float hold = h;
                float t = nsgn * (c0 - ulimv(Ginth, cfa[i - 1], cfa[i + 1]));
                h = (t2 < -(c0 + Ginth)) ? t : hw * h + (1.0f - hw) * t;
                h = hpos ? h : hold;
                h = Ginth > clip_pt ? t : h;

                const float Gintv = sgn * v + c0;
                t2 = sgn3 * v;
                const float vw = 1.0f + t2 / (eps + Gintv + c0);
                const int vpos = nsgn * v > 0.0f;
                const float vold = v;
                t = nsgn * (c0 - ulimv(Gintv, cfa[i - V1], cfa[i + V1]));
                v = (t2 < -(c0 + Gintv)) ? t : vw * v + (1.0f - vw) * t;
                v = vpos ? v : vold;
                v = Gintv > clip_pt ? t : v;
                hnew[l] = h; vnew[l] = v;
            }
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + l;
                hcd[i] = hnew[l]; vcd[i] = vnew[l];
                cddiffsq[i] = sq(vnew[l] - hnew[l]);
            }
        }

    /* ---- horizontal vs vertical weight from colour-difference variances, :876-920 ---- */
    for (rr = 6; rr < rr1 - 6; rr++)
        for (cc = 6 + (fc(rr, 2) & 1); cc < cc1 - 6; cc += 8)
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + 2 * l;
                float t = vcd[i];
                const float uave = t + vcd[i - V1] + vcd[i - V2] + vcd[i - V3], dave = t + vcd[i + V1] + vcd[i + V2] + vcd[i + V3];
                float Du = sq(t - uave) + sq(vcd[i - V1] - uave) + sq(vcd[i - V2] - uave) + sq(vcd[i - V3] - uave);
                float Dd = sq(t - dave) + sq(vcd[i + V1] - dave) + sq(vcd[i + V2] - dave) + sq(vcd[i + V3] - dave);
                const float hwt = dirwts1[i - 1] / (dirwts1[i - 1] + dirwts1[i + 1]);
                const float vwt = dirwts0[i - V1] / (dirwts0[i + V1] + dirwts0[i - V1]);
                t = hcd[i];
                const float lave = t + hcd[i - 1] + hcd[i - 2] + hcd[i - 3], rave = t + hcd[i + 1] + hcd[i + 2] + hcd[i + 3];
                float Dl = sq(t - lave) + sq(hcd[i - 1] - lave) + sq(hcd[i - 2] - lave) + sq(hcd[i - 3] - lave);
                float Dr = sq(t - rave) + sq(hcd[i + 1] - rave) + sq(hcd[i + 2] - rave) + sq(hcd[i + 3] - rave);
                const float vcdvar = epssq + vwt * Dd + (1.0f - vwt) * Du, hcdvar = epssq + hwt * Dr + (1.0f - hwt) * Dl;
                Du = dgintv[i] + dgintv[i - V1] + dgintv[i - V2];
                Dd = dgintv[i] + dgintv[i + V1] + dgintv[i + V2];
                Dl = dginth[i] + dginth[i - 1] + dginth[i - 2];
                Dr = dginth[i] + dginth[i + 1] + dginth[i + 2];
                const float vcdvar1 = epssq + vwt * Dd + (1.0f - vwt) * Du, hcdvar1 = epssq + hwt * Dr + (1.0f - hwt) * Dl;
                const float varwt = hcdvar / (vcdvar + hcdvar), diffwt = hcdvar1 / (vcdvar1 + hcdvar1);
                const int dec = ((0.5f - varwt) * (0.5f - diffwt) > 0.0f) & (fabsf(0.5f - diffwt) < fabsf(0.5f - varwt));
                hvwt[i >> 1] = dec ? varwt : diffwt;
            }

    /* ---- Nyquist texture test, :967-996 ---- */
    for (rr = 6; rr < rr1 - 6; rr++)
        for (cc = 6 + (fc(rr, 2) & 1); cc < cc1 - 6; cc += 2) {
            const int i = rr * TS + cc;
            float nyqtest = (gaussodd[0] * cddiffsq[i] +
                             gaussodd[1] * (cddiffsq[i - M1] + cddiffsq[i + P1] + cddiffsq[i - P1] + cddiffsq[i + M1]) +
                             gaussodd[2] * (cddiffsq[i - V2] + cddiffsq[i - 2] + cddiffsq[i + 2] + cddiffsq[i + V2]) +
                             gaussodd[3] * (cddiffsq[i - M2] + cddiffsq[i + P2] + cddiffsq[i - P2] + cddiffsq[i + M2]));
            nyqtest -= nyqthresh * (gaussgrad[0] * delhvsqsum[i] +
                                    gaussgrad[1] * (delhvsqsum[i - V1] + delhvsqsum[i + 1] + delhvsqsum[i - 1] + delhvsqsum[i + V1]) +
                                    gaussgrad[2] * (delhvsqsum[i - M1] + delhvsqsum[i + P1] + delhvsqsum[i - P1] + delhvsqsum[i + M1]) +
                                    gaussgrad[3] * (delhvsqsum[i - V2] + delhvsqsum[i - 2] + delhvsqsum[i + 2] + delhvsqsum[i + V2]) +
                                    gaussgrad[4] * (delhvsqsum[i - 2 * TS - 1] + delhvsqsum[i - 2 * TS + 1] + delhvsqsum[i - TS - 2] +
                                                    delhvsqsum[i - TS + 2] + delhvsqsum[i + TS - 2] + delhvsqsum[i + TS + 2] +
                                                    delhvsqsum[i + 2 * TS - 1] + delhvsqsum[i + 2 * TS + 1]) +
                                    gaussgrad[5] * (delhvsqsum[i - M2] + delhvsqsum[i + P2] + delhvsqsum[i - P2] + delhvsqsum[i + M2]));
            if (nyqtest > 0) nyquist[i >> 1] = 1;
        }
    /* 3x3 majority vote, sequential and in place, :998-1010 */
    for (rr = 8; rr < rr1 - 8; rr++)
        for (cc = 8 + (fc(rr, 2) & 1); cc < cc1 - 8; cc += 2) {
            const int i = rr * TS + cc;
            const unsigned n = nyquist[(i - V2) >> 1] + nyquist[(i - M1) >> 1] + nyquist[(i + P1) >> 1] + nyquist[(i - 2) >> 1] +
                               nyquist[i >> 1] + nyquist[(i + 2) >> 1] + nyquist[(i - P1) >> 1] + nyquist[(i + M1) >> 1] +
                               nyquist[(i + V2) >> 1];
            if (n > 4) nyquist[i >> 1] = 1;
            if (n < 4) nyquist[i >> 1] = 0;
        }
    /* area interpolation inside Nyquist regions, :1016-1045 */
    for (rr = 8; rr < rr1 - 8; rr++)
        for (cc = 8 + (fc(rr, 2) & 1); cc < cc1 - 8; cc += 2) {
            const int i = rr * TS + cc;
            if (!nyquist[i >> 1]) continue;
            float sumh = 0, sumv = 0, sumsqh = 0, sumsqv = 0, areawt = 0;
            for (int a = -6; a < 7; a += 2)
                for (int b = -6; b < 7; b += 2) {
                    const int j = (rr + a) * TS + cc + b;
                    if (!nyquist[j >> 1]) continue;
                    sumh += cfa[j] - HALF(cfa[j - 1] + cfa[j + 1]);
                    sumv += cfa[j] - HALF(cfa[j - V1] + cfa[j + V1]);
                    sumsqh += HALF(sq(cfa[j] - cfa[j - 1]) + sq(cfa[j] - cfa[j + 1]));
                    sumsqv += HALF(sq(cfa[j] - cfa[j - V1]) + sq(cfa[j] - cfa[j + V1]));
                    areawt += 1;
                }
            const float hv = epssq + fabsf(areawt * sumsqh - sumh * sumh), vv = epssq + fabsf(areawt * sumsqv - sumv * sumv);
            hvwt[i >> 1] = hv / (vv + hv);
        }

    /* ---- G at R/B sites; hvwt is refined in place row after row, :1050-1075 ---- */
    for (rr = 8; rr < rr1 - 8; rr++)
        for (cc = 8 + (fc(rr, 2) & 1); cc < cc1 - 8; cc += 2) {
            const int i = rr * TS + cc;
            const float alt = QUARTER(hvwt[(i - M1) >> 1] + hvwt[(i + P1) >> 1] + hvwt[(i - P1) >> 1] + hvwt[(i + M1) >> 1]);
            if (fabsf(0.5f - hvwt[i >> 1]) < fabsf(0.5f - alt)) hvwt[i >> 1] = alt;
            Dgrb[0][i >> 1] = hcd[i] * (1.0f - hvwt[i >> 1]) + vcd[i] * hvwt[i >> 1];
            rgbgreen[i] = cfa[i] + Dgrb[0][i >> 1];
            if (nyquist[i >> 1]) {
                Dgrb2[2 * (i >> 1)] = sq(rgbgreen[i] - HALF(rgbgreen[i - 1] + rgbgreen[i + 1]));
                Dgrb2[2 * (i >> 1) + 1] = sq(rgbgreen[i] - HALF(rgbgreen[i - V1] + rgbgreen[i + V1]));
            } else
                Dgrb2[2 * (i >> 1)] = Dgrb2[2 * (i >> 1) + 1] = 0;
        }
    /* refine Nyquist sites with the local G curvature, :1085-1102 */
#define D2H(k) Dgrb2[2 * ((k) >> 1)]
#define D2V(k) Dgrb2[2 * ((k) >> 1) + 1]
    for (rr = 8; rr < rr1 - 8; rr++)
        for (cc = 8 + (fc(rr, 2) & 1); cc < cc1 - 8; cc += 2) {
            const int i = rr * TS + cc;
            if (!nyquist[i >> 1]) continue;
            const float gvarh = epssq + (gquinc[0] * D2H(i) + gquinc[1] * (D2H(i - M1) + D2H(i + P1) + D2H(i - P1) + D2H(i + M1)) +
                                         gquinc[2] * (D2H(i - V2) + D2H(i - 2) + D2H(i + 2) + D2H(i + V2)) +
                                         gquinc[3] * (D2H(i - M2) + D2H(i + P2) + D2H(i - P2) + D2H(i + M2)));
            const float gvarv = epssq + (gquinc[0] * D2V(i) + gquinc[1] * (D2V(i - M1) + D2V(i + P1) + D2V(i - P1) + D2V(i + M1)) +
                                         gquinc[2] * (D2V(i - V2) + D2V(i - 2) + D2V(i + 2) + D2V(i + V2)) +
                                         gquinc[3] * (D2V(i - M2) + D2V(i + P2) + D2V(i - P2) + D2V(i + M2)));
            Dgrb[0][i >> 1] = (hcd[i] * gvarv + vcd[i] * gvarh) / (gvarv + gvarh);
            rgbgreen[i] = cfa[i] + Dgrb[0][i >> 1];
        }
#undef D2H
#undef D2V

    /* ---- diagonal (NW-SE "m", NE-SW "p") interpolation of the opposite colour at R/B sites, :1115-1180 ---- */
    for (rr = 8; rr < rr1 - 8; rr++)
        for (cc = 8 + (fc(rr, 2) & 1); cc < cc1 - 8; cc += 8)
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + 2 * l, i1 = i >> 1;
                const float c0 = cfa[i];
                float t1, t2, w, rbse, rbnw, rbne, rbsw;
                t1 = cfa[i + M1]; t2 = cfa[i + M2]; rbse = (t1 + t1) / (eps + c0 + t2);
                rbse = fabsf(1.0f - rbse) < arthresh ? c0 * rbse : t1 + 0.5f * (c0 - t2);
                t1 = cfa[i - M1]; t2 = cfa[i - M2]; rbnw = (t1 + t1) / (eps + c0 + t2);
                rbnw = fabsf(1.0f - rbnw) < arthresh ? c0 * rbnw : t1 + 0.5f * (c0 - t2);
                t1 = eps + delm[i1];
                const float wtse = t1 + delm[(i + M1) >> 1] + delm[(i + M2) >> 1], wtnw = t1 + delm[(i - M1) >> 1] + delm[(i - M2) >> 1];
                const float m = (wtse * rbnw + wtnw * rbse) / (wtse + wtnw);
                t1 = ulimv(m, cfa[i - M1], cfa[i + M1]);
                w = 2.0f * (c0 - m) / (eps + m + c0);
                t2 = w * m + (1.0f - w) * t1;
                t2 = (m + m < c0) ? t1 : t2;
                t2 = (m < c0) ? t2 : m;
                rbm[i1] = t2 > clip_pt ? ulimv(t2, cfa[i - M1], cfa[i + M1]) : t2;

                t1 = cfa[i + P1]; t2 = cfa[i + P2]; rbne = (t1 + t1) / (eps + c0 + t2);
                rbne = fabsf(1.0f - rbne) < arthresh ? c0 * rbne : t1 + 0.5f * (c0 - t2);
                t1 = cfa[i - P1]; t2 = cfa[i - P2]; rbsw = (t1 + t1) / (eps + c0 + t2);
                rbsw = fabsf(1.0f - rbsw) < arthresh ? c0 * rbsw : t1 + 0.5f * (c0 - t2);
                t1 = eps + delp[i1];
                const float wtne = t1 + delp[(i + P1) >> 1] + delp[(i + P2) >> 1], wtsw = t1 + delp[(i - P1) >> 1] + delp[(i - P2) >> 1];
                const float p = (wtne * rbsw + wtsw * rbne) / (wtne + wtsw);
                t1 = ulimv(p, cfa[i - P1], cfa[i + P1]);
                w = 2.0f * (c0 - p) / (eps + p + c0);
                t2 = w * p + (1.0f - w) * t1;
                t2 = (p + p < c0) ? t1 : t2;
                t2 = (p < c0) ? t2 : p;
                rbp[i1] = t2 > clip_pt ? ulimv(t2, cfa[i - P1], cfa[i + P1]) : t2;

#define EVEN8(A) (gausseven[0] * (A[(i - V1) >> 1] + A[(i - 1) >> 1] + A[(i + 1) >> 1] + A[(i + V1) >> 1]) +                      \
                  gausseven[1] * (A[(i - V2 - 1) >> 1] + A[(i - V2 + 1) >> 1] + A[(i - 2 - V1) >> 1] + A[(i + 2 - V1) >> 1] +       \
                                  A[(i - 2 + V1) >> 1] + A[(i + 2 + V1) >> 1] + A[(i + V2 - 1) >> 1] + A[(i + V2 + 1) >> 1]))
                const float rbvarm = epssq + EVEN8(Dgrbsq1m);
                pmwt[i1] = rbvarm / ((epssq + EVEN8(Dgrbsq1p)) + rbvarm);
#undef EVEN8
            }

    /* ---- plus/minus weight refined in place row after row; R+B estimate, :1264-1274 ---- */
    for (rr = 10; rr < rr1 - 10; rr++)
        for (cc = 10 + (fc(rr, 2) & 1); cc < cc1 - 10; cc += 8) {
            float wn[4], rn[4];
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + 2 * l, i1 = i >> 1;
                const float alt = 0.25f * (pmwt[(i - M1) >> 1] + pmwt[(i + P1) >> 1] + pmwt[(i - P1) >> 1] + pmwt[(i + M1) >> 1]);
                float t = pmwt[i1];
                t = fabsf(0.5f - t) < fabsf(0.5f - alt) ? alt : t;
                wn[l] = t;
                rn[l] = 0.5f * (cfa[i] + rbm[i1] * (1.0f - t) + rbp[i1] * t);
            }
            for (l = 0; l < 4; l++) {
                const int i1 = (rr * TS + cc + 2 * l) >> 1;
                pmwt[i1] = wn[l]; rbint[i1] = rn[l];
            }
        }

    /* ---- where the diagonal estimate discriminates better, redo G from R+B, :1287-1352 ---- */
    for (rr = 12; rr < rr1 - 12; rr++)
        for (cc = 12 + (fc(rr, 2) & 1); cc < cc1 - 12; cc += 2) {
            const int i = rr * TS + cc, i1 = i >> 1;
            if (fabsf(0.5f - pmwt[i1]) < fabsf(0.5f - hvwt[i1])) continue;
            const float rb = rbint[i1];
            const float cru = (float)(cfa[i - V1] * 2.0 / (eps + rb + rbint[i1 - V1]));
            const float crd = (float)(cfa[i + V1] * 2.0 / (eps + rb + rbint[i1 + V1]));
            const float crl = (float)(cfa[i - 1] * 2.0 / (eps + rb + rbint[i1 - 1]));
            const float crr = (float)(cfa[i + 1] * 2.0 / (eps + rb + rbint[i1 + 1]));
            const float gu = fabsf(1.0f - cru) < arthresh ? rb * cru : cfa[i - V1] + HALF(rb - rbint[i1 - V1]);
            const float gd = fabsf(1.0f - crd) < arthresh ? rb * crd : cfa[i + V1] + HALF(rb - rbint[i1 + V1]);
            const float gl = fabsf(1.0f - crl) < arthresh ? rb * crl : cfa[i - 1] + HALF(rb - rbint[i1 - 1]);
            const float gr = fabsf(1.0f - crr) < arthresh ? rb * crr : cfa[i + 1] + HALF(rb - rbint[i1 + 1]);
            float Gintv = (dirwts0[i - V1] * gd + dirwts0[i + V1] * gu) / (dirwts0[i + V1] + dirwts0[i - V1]);
            float Ginth = (dirwts1[i - 1] * gr + dirwts1[i + 1] * gl) / (dirwts1[i - 1] + dirwts1[i + 1]);
            if (Gintv < rb) {
                if (2 * Gintv < rb)
                    Gintv = ulims(Gintv, cfa[i - V1], cfa[i + V1]);
                else {
                    const float vw = (float)(2.0 * (rb - Gintv) / (eps + Gintv + rb));
                    Gintv = vw * Gintv + (1.0f - vw) * ulims(Gintv, cfa[i - V1], cfa[i + V1]);
                }
            }
            if (Ginth < rb) {
                if (2 * Ginth < rb)
                    Ginth = ulims(Ginth, cfa[i - 1], cfa[i + 1]);
                else {
                    const float hw = (float)(2.0 * (rb - Ginth) / (eps + Ginth + rb));
                    Ginth = hw * Ginth + (1.0f - hw) * ulims(Ginth, cfa[i - 1], cfa[i + 1]);
                }
            }
            if (Ginth > clip_pt) Ginth = ulims(Ginth, cfa[i - 1], cfa[i + 1]);
            if (Gintv > clip_pt) Gintv = ulims(Gintv, cfa[i - V1], cfa[i + V1]);
            rgbgreen[i] = Ginth * (1.0f - hvwt[i1]) + Gintv * hvwt[i1];
            Dgrb[0][i1] = rgbgreen[i] - cfa[i];
        }

    /* ---- split G-B out of the G-R plane (R at (0,0): B sites are odd/odd), :1358-1362 ---- */
    for (rr = 13; rr < rr1 - 12; rr += 2)
        for (cc = 13; cc < cc1 - 12; cc += 2) {
            const int i1 = (rr * TS + cc) >> 1;
            Dgrb[1][i1] = Dgrb[0][i1];
            Dgrb[0][i1] = 0;
        }
    /* ---- chroma at the opposite-colour sites from the four diagonal neighbours, :1369-1383 ---- */
    for (rr = 14; rr < rr1 - 14; rr++) {
        const int cs = 14 + (fc(rr, 2) & 1);
        float *D = Dgrb[1 - fc(rr, cs) / 2];
        for (cc = cs; cc < cc1 - 14; cc += 8) {
            float res[4];
            for (l = 0; l < 4; l++) {
                const int i = rr * TS + cc + 2 * l;
#define G(o) D[(i + (o)) >> 1]
                const float wtnw = 1.0f / (eps + fabsf(G(-M1) - G(M1)) + fabsf(G(-M1) - G(-M3)) + fabsf(G(M1) - G(-M3)));
                const float wtne = 1.0f / (eps + fabsf(G(P1) - G(-P1)) + fabsf(G(P1) - G(P3)) + fabsf(G(-P1) - G(P3)));
                const float wtsw = 1.0f / (eps + fabsf(G(-P1) - G(P1)) + fabsf(G(-P1) - G(M3)) + fabsf(G(P1) - G(-P3)));
                const float wtse = 1.0f / (eps + fabsf(G(M1) - G(-M1)) + fabsf(G(M1) - G(-P3)) + fabsf(G(-M1) - G(M3)));
                res[l] = (wtnw * (1.325f * G(-M1) - 0.175f * G(-M3) - 0.075f * G(-M1 - 2) - 0.075f * G(-M1 - V2)) +
                          wtne * (1.325f * G(P1) - 0.175f * G(P3) - 0.075f * G(P1 + 2) - 0.075f * G(P1 + V2)) +
                          wtsw * (1.325f * G(-P1) - 0.175f * G(-P3) - 0.075f * G(-P1 - 2) - 0.075f * G(-P1 - V2)) +
                          wtse * (1.325f * G(M1) - 0.175f * G(M3) - 0.075f * G(M1 + 2) - 0.075f * G(M1 + V2))) /
                         (wtnw + wtne + wtsw + wtse);
#undef G
            }
            for (l = 0; l < 4; l++) D[(rr * TS + cc + 2 * l) >> 1] = res[l];
        }
    }

    /* ---- red and blue planes, :1400-1445 ---- */
#define CROSS(k, i) ((hvwt[((i) - V1) >> 1]) * Dgrb[k][((i) - V1) >> 1] + (1.0f - hvwt[((i) + 1) >> 1]) * Dgrb[k][((i) + 1) >> 1] + \
                     (1.0f - hvwt[((i) - 1) >> 1]) * Dgrb[k][((i) - 1) >> 1] + (hvwt[((i) + V1) >> 1]) * Dgrb[k][((i) + V1) >> 1])
    for (rr = 16; rr < rr1 - 16; rr++) {
        const int row = rr + top, gfirst = (fc(rr, 2) & 1);      /* does the pair at even cc start with a green site? */
        float *R = red + (size_t)row * stride, *B = blue + (size_t)row * stride;
        for (cc = 16; cc < cc1 - 16; cc++) {
            const int i = rr * TS + cc, col = cc + left;
            const int is_green = ((cc & 1) == 0) ? gfirst : !gfirst;
            if (is_green) {
                const float temp = 1.0f / ((hvwt[(i - V1) >> 1]) + (1.0f - hvwt[(i + 1) >> 1]) + (1.0f - hvwt[(i - 1) >> 1]) + (hvwt[(i + V1) >> 1]));
                R[col] = 65535.0f * (rgbgreen[i] - CROSS(0, i) * temp);
                B[col] = 65535.0f * (rgbgreen[i] - CROSS(1, i) * temp);
            } else {
                R[col] = 65535.0f * (rgbgreen[i] - Dgrb[0][i >> 1]);
                B[col] = 65535.0f * (rgbgreen[i] - Dgrb[1][i >> 1]);
            }
        }
    }
#undef CROSS
    /* ---- green plane, 4 columns per store while cc < cc1-19, :1451-1455 ---- */
    for (rr = 16; rr < rr1 - 16; rr++)
        for (cc = 16; cc < cc1 - 19; cc += 4)
            for (l = 0; l < 4; l++) green[(size_t)(rr + top) * stride + cc + left + l] = rgbgreen[rr * TS + cc + l] * 65535.0f;
#undef RAW
}

/* amaze_demosaic_RT(rawData, red, green, blue, 0, 0, width, height), amaze_demosaic_RT.c:113-1487.
 * Planes are row-major with `stride` >= width + 16 floats per row; raw's columns >= width must be 0
 * (hdr.c:971).  Cells of red/green/blue that the reference leaves untouched are left untouched. */
void orc_amaze_demosaic(const float *raw, float *red, float *green, float *blue, int stride, int width, int height, int fresh_tiles)
{
    planes_t P;
    planes_alloc(&P);
    for (int top = -16; top < height; top += TS - 32)
        for (int left = -16; left < width; left += TS - 32) {
            if (fresh_tiles == 1) memset(P.block, 0, P.bytes);
            else if (fresh_tiles == -1) {               /* poison everything but pmwt with NaN, pmwt with 0 */
                memset(P.block, 0xFF, P.bytes);
                memset(P.pmwt, 0, sizeof(float) * TS * TSH);
            }
            else if (fresh_tiles > 1) {                 /* debug: bit k+1 clears plane k only */
                float *pl[24] = {P.rgbgreen, P.delhvsqsum, P.dirwts0, P.dirwts1, P.vcd, P.hcd, P.vcdalt, P.hcdalt, P.cddiffsq, P.hvwt,
                                 P.Dgrb0, P.Dgrb1, P.delp, P.delm, P.rbint, P.Dgrb2, P.dgintv, P.dginth, P.Dgrbsq1m, P.Dgrbsq1p, P.cfa,
                                 P.pmwt, P.rbm, P.rbp};
                const int fullp[24] = {1,1,1,1,1,1,1,1,1,0, 0,0,0,0,0,1,1,1,0,0,1, 0,0,0};
                for (int k = 0; k < 24; k++)
                    if (fresh_tiles & (2 << k)) memset(pl[k], 0, sizeof(float) * TS * (fullp[k] ? TS : TSH));
            }
            amaze_tile(&P, raw, red, green, blue, stride, width, height, top, left);
        }
    free(P.block);
}
