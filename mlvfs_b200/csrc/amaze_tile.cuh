// amaze_tile.cuh -- AMaZE demosaic of one 160x160 tile as a cooperative thread-block program.
//
// Replaces reference amaze_demosaic_RT.c:292-1470 (the body of the tile loop) in its __SSE2__ form, which is
// what the reference compiles to on x86-64: 4-wide vector loops that run a few columns past the scalar
// bounds, and in-place passes whose lanes see a mix of already-updated and original neighbours.  Every
// value equals the reference's: IEEE binary32 in the reference's association order (build with
// -fmad=false), IEEE division, the two "exponent decrement" halvings and the three fp64 promotions.
//
// The reference walks the tiles one after another through one calloc'ed block; here every tile runs in its
// own thread block on its own workspace.  That is exact whenever the reference's result does not depend on
// what the previous tile left behind, which holds for every plane except `pmwt` (tests/test_oracle_vs_ref.py
// ::test_amaze_tiles_are_independent); pmwt is zeroed per tile, which equals the reference's state for
// widths that are multiples of 128 (all BASELINE configs).  For other widths the last partial tile column
// can differ from the sequential reference in a few border pixels (documented in DESIGN.md).
//
// The body is one template over a "context" (thread id, block size, barrier, shared scratch) so that the
// same source runs as a CUDA block and, for tests, as a group of host threads (tests/emu/amaze_emu.cpp).
// Sequential dependencies of the reference are kept exactly:
//   * variance pass (:748-803): column-per-thread walk down the rows, one barrier per row for the
//     two left-neighbour lanes that must see updated values;
//   * hvwt (:1054-1058) and pmwt (:1269-1272) refinements: row after row through a shared row buffer;
//   * Nyquist 3x3 vote (:998-1010): raster-sequential, run only for rows that can change.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define AMZ_HD __host__ __device__ __forceinline__
#else
#define AMZ_HD inline
#endif

namespace amaze {

constexpr int TS = 160, TSH = 80;                                      // amaze_demosaic_RT.c:137-138
constexpr int V1 = TS, V2 = 2 * TS, V3 = 3 * TS, P1 = -TS + 1, P2 = -2 * TS + 2, P3 = -3 * TS + 3,
              M1 = TS + 1, M2 = 2 * TS + 2, M3 = 3 * TS + 3;           // :147
constexpr size_t FULL_B = sizeof(float) * TS * TS, HALF_B = sizeof(float) * TS * TSH, GAP_B = 64;
// the reference's block (:244): 22 full planes' worth of floats + nyquist bytes + 64-byte gaps
constexpr size_t WS_BYTES = (22 * FULL_B + TS * TSH + 23 * GAP_B + 255) / 256 * 256;

struct Ws {                                                            // plane order of :249-273
    float *rgbgreen, *delhvsqsum, *dirwts0, *dirwts1, *vcd, *hcd, *vcdalt, *hcdalt, *cddiffsq, *hvwt, *Dgrb0, *Dgrb1,
          *delp, *delm, *rbint, *Dgrb2, *dgintv, *dginth, *Dgrbsq1m, *Dgrbsq1p, *cfa, *pmwt, *rbm, *rbp;
    unsigned char *nyquist;
};

AMZ_HD Ws carve(char *p)
{
    Ws W;
#define AMZ_TAKE(name, sz) W.name = (float *)p; p += (sz) + GAP_B
    AMZ_TAKE(rgbgreen, FULL_B); AMZ_TAKE(delhvsqsum, FULL_B); AMZ_TAKE(dirwts0, FULL_B); AMZ_TAKE(dirwts1, FULL_B);
    AMZ_TAKE(vcd, FULL_B); AMZ_TAKE(hcd, FULL_B); AMZ_TAKE(vcdalt, FULL_B); AMZ_TAKE(hcdalt, FULL_B); AMZ_TAKE(cddiffsq, FULL_B);
    AMZ_TAKE(hvwt, HALF_B);
    W.Dgrb0 = (float *)p; W.Dgrb1 = W.Dgrb0 + TS * TSH; p += FULL_B + GAP_B;
    AMZ_TAKE(delp, HALF_B); AMZ_TAKE(delm, HALF_B); AMZ_TAKE(rbint, HALF_B); AMZ_TAKE(Dgrb2, FULL_B); AMZ_TAKE(dgintv, FULL_B);
    AMZ_TAKE(dginth, FULL_B); AMZ_TAKE(Dgrbsq1m, HALF_B); AMZ_TAKE(Dgrbsq1p, HALF_B); AMZ_TAKE(cfa, FULL_B);
    AMZ_TAKE(pmwt, HALF_B); AMZ_TAKE(rbm, HALF_B); AMZ_TAKE(rbp, HALF_B);
#undef AMZ_TAKE
    W.nyquist = (unsigned char *)p;
    return W;
}

struct Geom { int width, height, top, left, rr1, cc1, rrmin, rrmax, ccmin, ccmax; };

AMZ_HD Geom tile_geom(int width, int height, int top, int left)        // :296-303, :373-376
{
    Geom G;
    G.width = width; G.height = height; G.top = top; G.left = left;
    const int bottom = top + TS < height + 16 ? top + TS : height + 16;
    const int right = left + TS < width + 16 ? left + TS : width + 16;
    G.rr1 = bottom - top; G.cc1 = right - left;
    G.rrmin = top < 0 ? 16 : 0; G.ccmin = left < 0 ? 16 : 0;
    G.rrmax = bottom > height ? height - top : G.rr1;
    G.ccmax = right > width ? width - left : G.cc1;
    return G;
}
AMZ_HD int tiles_along(int extent) { return (extent + 16 + (TS - 32) - 1) / (TS - 32); }   // top = -16, -16+128, ... < extent

struct Shared {                       // block-shared scratch of the sequential passes
    float row[2][2][TS];              // [pass: hvwt / pmwt refinement][row parity][half-row site]: the previous row, updated
    int rowcnt[TS];
    int scan[2][32];                  // Nyquist vote: per-lane transition functions (two buffers for the warp scan)
    int votecnt;
    int anynyq;
};

// ---- scalar helpers with the reference's exact comparison semantics ----
AMZ_HD int fc(int r, int c) { return ((r & 1) == 0 && (c & 1) == 0) ? 0 : ((r & 1) && (c & 1)) ? 2 : 1; }   // :41-49
AMZ_HD float expdec(float d, int n)                                    // xdiv2f / xdivf, :88-100
{
#if defined(__CUDA_ARCH__)
    int i = __float_as_int(d);
    if (i & 0x7FFFFFFF) i -= n << 23;
    return __int_as_float(i);
#else
    int32_t i; memcpy(&i, &d, 4);
    if (i & 0x7FFFFFFF) i -= n << 23;
    memcpy(&d, &i, 4);
    return d;
#endif
}
AMZ_HD float sq(float a) { return a * a; }
AMZ_HD float ab(float a) { return fabsf(a); }
AMZ_HD float vmin(float a, float b) { return a < b ? a : b; }          // minps / maxps operand order
AMZ_HD float vmax(float a, float b) { return a > b ? a : b; }
AMZ_HD float limv(float a, float b, float c) { return vmax(b, vmin(a, c)); }                 // sleefsseavx.c:1295
AMZ_HD float ulimv(float a, float b, float c) { return b < c ? limv(a, b, c) : limv(a, c, b); }
AMZ_HD float lims(float x, float lo, float hi) { const float m = x < hi ? x : hi; return m > lo ? m : lo; }   // scalar LIM
AMZ_HD float ulims(float a, float b, float c) { return b < c ? lims(a, b, c) : lims(a, c, b); }
AMZ_HD int cdiv(int a, int b) { return a > 0 ? (a + b - 1) / b : 0; }

#define AMZ_EPS 1e-5f
#define AMZ_EPSSQ 1e-10f
#define AMZ_ARTHRESH 0.75f
#define AMZ_CLIP 1.0f
#define AMZ_CLIP8 0.8f

// one cell of the variance/bounding pass (:766-800) given the (possibly updated) left neighbour hm2 and the
// row-above neighbour vm2; everything else original
AMZ_HD float bound_cd(float cd, float c0, float n1, float n2, float sgn)
{
    const float nsgn = -sgn, sgn3 = 3.0f * sgn;
    const float Gint = sgn * cd + c0;
    const float t2 = sgn3 * cd;
    const float wt = 1.0f + t2 / (AMZ_EPS + Gint + c0);
    const bool pos = nsgn * cd > 0.0f;
    const float old = cd;
    const float t = nsgn * (c0 - ulimv(Gint, n1, n2));
    float r = (t2 < -(c0 + Gint)) ? t : wt * cd + (1.0f - wt) * t;
    r = pos ? r : old;
    r = Gint > AMZ_CLIP ? t : r;
    return r;
}
AMZ_HD float var3(float a, float b, float c) { return 3.0f * (sq(a) + sq(b) + sq(c)) - sq(a + b + c); }

// (row, column) of the flat index idx = tid, tid + nthr, ... over rows of n cells: one division per loop instead of one
// per iteration (the passes below are loops of ~90 iterations per thread)
struct Strider {
    int row, col, q, r, n;
    AMZ_HD Strider(int tid, int nthr, int n_) : n(n_ > 0 ? n_ : 1)
    {
        row = tid / n; col = tid - row * n;
        q = nthr / n; r = nthr - q * n;
    }
    AMZ_HD void step()
    {
        col += r; row += q;
        if (col >= n) { col -= n; row++; }
    }
};

// ------------------------------------------------------------------------------------------------------
template <class Ctx>
AMZ_HD void tile_body(Ctx &C, const Ws &W, const Geom &G, Shared &S, const float *__restrict__ raw, float *__restrict__ red,
                      float *__restrict__ green, float *__restrict__ blue, int stride)
{
    const int tid = C.tid, nthr = C.nthr;
    const int rr1 = G.rr1, cc1 = G.cc1, top = G.top, left = G.left, width = G.width, height = G.height;
    float *const cfa = W.cfa, *const rgbgreen = W.rgbgreen, *const hvwt = W.hvwt, *const pmwt = W.pmwt;
    unsigned char *const nyquist = W.nyquist;
    const float gaussodd[4] = {0.14659727707323927f, 0.103592713382435f, 0.0732036125103057f, 0.0365543548389495f};
    const float gaussgrad[6] = {0.07384411893421103f, 0.06207511968171489f, 0.0521818194747806f,
                                0.03687419286733595f, 0.03099732204057846f, 0.018413194161458882f};
    const float gausseven[2] = {0.13719494435797422f, 0.05640252782101291f};
    const float gquinc[4] = {0.169917f, 0.108947f, 0.069855f, 0.0287182f};
    // Software prefetch (a hint: no effect on any value).  A strided pass over the tile advances nthr / row-width rows
    // per iteration; every thread asks for the cell AMZ_PF_ROWS rows below the lowest row it reads now, so the swaths
    // of successive iterations cover the planes ahead of the loads.  The work planes of the ~600 resident tile
    // programs do not fit the L2 (amaze.cu), so without this every pass waits on DRAM with a handful of loads in
    // flight per warp.
#ifndef AMZ_PF_ROWS
#define AMZ_PF_ROWS 6
#endif
    const int pfd = AMZ_PF_ROWS * TS;
    auto pf = [&](const float *pl, int i) { if (AMZ_PF_ROWS && i < TS * TS) C.prefetch(pl + i); };          // full plane
    auto pfh = [&](const float *pl, int i) { if (AMZ_PF_ROWS && i < TS * TS) C.prefetch(pl + (i >> 1)); };  // half plane, full index

    // ---- per-tile clears (:294-295) + pmwt (see header) ----
    for (int k = tid; k < TS * TSH; k += nthr) { pmwt[k] = 0.0f; W.rbint[k] = 0.0f; nyquist[k] = 0; }
    for (int k = tid; k < TS; k += nthr) S.rowcnt[k] = 0;
    if (tid == 0) S.anynyq = 0;
    C.sync();
    C.mark(0);

    // ---- load + mirrored borders (:378-469); the bottom border may run past row 159 like the reference's ----
    {
        const int rrend = G.rrmax < rr1 ? G.rrmax + 16 : rr1;
        Strider sd0(tid, nthr, cc1);
        for (int idx = tid; idx < rrend * cc1; idx += nthr, sd0.step()) {
            const int rr = sd0.row, cc = sd0.col;
            const bool tb = rr < G.rrmin, bb = rr >= G.rrmax, lb = cc < G.ccmin, rb = cc >= G.ccmax;
            const int rl = rr - G.rrmax, cl = cc - G.ccmax;              // local indices inside the bottom / right border
            int sr, sc, fr = rr, fcc = cc;
            bool vec;                                                    // vector copies fill rgbgreen at every site
            if (tb && lb)      { sr = 32 - rr;          sc = 32 - (cc & ~3) + (cc & 3);             vec = true; }            // :435-443
            else if (bb && rb) { sr = height - rl - 2;  sc = width - (cl & ~3) - 2 + (cl & 3);      vec = true; }            // :444-452
            else if (tb && rb) { sr = 32 - rr;          sc = width - cl - 2;                        vec = false; fcc = cl; } // :453-461
            else if (bb && lb) { sr = height - rl - 2;  sc = 32 - cc;                               vec = false; fr = rl; }  // :462-469
            else if (tb)       { sr = 32 - rr + top;    sc = cc + left;                             vec = false; }           // :399-406
            else if (bb)       { sr = height - rl - 2;  sc = cc + left;                             vec = true; }            // :407-415
            else if (lb)       { sr = rr + top;         sc = 32 - cc + left;                        vec = false; }           // :417-424
            else if (rb)       { sr = rr + top;         sc = width - cl - 2;                        vec = false; fcc = cl; } // :426-433
            else               { sr = rr + top;         sc = cc + left;                             vec = true; }            // :381-387
            if (AMZ_PF_ROWS && sr + AMZ_PF_ROWS < height) C.prefetch(raw + (size_t)(sr + AMZ_PF_ROWS) * stride + sc);
            const float v = C.ld_stream(raw + (size_t)sr * stride + sc) / 65535.0f;   // read 1.56 times per frame, never again
            cfa[rr * TS + cc] = v;
            if (vec || fc(fr, fcc) == 1) rgbgreen[rr * TS + cc] = v;
        }
    }
    C.sync();
    C.mark(1);

    // ---- gradients, directional weights (:553-567) and diagonal gradients (:581-606) ----
    {
        const int cw = cdiv(cc1, 4) * 4, nrow = rr1 - 4;
        Strider sd1(tid, nthr, cw);
        for (int idx = tid; idx < nrow * cw; idx += nthr, sd1.step()) {
            const int rr = 2 + sd1.row, cc = sd1.col, i = rr * TS + cc;
            pf(cfa, i + V2 + pfd);
            const float delh = ab(cfa[i + 1] - cfa[i - 1]), delv = ab(cfa[i + V1] - cfa[i - V1]);
            W.dirwts1[i] = AMZ_EPS + ab(cfa[i + 2] - cfa[i]) + ab(cfa[i] - cfa[i - 2]) + delh;
            W.dirwts0[i] = AMZ_EPS + ab(cfa[i + V2] - cfa[i]) + ab(cfa[i] - cfa[i - V2]) + delv;
            W.delhvsqsum[i] = delh * delh + delv * delv;
        }
        const int np = 4 * cdiv(cc1 - 12, 8), nrow6 = rr1 - 12;
        Strider sd2(tid, nthr, np);
        for (int idx = tid; idx < nrow6 * np; idx += nthr, sd2.step()) {
            const int rr = 6 + sd2.row, cc = 6 + 2 * sd2.col, i = rr * TS + cc;
            const int o = (fc(rr, 2) & 1) ? 0 : 1, g = i + o, c = i + (1 - o);
            const float t = cfa[g];
            W.delp[i >> 1] = ab(cfa[c + P1] - cfa[c - P1]);
            W.delm[i >> 1] = ab(cfa[c + M1] - cfa[c - M1]);
            W.Dgrbsq1m[i >> 1] = sq(t - cfa[g - M1]) + sq(t - cfa[g + M1]);
            W.Dgrbsq1p[i >> 1] = sq(t - cfa[g - P1]) + sq(t - cfa[g + P1]);
        }
    }
    C.sync();
    C.mark(2);

    // ---- H/V colour differences (:633-689) and the diagonal R/B estimates (:1115-1180) ----
    {
        const int nc = 4 * cdiv(cc1 - 11, 4), nrow = rr1 - 8;
        Strider sd3(tid, nthr, nc);
        for (int idx = tid; idx < nrow * nc; idx += nthr, sd3.step()) {
            const int rr = 4 + sd3.row, cc = 4 + sd3.col, i = rr * TS + cc;
            const float sgn = ((rr + cc) & 1) ? -1.0f : 1.0f;
            const float *d0 = W.dirwts0, *d1 = W.dirwts1;
            pf(cfa, i + V2 + pfd); pf(d0, i + V2 + pfd); pf(d1, i + pfd);
            const float c0 = cfa[i];
            const float cru = cfa[i - V1] * (d0[i - V2] + d0[i]) / (d0[i - V2] * (AMZ_EPS + c0) + d0[i] * (AMZ_EPS + cfa[i - V2]));
            const float crd = cfa[i + V1] * (d0[i + V2] + d0[i]) / (d0[i + V2] * (AMZ_EPS + c0) + d0[i] * (AMZ_EPS + cfa[i + V2]));
            const float crl = cfa[i - 1] * (d1[i - 2] + d1[i]) / (d1[i - 2] * (AMZ_EPS + c0) + d1[i] * (AMZ_EPS + cfa[i - 2]));
            const float crr = cfa[i + 1] * (d1[i + 2] + d1[i]) / (d1[i + 2] * (AMZ_EPS + c0) + d1[i] * (AMZ_EPS + cfa[i + 2]));
            const float guha = cfa[i - V1] + 0.5f * (c0 - cfa[i - V2]), gdha = cfa[i + V1] + 0.5f * (c0 - cfa[i + V2]);
            const float glha = cfa[i - 1] + 0.5f * (c0 - cfa[i - 2]), grha = cfa[i + 1] + 0.5f * (c0 - cfa[i + 2]);
            float guar = ab(1.0f - cru) < AMZ_ARTHRESH ? c0 * cru : guha;
            float gdar = ab(1.0f - crd) < AMZ_ARTHRESH ? c0 * crd : gdha;
            float glar = ab(1.0f - crl) < AMZ_ARTHRESH ? c0 * crl : glha;
            float grar = ab(1.0f - crr) < AMZ_ARTHRESH ? c0 * crr : grha;
            const float hwt = d1[i - 1] / (d1[i - 1] + d1[i + 1]);
            const float vwt = d0[i - V1] / (d0[i + V1] + d0[i - V1]);
            const float Ginthha = hwt * grha + (1.0f - hwt) * glha, Gintvha = vwt * gdha + (1.0f - vwt) * guha;
            const float ha = sgn * (Ginthha - c0), va = sgn * (Gintvha - c0);
            W.hcdalt[i] = ha; W.vcdalt[i] = va;
            const bool clip = (c0 > AMZ_CLIP8) | (Gintvha > AMZ_CLIP8) | (Ginthha > AMZ_CLIP8);
            if (clip) { guar = guha; gdar = gdha; glar = glha; grar = grha; }
            W.vcd[i] = clip ? va : sgn * ((vwt * gdar + (1.0f - vwt) * guar) - c0);
            W.hcd[i] = clip ? ha : sgn * ((hwt * grar + (1.0f - hwt) * glar) - c0);
            W.dgintv[i] = vmin(sq(guha - gdha), sq(guar - gdar));
            W.dginth[i] = vmin(sq(glha - grha), sq(glar - grar));
        }
        const int nrow8 = rr1 - 16;
        for (int par = 0; par < 2; par++) {                               // rows of one parity share a site count
            const int ns = 4 * cdiv(cc1 - 16 - par, 8), nr = (nrow8 + 1 - par) / 2;      // rows 8+par, 10+par, ...
            Strider sd4(tid, nthr, ns);
            for (int idx = tid; idx < nr * ns; idx += nthr, sd4.step()) {
                const int rr = 8 + par + 2 * (sd4.row), cc = 8 + par + 2 * sd4.col, i = rr * TS + cc, i1 = i >> 1;
                pf(cfa, i + M2 + 2 * pfd); pfh(W.delm, i + M2 + 2 * pfd); pfh(W.delp, i + M2 + 2 * pfd);
                pfh(W.Dgrbsq1m, i + V2 + 2 * pfd); pfh(W.Dgrbsq1p, i + V2 + 2 * pfd);
                const float c0 = cfa[i];
                float t1, t2, w;
                t1 = cfa[i + M1]; t2 = cfa[i + M2];
                float rbse = (t1 + t1) / (AMZ_EPS + c0 + t2);
                rbse = ab(1.0f - rbse) < AMZ_ARTHRESH ? c0 * rbse : t1 + 0.5f * (c0 - t2);
                t1 = cfa[i - M1]; t2 = cfa[i - M2];
                float rbnw = (t1 + t1) / (AMZ_EPS + c0 + t2);
                rbnw = ab(1.0f - rbnw) < AMZ_ARTHRESH ? c0 * rbnw : t1 + 0.5f * (c0 - t2);
                t1 = AMZ_EPS + W.delm[i1];
                const float wtse = t1 + W.delm[(i + M1) >> 1] + W.delm[(i + M2) >> 1], wtnw = t1 + W.delm[(i - M1) >> 1] + W.delm[(i - M2) >> 1];
                const float m = (wtse * rbnw + wtnw * rbse) / (wtse + wtnw);
                t1 = ulimv(m, cfa[i - M1], cfa[i + M1]);
                w = 2.0f * (c0 - m) / (AMZ_EPS + m + c0);
                t2 = w * m + (1.0f - w) * t1;
                t2 = (m + m < c0) ? t1 : t2;
                t2 = (m < c0) ? t2 : m;
                W.rbm[i1] = t2 > AMZ_CLIP ? ulimv(t2, cfa[i - M1], cfa[i + M1]) : t2;

                t1 = cfa[i + P1]; t2 = cfa[i + P2];
                float rbne = (t1 + t1) / (AMZ_EPS + c0 + t2);
                rbne = ab(1.0f - rbne) < AMZ_ARTHRESH ? c0 * rbne : t1 + 0.5f * (c0 - t2);
                t1 = cfa[i - P1]; t2 = cfa[i - P2];
                float rbsw = (t1 + t1) / (AMZ_EPS + c0 + t2);
                rbsw = ab(1.0f - rbsw) < AMZ_ARTHRESH ? c0 * rbsw : t1 + 0.5f * (c0 - t2);
                t1 = AMZ_EPS + W.delp[i1];
                const float wtne = t1 + W.delp[(i + P1) >> 1] + W.delp[(i + P2) >> 1], wtsw = t1 + W.delp[(i - P1) >> 1] + W.delp[(i - P2) >> 1];
                const float p = (wtne * rbsw + wtsw * rbne) / (wtne + wtsw);
                t1 = ulimv(p, cfa[i - P1], cfa[i + P1]);
                w = 2.0f * (c0 - p) / (AMZ_EPS + p + c0);
                t2 = w * p + (1.0f - w) * t1;
                t2 = (p + p < c0) ? t1 : t2;
                t2 = (p < c0) ? t2 : p;
                W.rbp[i1] = t2 > AMZ_CLIP ? ulimv(t2, cfa[i - P1], cfa[i + P1]) : t2;
#define AMZ_EVEN8(A) (gausseven[0] * (A[(i - V1) >> 1] + A[(i - 1) >> 1] + A[(i + 1) >> 1] + A[(i + V1) >> 1]) +                 \
                      gausseven[1] * (A[(i - V2 - 1) >> 1] + A[(i - V2 + 1) >> 1] + A[(i - 2 - V1) >> 1] + A[(i + 2 - V1) >> 1] +  \
                                      A[(i - 2 + V1) >> 1] + A[(i + 2 + V1) >> 1] + A[(i + V2 - 1) >> 1] + A[(i + V2 + 1) >> 1]))
                const float rbvarm = AMZ_EPSSQ + AMZ_EVEN8(W.Dgrbsq1m);
                pmwt[i1] = rbvarm / ((AMZ_EPSSQ + AMZ_EVEN8(W.Dgrbsq1p)) + rbvarm);
#undef AMZ_EVEN8
            }
        }
    }
    C.sync();
    C.mark(3);

    // ---- variance-based choice + bounding (:748-803) ----
    // The reference runs this in place, four columns at a time: lanes 0 and 1 of a vector see the already updated
    // hcd two columns to their left (lanes 2 and 3 of the previous vector, which themselves read originals only),
    // and vcd sees the updated row two above.  Nothing else it reads has been rewritten.  So:
    //   h: every cell from ORIGINAL values only -- a lane 0 / 1 cell first re-derives its left neighbour's updated
    //      value (the same expression on that column's originals).  All cells in parallel, reading a copy of hcd
    //      taken before anything is overwritten; no row-by-row barriers.
    //   v: a recurrence down each column (row rr needs the updated row rr - 2 of its own column only): one thread per
    //      column walks the rows, the updated values of the two parities in registers; threads never read each
    //      other's cells, so no barriers either.
    {
        const int ncol = 4 * cdiv(cc1 - 8, 4);                            // columns 4 .. 4+ncol-1
        const int nrow = rr1 - 8;                                         // rows 4 .. rr1-5
        float *const horig = W.Dgrb2;                                     // free until the G pass writes it
        for (int idx = tid; idx < (nrow > 0 ? nrow : 0) * TS; idx += nthr) { pf(W.hcd, 4 * TS + idx + pfd); horig[4 * TS + idx] = W.hcd[4 * TS + idx]; }
        C.sync();
        C.mark(4);
        auto h_from_originals = [&](int i, float sgn, float hm2) {        // hm2: original or updated left neighbour
            const float h0 = horig[i], ha = W.hcdalt[i];
            const float havar = var3(W.hcdalt[i - 2], ha, W.hcdalt[i + 2]);
            const float h = (havar < var3(hm2, h0, horig[i + 2])) ? ha : h0;
            return bound_cd(h, cfa[i], cfa[i - 1], cfa[i + 1], sgn);
        };
        Strider sd5(tid, nthr, ncol);
        for (int idx = tid; idx < (nrow > 0 ? nrow : 0) * ncol; idx += nthr, sd5.step()) {
            const int rr = 4 + sd5.row, t = sd5.col, cc = 4 + t, i = rr * TS + cc;
            const float sgn = ((rr + cc) & 1) ? -1.0f : 1.0f;             // the same for column cc - 2
            pf(horig, i + pfd); pf(W.hcdalt, i + pfd); pf(cfa, i + pfd);
            float hm2 = horig[i - 2];
            if ((t & 3) < 2 && t >= 2) hm2 = h_from_originals(i - 2, sgn, horig[i - 4]);
            W.hcd[i] = h_from_originals(i, sgn, hm2);
        }
        C.sync();
        C.mark(5);
        if (tid < ncol) {
            const int cc = 4 + tid;
            float vup[2] = {0.0f, 0.0f};                                  // updated vcd of row rr-2 (same parity)
            struct VIn { float c0, cu, cd, v0, vm2, vp2, va, vam2, vap2, h; };
            auto load_v = [&](int rr) {
                VIn Q;
                const int i = rr * TS + cc;
                Q.c0 = cfa[i]; Q.cu = cfa[i - V1]; Q.cd = cfa[i + V1];
                Q.v0 = W.vcd[i]; Q.vm2 = W.vcd[i - V2]; Q.vp2 = W.vcd[i + V2];   // vm2 only matters for rows 4, 5 (rows 2, 3 are never updated)
                Q.va = W.vcdalt[i]; Q.vam2 = W.vcdalt[i - V2]; Q.vap2 = W.vcdalt[i + V2];
                Q.h = W.hcd[i];
                return Q;
            };
            VIn cur = {}, nx1 = {};
            if (4 < rr1 - 4) cur = load_v(4);
            if (5 < rr1 - 4) nx1 = load_v(5);
            for (int rr = 4; rr < rr1 - 4; rr++) {
                VIn nx2 = {};
                if (rr + 2 < rr1 - 4) nx2 = load_v(rr + 2);               // two rows ahead: rows rr+1, rr+2 are still original
                { const int ip = (rr + 4) * TS + cc + pfd; pf(cfa, ip); pf(W.vcd, ip); pf(W.vcdalt, ip); pf(W.hcd, ip - V2); }
                const int i = rr * TS + cc;
                const float sgn = ((rr + cc) & 1) ? -1.0f : 1.0f;
                const float vm2 = rr >= 6 ? vup[rr & 1] : cur.vm2;
                float v = (var3(cur.vam2, cur.va, cur.vap2) < var3(vm2, cur.v0, cur.vp2)) ? cur.va : cur.v0;
                v = bound_cd(v, cur.c0, cur.cu, cur.cd, sgn);
                vup[rr & 1] = v;
                W.vcd[i] = v;
                W.cddiffsq[i] = sq(v - cur.h);
                cur = nx1; nx1 = nx2;
            }
        }
    }
    C.sync();
    C.mark(6);

    // ---- H/V weight (:876-920) and Nyquist texture test (:967-996) ----
    for (int par = 0; par < 2; par++) {
        const int ns = 4 * cdiv(cc1 - 12 - par, 8), nr = (rr1 - 12 + 1 - par) / 2;
        Strider sd6(tid, nthr, ns);
        for (int idx = tid; idx < nr * ns; idx += nthr, sd6.step()) {
            const int rr = 6 + par + 2 * (sd6.row), cc = 6 + par + 2 * sd6.col, i = rr * TS + cc;
            const float *vcd = W.vcd, *hcd = W.hcd, *d0 = W.dirwts0, *d1 = W.dirwts1;
            pf(vcd, i + V3 + 2 * pfd); pf(hcd, i + 2 * pfd); pf(d0, i + V1 + 2 * pfd); pf(d1, i + 2 * pfd);
            pf(W.dgintv, i + V2 + 2 * pfd); pf(W.dginth, i + 2 * pfd);
            float t = vcd[i];
            const float uave = t + vcd[i - V1] + vcd[i - V2] + vcd[i - V3], dave = t + vcd[i + V1] + vcd[i + V2] + vcd[i + V3];
            float Du = sq(t - uave) + sq(vcd[i - V1] - uave) + sq(vcd[i - V2] - uave) + sq(vcd[i - V3] - uave);
            float Dd = sq(t - dave) + sq(vcd[i + V1] - dave) + sq(vcd[i + V2] - dave) + sq(vcd[i + V3] - dave);
            const float hwt = d1[i - 1] / (d1[i - 1] + d1[i + 1]);
            const float vwt = d0[i - V1] / (d0[i + V1] + d0[i - V1]);
            t = hcd[i];
            const float lave = t + hcd[i - 1] + hcd[i - 2] + hcd[i - 3], rave = t + hcd[i + 1] + hcd[i + 2] + hcd[i + 3];
            float Dl = sq(t - lave) + sq(hcd[i - 1] - lave) + sq(hcd[i - 2] - lave) + sq(hcd[i - 3] - lave);
            float Dr = sq(t - rave) + sq(hcd[i + 1] - rave) + sq(hcd[i + 2] - rave) + sq(hcd[i + 3] - rave);
            const float vcdvar = AMZ_EPSSQ + vwt * Dd + (1.0f - vwt) * Du, hcdvar = AMZ_EPSSQ + hwt * Dr + (1.0f - hwt) * Dl;
            Du = W.dgintv[i] + W.dgintv[i - V1] + W.dgintv[i - V2];
            Dd = W.dgintv[i] + W.dgintv[i + V1] + W.dgintv[i + V2];
            Dl = W.dginth[i] + W.dginth[i - 1] + W.dginth[i - 2];
            Dr = W.dginth[i] + W.dginth[i + 1] + W.dginth[i + 2];
            const float vcdvar1 = AMZ_EPSSQ + vwt * Dd + (1.0f - vwt) * Du, hcdvar1 = AMZ_EPSSQ + hwt * Dr + (1.0f - hwt) * Dl;
            const float varwt = hcdvar / (vcdvar + hcdvar), diffwt = hcdvar1 / (vcdvar1 + hcdvar1);
            const bool dec = ((0.5f - varwt) * (0.5f - diffwt) > 0.0f) & (ab(0.5f - diffwt) < ab(0.5f - varwt));
            hvwt[i >> 1] = dec ? varwt : diffwt;
        }
        const int nq = cdiv(cc1 - 12 - par, 2);
        Strider sd7(tid, nthr, nq);
        for (int idx = tid; idx < nr * nq; idx += nthr, sd7.step()) {
            const int rr = 6 + par + 2 * (sd7.row), cc = 6 + par + 2 * sd7.col, i = rr * TS + cc;
            const float *q = W.cddiffsq, *d = W.delhvsqsum;
            pf(q, i + V2 + 2 * pfd); pf(d, i + V2 + 2 * pfd);
            float nyqtest = (gaussodd[0] * q[i] + gaussodd[1] * (q[i - M1] + q[i + P1] + q[i - P1] + q[i + M1]) +
                             gaussodd[2] * (q[i - V2] + q[i - 2] + q[i + 2] + q[i + V2]) +
                             gaussodd[3] * (q[i - M2] + q[i + P2] + q[i - P2] + q[i + M2]));
            nyqtest -= 0.5f * (gaussgrad[0] * d[i] + gaussgrad[1] * (d[i - V1] + d[i + 1] + d[i - 1] + d[i + V1]) +
                               gaussgrad[2] * (d[i - M1] + d[i + P1] + d[i - P1] + d[i + M1]) +
                               gaussgrad[3] * (d[i - V2] + d[i - 2] + d[i + 2] + d[i + V2]) +
                               gaussgrad[4] * (d[i - 2 * TS - 1] + d[i - 2 * TS + 1] + d[i - TS - 2] + d[i - TS + 2] +
                                               d[i + TS - 2] + d[i + TS + 2] + d[i + 2 * TS - 1] + d[i + 2 * TS + 1]) +
                               gaussgrad[5] * (d[i - M2] + d[i + P2] + d[i - P2] + d[i + M2]));
            if (nyqtest > 0) { nyquist[i >> 1] = 1; C.atomic_add(&S.rowcnt[rr], 1); C.atomic_add(&S.anynyq, 1); }
        }
    }
    C.sync();
    C.mark(7);
    const bool anynyq = S.anynyq != 0;                                    // block-uniform

    if (anynyq) {
        // ---- 3x3 vote, raster-sequential and in place (:998-1010) ----
        // A cell's new value depends on the new value of the cell two columns to its left: x' = vote(sum of the seven
        // other neighbours + own old value + x'_left).  For a fixed row that is a two-state automaton, so the row is
        // a scan over function composition: warp 0 alone walks the rows (rows still run in order -- each reads the
        // voted row above), three neighbouring cells per lane, the lanes' composed transition functions combined by
        // a warp scan through shared memory.  The other warps wait at the barrier below.
        if (tid < 32) {
            const int lane = tid;
            for (int rr = 8; rr < rr1 - 8; rr++) {
                if (S.rowcnt[rr - 2] + S.rowcnt[rr - 1] + S.rowcnt[rr] + S.rowcnt[rr + 1] + S.rowcnt[rr + 2] == 0) continue;   // stays all zero
                const int par = fc(rr, 2) & 1, ns = cdiv(cc1 - 16 - par, 2);
                const int per = cdiv(ns, 32);                             // cells per lane (3 for a full tile), lane's cells are adjacent
                int f0[4], f1[4], oldv[4];                                // per cell: next state for x'_left = 0 / 1
                int F0 = 0, F1 = 1;                                       // the lane's composed function
                for (int j = 0; j < per && j < 4; j++) {
                    const int k = lane * per + j;
                    f0[j] = 0; f1[j] = 1; oldv[j] = 0;
                    if (k < ns) {
                        const int i = rr * TS + 8 + par + 2 * k;
                        const int oth = nyquist[(i - V2) >> 1] + nyquist[(i - M1) >> 1] + nyquist[(i + P1) >> 1] + nyquist[(i + 2) >> 1] +
                                        nyquist[(i - P1) >> 1] + nyquist[(i + M1) >> 1] + nyquist[(i + V2) >> 1];
                        const int old = nyquist[i >> 1], n = oth + old;
                        oldv[j] = old;
                        f0[j] = n > 4 ? 1 : (n < 4 ? 0 : old);
                        f1[j] = n + 1 > 4 ? 1 : (n + 1 < 4 ? 0 : old);
                    }
                    F0 = F0 ? f1[j] : f0[j];
                    F1 = F1 ? f1[j] : f0[j];
                }
                // inclusive scan of "apply the lanes to the left first": two bits per lane, double-buffered
                int buf = 0;
                S.scan[0][lane] = F0 | (F1 << 1);
                C.syncwarp();
                for (int d = 1; d < 32; d <<= 1) {
                    int mine = S.scan[buf][lane];
                    if (lane >= d) {
                        const int left = S.scan[buf][lane - d];           // covers the lanes before mine's range
                        const int m0 = (left & 1) ? (mine >> 1) & 1 : mine & 1, m1 = (left & 2) ? (mine >> 1) & 1 : mine & 1;
                        mine = m0 | (m1 << 1);
                    }
                    S.scan[buf ^ 1][lane] = mine;
                    buf ^= 1;
                    C.syncwarp();
                }
                int x = nyquist[(rr * TS + 8 + par - 2) >> 1];             // the cell left of the voted range keeps its value
                if (lane > 0) { const int pre = S.scan[buf][lane - 1]; x = x ? (pre >> 1) & 1 : pre & 1; }
                int cnt = 0;
                for (int j = 0; j < per && j < 4; j++) {
                    const int k = lane * per + j;
                    if (k < ns) {
                        x = x ? f1[j] : f0[j];
                        nyquist[(rr * TS + 8 + par + 2 * k) >> 1] = (unsigned char)x;
                        cnt += x;
                    }
                }
                if (lane == 0) {
                    // cells of this row outside the voted range keep their test result
                    const int i0 = rr * TS;
                    for (int c2 = 6 + par; c2 < 8 + par; c2 += 2) cnt += nyquist[(i0 + c2) >> 1];
                    for (int c2 = 8 + par + 2 * ns; c2 < cc1 - 6; c2 += 2) cnt += nyquist[(i0 + c2) >> 1];
                    S.votecnt = 0;
                }
                C.syncwarp();
                if (cnt) C.atomic_add(&S.votecnt, cnt);
                C.syncwarp();
                if (lane == 0) S.rowcnt[rr] = S.votecnt;
                C.syncwarp();
            }
        }
        C.sync();
        C.mark(8);

        // ---- area interpolation inside Nyquist regions (:1016-1045) ----
        for (int par = 0; par < 2; par++) {
            const int ns = cdiv(cc1 - 16 - par, 2), nr = (rr1 - 16 + 1 - par) / 2;
            Strider sd8(tid, nthr, ns);
            for (int idx = tid; idx < nr * ns; idx += nthr, sd8.step()) {
                const int rr = 8 + par + 2 * (sd8.row), cc = 8 + par + 2 * sd8.col, i = rr * TS + cc;
                pf(cfa, i + 7 * TS + 2 * pfd);
                if (!nyquist[i >> 1]) continue;
                float sumh = 0, sumv = 0, sumsqh = 0, sumsqv = 0, areawt = 0;
                for (int a = -6; a < 7; a += 2)
                    for (int b = -6; b < 7; b += 2) {
                        const int j = (rr + a) * TS + cc + b;
                        if (!nyquist[j >> 1]) continue;
                        sumh += cfa[j] - expdec(cfa[j - 1] + cfa[j + 1], 1);
                        sumv += cfa[j] - expdec(cfa[j - V1] + cfa[j + V1], 1);
                        sumsqh += expdec(sq(cfa[j] - cfa[j - 1]) + sq(cfa[j] - cfa[j + 1]), 1);
                        sumsqv += expdec(sq(cfa[j] - cfa[j - V1]) + sq(cfa[j] - cfa[j + V1]), 1);
                        areawt += 1;
                    }
                const float hv = AMZ_EPSSQ + ab(areawt * sumsqh - sumh * sumh), vv = AMZ_EPSSQ + ab(areawt * sumsqv - sumv * sumv);
                hvwt[i >> 1] = hv / (vv + hv);
            }
        }
        C.sync();
        C.mark(9);
    }

    // ---- hvwt refinement from the diagonal neighbours, row after row (:1054-1058), and
    //      pmwt refinement + R+B estimate the same way (:1264-1274) ----
    // Each is a recurrence over the rows (a row reads the UPDATED row above and the ORIGINAL row below) with no
    // dependency inside a row, and the two do not touch each other's planes: warp 0 walks hvwt while warp 1 walks
    // pmwt, each with its own shared row buffer and warp-level synchronisation only (a block barrier per row would
    // make all eight warps wait ~290 times per tile).  The originals of the next row are loaded before the current
    // row's barrier.
    if (tid < 64) {
        const int pass = tid >> 5, lane = tid & 31;
        float *const P = pass ? pmwt : hvwt;
        const int r0 = pass ? 10 : 8;
        float (*rowbuf)[TS] = S.row[pass];
        for (int k = lane; k < TSH; k += 32) rowbuf[(r0 - 1) & 1][k] = P[(r0 - 1) * TSH + k];
        C.syncwarp();
        struct RefIn { float t, dl, dr, c, m, p; bool in; };
        auto load_ref = [&](int rr, int k) {
            RefIn Q = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, false};
            if (k >= TSH) return Q;
            const int par = fc(rr, 2) & 1;
            // processed sites: scalar bound for hvwt, whole 4-site vectors for pmwt
            const int ns = pass ? 4 * cdiv(cc1 - 20 - par, 8) : cdiv(cc1 - 16 - par, 2);
            const int k0 = (r0 + par) >> 1;                                // half index of the first processed site
            Q.t = P[rr * TSH + k];
            Q.in = k >= k0 && k < k0 + ns;
            if (Q.in) {
                Q.dl = P[(rr + 1) * TSH + k - 1 + par]; Q.dr = P[(rr + 1) * TSH + k + par];
                if (pass) { Q.c = cfa[rr * TS + 2 * k + par]; Q.m = W.rbm[rr * TSH + k]; Q.p = W.rbp[rr * TSH + k]; }
            }
            return Q;
        };
        constexpr int NS = (TSH + 31) / 32;                                // sites per lane: k = lane, lane + 32, lane + 64
        RefIn cur_in[NS] = {}, nxt_in[NS] = {};
        if (r0 < rr1 - r0)
            for (int j = 0; j < NS; j++) cur_in[j] = load_ref(r0, lane + 32 * j);
        for (int rr = r0; rr < rr1 - r0; rr++) {
            const int par = fc(rr, 2) & 1;
            const float *prev = rowbuf[(rr - 1) & 1];
            float *cur = rowbuf[rr & 1];
            if (rr + 1 < rr1 - r0)
                for (int j = 0; j < NS; j++) nxt_in[j] = load_ref(rr + 1, lane + 32 * j);
            if (AMZ_PF_ROWS && rr + 2 + AMZ_PF_ROWS < TS) {                 // a half-plane row is 2.5 lines, a cfa row 5
                const int rp = rr + 2 + AMZ_PF_ROWS;
                if (lane < 3) C.prefetch(P + rp * TSH + 32 * lane);
                else if (pass && lane < 6) C.prefetch(W.rbm + rp * TSH + 32 * (lane - 3));
                else if (pass && lane < 9) C.prefetch(W.rbp + rp * TSH + 32 * (lane - 6));
                else if (pass && lane < 14) C.prefetch(cfa + rp * TS + 32 * (lane - 9));
            }
            for (int j = 0; j < NS; j++) {
                const int k = lane + 32 * j;
                if (k >= TSH) continue;
                float t = cur_in[j].t;
                if (cur_in[j].in) {
                    // diagonal neighbours: previous row (updated) at half indices k-1+par, k+par; next row (original)
                    const float ul = prev[k - 1 + par], ur = prev[k + par];
                    const float s4 = ul + ur + cur_in[j].dl + cur_in[j].dr;
                    const float alt = pass ? 0.25f * s4 : expdec(s4, 2);
                    t = ab(0.5f - t) < ab(0.5f - alt) ? alt : t;
                    P[rr * TSH + k] = t;
                    if (pass) W.rbint[rr * TSH + k] = 0.5f * (cur_in[j].c + cur_in[j].m * (1.0f - t) + cur_in[j].p * t);
                }
                cur[k] = t;
            }
            C.syncwarp();
            for (int j = 0; j < NS; j++) cur_in[j] = nxt_in[j];
        }
    }
    C.sync();
    C.mark(10);
    {

        // ---- G at R/B sites with the final hvwt (:1063-1074) ----
        for (int par = 0; par < 2; par++) {
            const int ns = cdiv(cc1 - 16 - par, 2), nr = (rr1 - 16 + 1 - par) / 2;
            Strider sd9(tid, nthr, ns);
            for (int idx = tid; idx < nr * ns; idx += nthr, sd9.step()) {
                const int rr = 8 + par + 2 * (sd9.row), cc = 8 + par + 2 * sd9.col, i = rr * TS + cc, i1 = i >> 1;
                pf(W.hcd, i + 2 * pfd); pf(W.vcd, i + 2 * pfd); pfh(hvwt, i + 2 * pfd); pf(cfa, i + 2 * pfd); pf(rgbgreen, i + V1 + 2 * pfd);
                const float d = W.hcd[i] * (1.0f - hvwt[i1]) + W.vcd[i] * hvwt[i1];
                W.Dgrb0[i1] = d;
                const float g = cfa[i] + d;
                rgbgreen[i] = g;
                if (anynyq && nyquist[i1]) {
                    W.Dgrb2[2 * i1] = sq(g - expdec(rgbgreen[i - 1] + rgbgreen[i + 1], 1));
                    W.Dgrb2[2 * i1 + 1] = sq(g - expdec(rgbgreen[i - V1] + rgbgreen[i + V1], 1));
                } else
                    W.Dgrb2[2 * i1] = W.Dgrb2[2 * i1 + 1] = 0.0f;
            }
        }
        C.sync();
        C.mark(11);
        // ---- refine Nyquist sites with the local G curvature (:1085-1102) ----
        if (anynyq) {
#define AMZ_D2H(k) W.Dgrb2[2 * ((k) >> 1)]
#define AMZ_D2V(k) W.Dgrb2[2 * ((k) >> 1) + 1]
            for (int par = 0; par < 2; par++) {
                const int ns = cdiv(cc1 - 16 - par, 2), nr = (rr1 - 16 + 1 - par) / 2;
                Strider sd10(tid, nthr, ns);
                for (int idx = tid; idx < nr * ns; idx += nthr, sd10.step()) {
                    const int rr = 8 + par + 2 * (sd10.row), cc = 8 + par + 2 * sd10.col, i = rr * TS + cc;
                    pf(W.Dgrb2, i + V2 + 2 * pfd);
                    if (!nyquist[i >> 1]) continue;
                    const float gvarh = AMZ_EPSSQ + (gquinc[0] * AMZ_D2H(i) + gquinc[1] * (AMZ_D2H(i - M1) + AMZ_D2H(i + P1) + AMZ_D2H(i - P1) + AMZ_D2H(i + M1)) +
                                                     gquinc[2] * (AMZ_D2H(i - V2) + AMZ_D2H(i - 2) + AMZ_D2H(i + 2) + AMZ_D2H(i + V2)) +
                                                     gquinc[3] * (AMZ_D2H(i - M2) + AMZ_D2H(i + P2) + AMZ_D2H(i - P2) + AMZ_D2H(i + M2)));
                    const float gvarv = AMZ_EPSSQ + (gquinc[0] * AMZ_D2V(i) + gquinc[1] * (AMZ_D2V(i - M1) + AMZ_D2V(i + P1) + AMZ_D2V(i - P1) + AMZ_D2V(i + M1)) +
                                                     gquinc[2] * (AMZ_D2V(i - V2) + AMZ_D2V(i - 2) + AMZ_D2V(i + 2) + AMZ_D2V(i + V2)) +
                                                     gquinc[3] * (AMZ_D2V(i - M2) + AMZ_D2V(i + P2) + AMZ_D2V(i - P2) + AMZ_D2V(i + M2)));
                    const float d = (W.hcd[i] * gvarv + W.vcd[i] * gvarh) / (gvarv + gvarh);
                    W.Dgrb0[i >> 1] = d;
                    rgbgreen[i] = cfa[i] + d;
                }
            }
#undef AMZ_D2H
#undef AMZ_D2V
            C.sync();
            C.mark(12);
        }
    }

    // ---- where the diagonal estimate discriminates better, redo G from R+B (:1287-1352) ----
    for (int par = 0; par < 2; par++) {
        const int ns = cdiv(cc1 - 24 - par, 2), nr = (rr1 - 24 + 1 - par) / 2;
        Strider sd11(tid, nthr, ns);
        for (int idx = tid; idx < nr * ns; idx += nthr, sd11.step()) {
            const int rr = 12 + par + 2 * (sd11.row), cc = 12 + par + 2 * sd11.col, i = rr * TS + cc, i1 = i >> 1;
            pfh(pmwt, i + 2 * pfd); pfh(hvwt, i + 2 * pfd); pfh(W.rbint, i + V1 + 2 * pfd); pf(cfa, i + V1 + 2 * pfd);
            pf(W.dirwts0, i + V1 + 2 * pfd); pf(W.dirwts1, i + 2 * pfd);
            if (ab(0.5f - pmwt[i1]) < ab(0.5f - hvwt[i1])) continue;
            const float *rbint = W.rbint, *d0 = W.dirwts0, *d1 = W.dirwts1;
            const float rb = rbint[i1];
            const float cru = (float)((double)cfa[i - V1] * 2.0 / (double)(AMZ_EPS + rb + rbint[i1 - V1]));
            const float crd = (float)((double)cfa[i + V1] * 2.0 / (double)(AMZ_EPS + rb + rbint[i1 + V1]));
            const float crl = (float)((double)cfa[i - 1] * 2.0 / (double)(AMZ_EPS + rb + rbint[i1 - 1]));
            const float crr = (float)((double)cfa[i + 1] * 2.0 / (double)(AMZ_EPS + rb + rbint[i1 + 1]));
            const float gu = ab(1.0f - cru) < AMZ_ARTHRESH ? rb * cru : cfa[i - V1] + expdec(rb - rbint[i1 - V1], 1);
            const float gd = ab(1.0f - crd) < AMZ_ARTHRESH ? rb * crd : cfa[i + V1] + expdec(rb - rbint[i1 + V1], 1);
            const float gl = ab(1.0f - crl) < AMZ_ARTHRESH ? rb * crl : cfa[i - 1] + expdec(rb - rbint[i1 - 1], 1);
            const float gr = ab(1.0f - crr) < AMZ_ARTHRESH ? rb * crr : cfa[i + 1] + expdec(rb - rbint[i1 + 1], 1);
            float Gintv = (d0[i - V1] * gd + d0[i + V1] * gu) / (d0[i + V1] + d0[i - V1]);
            float Ginth = (d1[i - 1] * gr + d1[i + 1] * gl) / (d1[i - 1] + d1[i + 1]);
            if (Gintv < rb) {
                if (2 * Gintv < rb)
                    Gintv = ulims(Gintv, cfa[i - V1], cfa[i + V1]);
                else {
                    const float vw = (float)(2.0 * (double)(rb - Gintv) / (double)(AMZ_EPS + Gintv + rb));
                    Gintv = vw * Gintv + (1.0f - vw) * ulims(Gintv, cfa[i - V1], cfa[i + V1]);
                }
            }
            if (Ginth < rb) {
                if (2 * Ginth < rb)
                    Ginth = ulims(Ginth, cfa[i - 1], cfa[i + 1]);
                else {
                    const float hw = (float)(2.0 * (double)(rb - Ginth) / (double)(AMZ_EPS + Ginth + rb));
                    Ginth = hw * Ginth + (1.0f - hw) * ulims(Ginth, cfa[i - 1], cfa[i + 1]);
                }
            }
            if (Ginth > AMZ_CLIP) Ginth = ulims(Ginth, cfa[i - 1], cfa[i + 1]);
            if (Gintv > AMZ_CLIP) Gintv = ulims(Gintv, cfa[i - V1], cfa[i + V1]);
            const float g = Ginth * (1.0f - hvwt[i1]) + Gintv * hvwt[i1];
            rgbgreen[i] = g;
            W.Dgrb0[i1] = g - cfa[i];
        }
    }
    C.sync();
    C.mark(13);

    // ---- split G-B out of the G-R plane at the B sites (:1358-1362) ----
    {
        const int nr = cdiv(rr1 - 12 - 13, 2), ns = cdiv(cc1 - 12 - 13, 2);
        Strider sd12(tid, nthr, ns);
        for (int idx = tid; idx < nr * ns; idx += nthr, sd12.step()) {
            const int rr = 13 + 2 * (sd12.row), cc = 13 + 2 * sd12.col, i1 = (rr * TS + cc) >> 1;
            W.Dgrb1[i1] = W.Dgrb0[i1];
            W.Dgrb0[i1] = 0.0f;
        }
    }
    C.sync();
    C.mark(14);

    // ---- chroma at the opposite-colour sites from the four diagonal neighbours (:1369-1383) ----
    for (int par = 0; par < 2; par++) {
        const int ns = 4 * cdiv(cc1 - 28 - par, 8), nr = (rr1 - 28 + 1 - par) / 2;
        float *const D = par ? W.Dgrb0 : W.Dgrb1;                          // c = 1 - FC/2: R rows fill G-B, B rows fill G-R
        Strider sd13(tid, nthr, ns);
        for (int idx = tid; idx < nr * ns; idx += nthr, sd13.step()) {
            const int rr = 14 + par + 2 * (sd13.row), cc = 14 + par + 2 * sd13.col, i = rr * TS + cc;
#define AMZ_G(o) D[(i + (o)) >> 1]
            pfh(D, i + M3 + 2 * pfd);
            const float wtnw = 1.0f / (AMZ_EPS + ab(AMZ_G(-M1) - AMZ_G(M1)) + ab(AMZ_G(-M1) - AMZ_G(-M3)) + ab(AMZ_G(M1) - AMZ_G(-M3)));
            const float wtne = 1.0f / (AMZ_EPS + ab(AMZ_G(P1) - AMZ_G(-P1)) + ab(AMZ_G(P1) - AMZ_G(P3)) + ab(AMZ_G(-P1) - AMZ_G(P3)));
            const float wtsw = 1.0f / (AMZ_EPS + ab(AMZ_G(-P1) - AMZ_G(P1)) + ab(AMZ_G(-P1) - AMZ_G(M3)) + ab(AMZ_G(P1) - AMZ_G(-P3)));
            const float wtse = 1.0f / (AMZ_EPS + ab(AMZ_G(M1) - AMZ_G(-M1)) + ab(AMZ_G(M1) - AMZ_G(-P3)) + ab(AMZ_G(-M1) - AMZ_G(M3)));
            D[i >> 1] = (wtnw * (1.325f * AMZ_G(-M1) - 0.175f * AMZ_G(-M3) - 0.075f * AMZ_G(-M1 - 2) - 0.075f * AMZ_G(-M1 - V2)) +
                         wtne * (1.325f * AMZ_G(P1) - 0.175f * AMZ_G(P3) - 0.075f * AMZ_G(P1 + 2) - 0.075f * AMZ_G(P1 + V2)) +
                         wtsw * (1.325f * AMZ_G(-P1) - 0.175f * AMZ_G(-P3) - 0.075f * AMZ_G(-P1 - 2) - 0.075f * AMZ_G(-P1 - V2)) +
                         wtse * (1.325f * AMZ_G(M1) - 0.175f * AMZ_G(M3) - 0.075f * AMZ_G(M1 + 2) - 0.075f * AMZ_G(M1 + V2))) /
                        (wtnw + wtne + wtsw + wtse);
#undef AMZ_G
        }
    }
    C.sync();
    C.mark(15);

    // ---- write red, blue (:1400-1445) and green (:1451-1455) ----
    {
        const int nc = cc1 - 32, nrow = rr1 - 32;
        Strider sd14(tid, nthr, nc);
        for (int idx = tid; idx < nrow * nc; idx += nthr, sd14.step()) {
            const int rr = 16 + sd14.row, cc = 16 + sd14.col, i = rr * TS + cc;
            const size_t o = (size_t)(rr + top) * stride + cc + left;
            const bool is_green = ((rr + cc) & 1) != 0;
            pf(rgbgreen, i + pfd); pfh(hvwt, i + V1 + pfd); pfh(W.Dgrb0, i + V1 + pfd); pfh(W.Dgrb1, i + V1 + pfd);
            const float g = rgbgreen[i];
            if (is_green) {
                const float wu = hvwt[(i - V1) >> 1], wr = 1.0f - hvwt[(i + 1) >> 1], wl = 1.0f - hvwt[(i - 1) >> 1], wd = hvwt[(i + V1) >> 1];
                const float temp = 1.0f / (wu + wr + wl + wd);
                // the finished planes are streamed out (evict-first): they must not push workspace lines out of the L2
                C.st_stream(red + o, 65535.0f * (g - (wu * W.Dgrb0[(i - V1) >> 1] + wr * W.Dgrb0[(i + 1) >> 1] + wl * W.Dgrb0[(i - 1) >> 1] + wd * W.Dgrb0[(i + V1) >> 1]) * temp));
                C.st_stream(blue + o, 65535.0f * (g - (wu * W.Dgrb1[(i - V1) >> 1] + wr * W.Dgrb1[(i + 1) >> 1] + wl * W.Dgrb1[(i - 1) >> 1] + wd * W.Dgrb1[(i + V1) >> 1]) * temp));
            } else {
                C.st_stream(red + o, 65535.0f * (g - W.Dgrb0[i >> 1]));
                C.st_stream(blue + o, 65535.0f * (g - W.Dgrb1[i >> 1]));
            }
        }
        const int ng = 4 * cdiv(cc1 - 35, 4);
        Strider sd15(tid, nthr, ng);
        for (int idx = tid; idx < nrow * ng; idx += nthr, sd15.step()) {
            const int rr = 16 + sd15.row, cc = 16 + sd15.col;
            C.st_stream(green + (size_t)(rr + top) * stride + cc + left, rgbgreen[rr * TS + cc] * 65535.0f);
        }
    }
}

}  // namespace amaze
