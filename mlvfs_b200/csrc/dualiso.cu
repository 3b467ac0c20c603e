// dualiso.cu -- full dual-ISO conversion ("cr2hdr 20-bit"), both interpolation methods.
//
// Replaces reference hdr.c:1932-1957 (cr2hdr20_convert_data) and hdr_interpolate hdr.c:1774-1930 with
// its stages: hdr_check :407, identify_rggb_or_gbrg :441, identify_bright_and_dark_fields :497,
// white_detect :250, convert_to_20bit :825, match_exposures :638, build_ev2raw_lut :839,
// mean32_interpolate :1231 (mean2/mean3 :341-368), border_interpolate :1306, fullres_reconstruction
// :1355, mix_images :1524, hdr_chroma_smooth :1502 (kernel in chroma.cu), build_alias_map :1382,
// final_blend :1663, convert_20_to_16bit :1760.  The AMaZE + edge-directed interpolation
// (amaze_interpolate hdr.c:954-1229, amaze_demosaic_RT.c) lives in amaze.cu / amaze_tile.cuh; its last
// step (edge_interp :940-952, :1182-1210) is fused into this file's per-pixel interpolation kernel.
//
// Structure: three frame-global statistics barriers (row-field detection -> white levels -> exposure
// matching), each a histogram / selection kernel plus a tiny scalar epilogue on the host (the same
// integer logic as the reference, O(bins)); everything per-pixel runs on the GPU.  Statistics are
// exact integers; the per-pixel blends use IEEE fp64 without FMA contraction (-fmad=false) like the
// reference's SSE2 code.  The 20-bit EV tables and fullres_curve (a function of black only) are built on the host with
// glibc, so they are the reference's own values; only the per-frame mix_curve is evaluated with CUDA's fp64 log2 / cos
// (<= 2 ulp from glibc), which can move a half-resolution blend by one EV-LUT step (1/32768 EV) at most: the stage is
// a tolerance stage (<= 1 DN on the 16-bit output) as the north star states; every test so far measures 0 DN.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "amaze.cuh"
#include "context.cuh"

namespace {

constexpr int EVR = MLVB_EV_RES;
constexpr int N20 = 1 << 20;
constexpr int ALIAS_MAP_MAX = 15000;
constexpr double FULLRES_THR = 0.8;
constexpr int DARK_NOISE = 512;         // compute_noise sees an empty window: 8 DN * 64
constexpr double DARK_NOISE_EV = 9.0;
static_assert(4 * DARK_NOISE == 2048, "final blend multiplies by the exact reciprocal of 4 * DARK_NOISE");

// div_const (common.cuh) against the IEEE quotient for every numerator the final blend can pass: alias-map values
// (uint16) over ALIAS_MAP_MAX and overexposure weights (uint8) over 200.  std::fma is exact, like the device's DFMA.
static bool fast_div_ok()
{
    static const bool ok = [] {
        auto same = [](double a, double D) {
            const double R = 1.0 / D, q = a * R, v = fma(fma(-q, D, a), R, q), want = a / D;
            return memcmp(&v, &want, sizeof v) == 0;
        };
        for (int a = 0; a <= 65535; a++) if (!same((double)a, (double)ALIAS_MAP_MAX)) return false;
        for (int a = 0; a <= 255; a++) if (!same((double)a, 200.0)) return false;
        return true;
    }();
    return ok;
}

// ------------------------------------------------------------------ phase A: frame statistics

struct StatsA {
    double ev_sum;                      // hdr_check
    unsigned long long ev_num;
    unsigned hist_cfa[4][16384];        // identify_rggb_or_gbrg
    unsigned hist_field[2][4][16384];   // identify_bright_and_dark_fields for row offset 0 (RGGB) and 1 (GBRG)
};

// hdr_check (hdr.c:407-439) and, with HIST, the CFA / row-field histograms (hdr.c:441-636) in one pass.
// HIST = false: only the hdr_check sum (the histograms are taken later, from the pixel-fixed frame).
//
// One 1024-thread block per SM; block b takes the rows of one class c = b % 4 (y % 4 == c), every thread two
// adjacent pixels.  The rows of one class feed only three histograms, which fit in shared memory as 32-bit
// counters (3 x 64 KB): S0 / S1 = even / odd columns over rows < h/4*4 (-> hist_cfa and, for the green
// parity, hist_field[0][c]) and S2 = the GBRG reading of the same rows (hist_field[1][(c + 3) % 4], green
// parity x % 2 == y % 2, rows 5 .. (h-1)/4*4).  The block adds its non-zero bins to the global tables once.
constexpr int STATS_A_THREADS = 1024;
constexpr int STATS_A_SMEM = 3 * 16384 * (int)sizeof(unsigned);

template <bool HIST>
__global__ void __launch_bounds__(STATS_A_THREADS, 1)
diso_stats_a_kernel(const uint16_t *__restrict__ img, int w, int h, int black, int white, const double *__restrict__ raw2evf,
                    StatsA *__restrict__ S)
{
    extern __shared__ unsigned sa_hist[];
    unsigned *S0 = sa_hist, *S1 = sa_hist + 16384, *S2 = sa_hist + 32768;
    const int c = blockIdx.x & 3, nper = gridDim.x >> 2;                  // gridDim.x is a multiple of 4
    if (HIST) {
        for (int i = threadIdx.x; i < 3 * 16384; i += blockDim.x) sa_hist[i] = 0u;
        __syncthreads();
    }
    double ev = 0.0;
    unsigned num = 0;
    const int h0 = h / 4 * 4, h1 = (h - 1) / 4 * 4;
    for (int y = c + 4 * (blockIdx.x >> 2); y < h; y += 4 * nper) {
        const bool in0 = y < h0, in1 = (y - 1) >= 4 && (y - 1) < h1;
        const bool inner_y = y >= 2 && y < h - 2;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(img + (size_t)y * w);
        const uint32_t *row2 = reinterpret_cast<const uint32_t *>(img + (size_t)(y + 2) * w);
        for (int xp = threadIdx.x; xp < (w >> 1); xp += blockDim.x) {     // w is even
            const uint32_t pp = __ldg(row + xp);
            const int pe = (int)(pp & 0xFFFFu), po = (int)(pp >> 16);
            if (HIST) {
                if (in0) { atomicAdd(&S0[pe & 16383], 1u); atomicAdd(&S1[po & 16383], 1u); }      // hdr.c:461-465, 540-550
                if (in1) atomicAdd(&S2[((c & 1) ? po : pe) & 16383], 1u);                          // GBRG: frame starts one row lower
            }
            if (inner_y) {                                                                         // hdr.c:419-433
                const uint32_t qq = __ldg(row2 + xp);
                const int x = 2 * xp;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int p = k ? po : pe, p2 = k ? (int)(qq >> 16) : (int)(qq & 0xFFFFu);
                    if (x + k >= 2 && x + k < w - 2 && (p > black + 32 || p2 > black + 32) && p < white && p2 < white) {
                        ev += fabs(raw2evf[p2] - raw2evf[p]);
                        num++;
                    }
                }
            }
        }
    }
    // block reduction of the hdr_check sum
    __shared__ double s_ev[32];
    __shared__ unsigned s_num[32];
    for (int o = 16; o; o >>= 1) { ev += __shfl_xor_sync(0xFFFFFFFFu, ev, o); num += __shfl_xor_sync(0xFFFFFFFFu, num, o); }
    if ((threadIdx.x & 31) == 0) { s_ev[threadIdx.x >> 5] = ev; s_num[threadIdx.x >> 5] = num; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e = 0; unsigned n = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) { e += s_ev[i]; n += s_num[i]; }
        if (n) { atomicAdd(&S->ev_sum, e); atomicAdd(&S->ev_num, (unsigned long long)n); }
    }
    if (HIST) {
        unsigned *g_cfa0 = S->hist_cfa[(c & 1) * 2], *g_cfa1 = S->hist_cfa[(c & 1) * 2 + 1];
        unsigned *g_f0 = S->hist_field[0][c], *g_f1 = S->hist_field[1][(c + 3) & 3];
        for (int i = threadIdx.x; i < 16384; i += blockDim.x) {
            const unsigned a0 = S0[i], a1 = S1[i], a2 = S2[i];
            if (a0) atomicAdd(&g_cfa0[i], a0);
            if (a1) atomicAdd(&g_cfa1[i], a1);
            const unsigned g = (c & 1) ? a0 : a1;                          // green columns of this row class: x % 2 != y % 2
            if (g) atomicAdd(&g_f0[i], g);
            if (a2) atomicAdd(&g_f1[i], a2);
        }
    }
}

static int launch_stats_a(bool hist, const uint16_t *d_img, int w, int h, int black, int white, const double *raw2evf, StatsA *S,
                          int sm_count, cudaStream_t st)
{
    if ((w & 1) || ((uintptr_t)d_img & 3)) return MLVB_ERR_UNSUPPORTED;   // pixel pairs are read as 32-bit words
    const int nblocks = std::max(4, std::min(sm_count > 0 ? sm_count : 148, (h + 3) / 4 * 4) / 4 * 4);
    if (hist) {
        cudaFuncSetAttribute(diso_stats_a_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STATS_A_SMEM);   // per device
        diso_stats_a_kernel<true><<<nblocks, STATS_A_THREADS, STATS_A_SMEM, st>>>(d_img, w, h, black, white, raw2evf, S);
    } else {
        diso_stats_a_kernel<false><<<nblocks, STATS_A_THREADS, 0, st>>>(d_img, w, h, black, white, raw2evf, S);
    }
    return MLVB_OK;
}

// ------------------------------------------------------------------ phase B: white levels (hdr.c:250-300)

struct FieldInfo { int is_bright[4]; int y1; };

__global__ void diso_white_kernel(const uint16_t *__restrict__ img, int w, int h, FieldInfo F, int max_pix,
                                  unsigned n_dark, unsigned n_bright, unsigned *__restrict__ hist /* [2][65536] */)
{
    const int nx = (w + 2) / 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int x = 3 * i, y = F.y1 + 3 * j;
    if (i >= nx || y >= h) return;
    const int c = F.is_bright[y % 4];
    // rank of this sample among its class in raster order: class rows before row j (period 4 in j)
    int per_period = 0, partial = 0;
    for (int k = 0; k < 4; k++) per_period += (F.is_bright[(F.y1 + 3 * k) % 4] == c);
    for (int k = 0; k < (j & 3); k++) partial += (F.is_bright[(F.y1 + 3 * k) % 4] == c);
    const unsigned rank = (unsigned)((j >> 2) * per_period + partial) * nx + i;
    const unsigned n_c = c ? n_bright : n_dark;
    // a full array keeps overwriting its last slot: only the last sample survives there (hdr.c:279-281)
    if (rank < (unsigned)(max_pix - 1) || rank == n_c - 1) atomicAdd(&hist[c * 65536 + img[x + (size_t)y * w]], 1u);
}

// ------------------------------------------------------------------ phase C: exposure matching statistics (hdr.c:638-722)

struct ExpoParams { int black16, clip, clip0, y0; };

__device__ __forceinline__ int p16_of(uint16_t v) { return (int)(((uint32_t)v << 2) & 0xFFFF); }   // ((v << 6) & 0xFFFFF) >> 4

__global__ void diso_expo_pairs_kernel(const uint16_t *__restrict__ img, int w, int h, FieldInfo F, ExpoParams E,
                                       int2 *__restrict__ pairs, unsigned *__restrict__ hist /* [2][HB] bright, dark */, int HB)
{
    const int nx = (w + 2) / 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    const int x = 3 * i, y = E.y0 + 3 * j;
    if (i >= nx || y >= h - 2) return;
    const int pa = p16_of(img[x + (size_t)(y - 2) * w]) - E.black16, pb = p16_of(img[x + (size_t)(y + 2) * w]) - E.black16;
    int pn = p16_of(img[x + (size_t)y * w]) - E.black16;
    int pi = (pa + pb + 1) / 2;
    if (pa >= E.clip || pb >= E.clip) pi = E.clip0;
    if (pi >= E.clip) pn = E.clip0;
    const int b = F.is_bright[y % 4] ? pn : pi, d = F.is_bright[y % 4] ? pi : pn;
    pairs[(size_t)j * nx + i] = make_int2(b, d);
    if (b < E.clip) {
        atomicAdd(&hist[min(max(b + E.black16, 0), HB - 1)], 1u);
        atomicAdd(&hist[HB + min(max(d + E.black16, 0), HB - 1)], 1u);
    }
}

__global__ void diso_highlights_kernel(const int2 *__restrict__ pairs, unsigned n, int b_lo, int b_hi, int2 *__restrict__ sel,
                                       unsigned *__restrict__ nsel, unsigned cap)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 p = pairs[i];
    if (p.x >= b_hi || p.x <= b_lo) return;                        // hdr.c:737-738
    const unsigned k = atomicAdd(nsel, 1u);
    if (k < cap) sel[k] = p;
}

// SCORE_CAND candidate slopes per CTA (hdr.c:751-772): a selected pair is loaded and converted once for all of them
constexpr int SCORE_CAND = 4;
__global__ void diso_score_kernel(const int2 *__restrict__ sel, const unsigned *__restrict__ nsel, unsigned cap,
                                  const double *__restrict__ test_a, unsigned ncand, int dmed, int bmed, unsigned *__restrict__ scores)
{
    double ta[SCORE_CAND], tb[SCORE_CAND];
    unsigned s[SCORE_CAND];
#pragma unroll
    for (int c = 0; c < SCORE_CAND; c++) {
        const unsigned k = min(blockIdx.x * SCORE_CAND + c, ncand - 1);
        ta[c] = test_a[k]; tb[c] = (double)dmed - (double)bmed * ta[c];
        s[c] = 0;
    }
    const unsigned n = min(*nsel, cap);
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 p = sel[i];
        // abs((int)v) < 50 <=> |v| < 50 for the truncating conversion (|v| is far below 2^31 here): no F2I.F64,
        // and the int -> double conversions on the add pipe (common.cuh)
        const double px = i2d(p.x), py = i2d(p.y);
#pragma unroll
        for (int c = 0; c < SCORE_CAND; c++) s[c] += (fabs(py - (px * ta[c] + tb[c])) < 50.0);
    }
    __shared__ unsigned sw[SCORE_CAND][8];
#pragma unroll
    for (int c = 0; c < SCORE_CAND; c++) {
        for (int o = 16; o; o >>= 1) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], o);
        if ((threadIdx.x & 31) == 0) sw[c][threadIdx.x >> 5] = s[c];
    }
    __syncthreads();
    if (threadIdx.x < SCORE_CAND && blockIdx.x * SCORE_CAND + threadIdx.x < ncand) {
        unsigned t = 0;
        for (unsigned i = 0; i < blockDim.x / 32; i++) t += sw[threadIdx.x][i];
        scores[blockIdx.x * SCORE_CAND + threadIdx.x] = t;
    }
}

// ------------------------------------------------------------------ phase D: per-pixel pipeline

struct PixParams {
    int w, h;
    int is_bright[4];
    int black, white, white_darkened;                 // 20-bit
    double a, b20;                                    // exposure match (hdr.c:784-808)
    double corr_ev, overlap, max_ev;                  // mixing curve (hdr.c:1540-1571)
    const int *raw2ev;                                // 20-bit tables (hdr.c:839-874)
    const int *ev2raw;                                // pointer pre-offset by 10 EV
    const double *fullres_curve;                      // [2^20], keyed by black (hdr.c:890-913: the reference keeps the same table)
    const double *mix_curve;                          // [2^20], rebuilt for every frame (hdr.c:1562-1571)
    // where the curves are flat: [0] first index with a non-zero value, [1] one past the last index with a value
    // other than 1.0, [2] first index above FULLRES_THR, [3] one past the last index not above it ([2] == [3]: the
    // threshold is a plain comparison).  Outside [0], [1] the table value is exactly 0.0 / 1.0 and is not fetched.
    const int *fullres_lim, *mix_lim;
    int use_fullres, use_alias;
    int fast_div;                                     // div_const verified on the host for every numerator (fast_div_ok)
    int method;                                       // 0 AMaZE + edge-directed, 1 mean23 (hdr.c:1888-1896)
    AmazeView amz;
};

// 14 -> 20 bit (hdr.c:825-837) + exposure matching apply (hdr.c:784-808), one sample
__device__ __forceinline__ int to20_sample(uint16_t v, int y, const PixParams &P)
{
    int p = (int)(((uint32_t)v << 6) & 0xFFFFF);
    if (p != 0) {
        // the result is clamped to >= 0, so floor and the reference's truncation agree
        if (P.is_bright[y % 4]) p = d2i_floor(i2d(p - P.black) * P.a + (double)P.black + P.b20 * P.a);
        else p = d2i_floor(u2d((unsigned)p) - P.b20 + P.b20 * P.a);
        p = min(max(p, 0), 0xFFFFF);
    }
    return p;
}

__global__ void diso_to20_kernel(const uint16_t *__restrict__ img, uint32_t *__restrict__ raw32, const PixParams P)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.w) return;
    const size_t i = x + (size_t)y * P.w;
    raw32[i] = (uint32_t)to20_sample(img[i], y, P);
}

// mean23 path: the 20-bit sample and its EV, each converted / looked up once per pixel (the interpolation reads three or
// four neighbours' EVs per pixel: as plane reads instead of conversions + table gathers).  Four pixels per thread.
__global__ void diso_to20ev_kernel(const uint16_t *__restrict__ img, uint32_t *__restrict__ raw32, int *__restrict__ ev32, const PixParams P)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= P.w) return;
    const size_t i = x4 + (size_t)y * P.w;
    if (x4 + 3 < P.w && (i & 3) == 0) {
        const uint2 v = *reinterpret_cast<const uint2 *>(img + i);
        uint4 r;
        r.x = (uint32_t)to20_sample((uint16_t)(v.x & 0xFFFF), y, P); r.y = (uint32_t)to20_sample((uint16_t)(v.x >> 16), y, P);
        r.z = (uint32_t)to20_sample((uint16_t)(v.y & 0xFFFF), y, P); r.w = (uint32_t)to20_sample((uint16_t)(v.y >> 16), y, P);
        const int4 e = make_int4(__ldg(P.raw2ev + r.x), __ldg(P.raw2ev + r.y), __ldg(P.raw2ev + r.z), __ldg(P.raw2ev + r.w));
        *reinterpret_cast<uint4 *>(raw32 + i) = r;
        *reinterpret_cast<int4 *>(ev32 + i) = e;
    } else {
        for (int k = 0; k < 4 && x4 + k < P.w; k++) {
            const uint32_t r = (uint32_t)to20_sample(img[i + k], y, P);
            raw32[i + k] = r;
            ev32[i + k] = __ldg(P.raw2ev + r);
        }
    }
}

__device__ __forceinline__ int mean2_ev(int a, int b, int white) { return (a >= white || b >= white) ? white : (a + b) / 2; }
__device__ __forceinline__ int mean3_ev(int a, int b, int c, int white)
{
    const int m = (a + b + c) / 3;
    return (a >= white || b >= white || c >= white) ? max(m, white) : m;
}

// mean32_interpolate (or the edge-directed interpolation of amaze_interpolate, hdr.c:1182-1210) +
// border_interpolate + fullres_reconstruction, one thread per pixel, from the 20-bit plane.  FROM14 (the mean23 path):
// the neighbours' EVs come from the plane diso_to20ev_kernel wrote next to it (raw32 / ev32 are converted and looked
// up once per pixel; round 1 converted the three or four neighbours of every pixel on the fly: 286 instructions and
// 4.5 table gathers per pixel).
template <bool FROM14>
__global__ void diso_interp_kernel(const int *__restrict__ ev32, const uint32_t *__restrict__ raw32, uint32_t *__restrict__ dark,
                                   uint32_t *__restrict__ bright, uint32_t *__restrict__ fullres, const PixParams P)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int w = P.w, h = P.h;
    if (x >= w) return;
#define R(xx, yy) ((int)raw32[(xx) + (size_t)(yy) * w])
#define EVP(xx, yy) (ev32[(xx) + (size_t)(yy) * w])                    // raw2ev[R(xx, yy)], from diso_to20ev_kernel
    const int br = P.is_bright[y % 4];
    uint32_t native, interp;
    // precedence = order of the loops in border_interpolate (hdr.c:1312-1352), later loops win
    if (y >= 2 && x < 2) { interp = R(x, y - 2); native = R(x, y); }
    else if (y >= 2 && x >= w - 3) { interp = R(x - 2, y - 2); native = R(x - 2, y); }
    else if (y >= h - 4) { interp = R(x, y - 2); native = R(x, y); }
    else if (y < 3) { interp = R(x, y + 2); native = R(x, y); }
    else if (!FROM14) {
        // edge-directed: average three neighbouring directions in EV space (hdr.c:1198-1206)
        const int s = (P.is_bright[y % 4] == P.is_bright[(y + 1) % 4]) ? -1 : 1;
        const int dir = P.amz.edir[x + (size_t)y * w];
        const int pi0 = amz_edge_interp(P.amz, P.raw2ev, dir, x, y, s, P.black);
        const int pip = amz_edge_interp(P.amz, P.raw2ev, min(dir + 1, 10), x, y, s, P.black);
        const int pim = amz_edge_interp(P.amz, P.raw2ev, max(dir - 1, 0), x, y, s, P.black);
        interp = (uint32_t)__ldg(P.ev2raw + (2 * pi0 + pip + pim) / 4);
        native = R(x, y);
    } else {
        // mean23 interior (hdr.c:1255-1299); pairs start at even x
        const int white = !br ? P.white_darkened : P.white;
        const int wev = __ldg(P.raw2ev + white);
        const int s = (P.is_bright[y % 4] == P.is_bright[(y + 1) % 4]) ? -1 : 1;
        const int xe = x & ~1;
        int ev;
        if ((y & 1) == 0) {
            if (x == xe) ev = mean2_ev(EVP(x, y - 2), EVP(x, y + 2), wev);
            else ev = mean3_ev(EVP(xe + 2, y + s), EVP(xe, y + s), EVP(xe + 1, y - 2 * s), wev);
        } else {
            if (x == xe) ev = mean3_ev(EVP(xe + 1, y + s), EVP(xe - 1, y + s), EVP(xe, y - 2 * s), wev);
            else ev = mean2_ev(EVP(x, y - 2), EVP(x, y + 2), wev);
        }
        interp = (uint32_t)__ldg(P.ev2raw + ev);
        native = R(x, y);
    }
#undef R
#undef EVP
    const size_t i = x + (size_t)y * w;
    const uint32_t d = br ? interp : native, b = br ? native : interp;
    dark[i] = d;
    bright[i] = b;
    uint32_t f = 0;
    if (P.use_fullres) {                                                     // hdr.c:1365-1378
        if (br) f = ((int)b < P.white_darkened) ? b : max(b, d);
        else f = d;
    }
    fullres[i] = f;
}

// the two blending curves as tables over the 20-bit bright sample, like the reference's own mix_curve /
// fullres_curve arrays: fullres_curve depends on black only, mix_curve on this frame's exposure match
__device__ __forceinline__ void curve_limits(int i, double v, int *__restrict__ lim, bool valid = true)
{
    // block-aggregated (256 threads): at most one atomic per block and limit, none where it cannot change the limit
    int a0 = v != 0.0 ? i : N20, a1 = v != 1.0 ? i + 1 : 0, a2 = v > FULLRES_THR ? i : N20, a3 = !(v > FULLRES_THR) ? i + 1 : 0;
    if (!valid) { a0 = N20; a1 = 0; a2 = N20; a3 = 0; }
    for (int o = 16; o; o >>= 1) {
        a0 = min(a0, __shfl_xor_sync(0xFFFFFFFFu, a0, o)); a1 = max(a1, __shfl_xor_sync(0xFFFFFFFFu, a1, o));
        a2 = min(a2, __shfl_xor_sync(0xFFFFFFFFu, a2, o)); a3 = max(a3, __shfl_xor_sync(0xFFFFFFFFu, a3, o));
    }
    __shared__ int s_lim[8][4];
    if ((threadIdx.x & 31) == 0) { s_lim[threadIdx.x >> 5][0] = a0; s_lim[threadIdx.x >> 5][1] = a1; s_lim[threadIdx.x >> 5][2] = a2; s_lim[threadIdx.x >> 5][3] = a3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        int r = s_lim[0][k];
        for (int wp = 1; wp < (int)(blockDim.x >> 5); wp++) r = (k & 1) ? max(r, s_lim[wp][k]) : min(r, s_lim[wp][k]);
        if (k & 1) { if (r > 0 && r > lim[k]) atomicMax(&lim[k], r); }
        else if (r < N20 && r < lim[k]) atomicMin(&lim[k], r);
    }
}
__global__ void diso_curve_lim_init_kernel(int *__restrict__ lim) { lim[0] = N20; lim[1] = 0; lim[2] = N20; lim[3] = 0; }

// Only [i0, i1) is evaluated: the host derives from the curve's formula where it is flat (exactly 0.0 up to
// ev = max_ev - overlap, exactly 1.0 from ev = max_ev on) and pads that by two sample units; entries outside are never
// read (the readers clamp the index into [lim[0], lim[1]), and both limits lie inside the evaluated range).
__global__ void diso_mix_curve_kernel(double *__restrict__ curve, int *__restrict__ lim, int black, double corr_ev, double max_ev,
                                      double overlap, int i0, int i1)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < i1;
    double k = 0.0;
    if (in) {
        const double ev = log2(fmax((double)i / 64.0 - (double)black / 64.0, 1.0)) + corr_ev;
        const double c = -cos(fmax(fmin(ev - (max_ev - overlap), overlap), 0.0) * M_PI / overlap);
        k = fmax(fmin((c + 1.0) / 2.0, 1.0), 0.0);
        curve[i] = k;
    }
    curve_limits(i, k, lim, in);
}
// Table value with the flat ends answered from the limits (exactly 0.0 below lo = lim[0], exactly 1.0 from hi = lim[1] on).
// No branch: the fetch always happens, from an index clamped into the non-flat range (a neighbour
// of what the other lanes fetch), and the flat ends are selected afterwards.  Lets the compiler issue the gathers of
// several pixels of one thread back to back.
__device__ __forceinline__ double curve_at_nb(const double *__restrict__ curve, int lo, int hi, int i)
{
    const int j = hi > lo ? min(max(i, lo), hi - 1) : 0;
    const double v = __ldg(curve + j);
    return i < lo ? 0.0 : (i >= hi ? 1.0 : v);
}

// half-res blend (hdr.c:1562-1611) + overexposure flags (hdr.c:1627-1633) + alias-map skip mask
__global__ void diso_mix_kernel(const uint32_t *__restrict__ dark, const uint32_t *__restrict__ bright, uint32_t *__restrict__ halfres,
                                uint8_t *__restrict__ over, uint8_t *__restrict__ skip, const PixParams P)
{
    // two pixels per thread (i and i + half), each stage for both before the next: see diso_final_kernel
    constexpr int K = 2;
    const size_t np = (size_t)P.w * P.h, half = (np + K - 1) / K;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    const int mlo = __ldg(P.mix_lim), mhi = __ldg(P.mix_lim + 1);
    const int tlo = __ldg(P.fullres_lim + 2), thi = __ldg(P.fullres_lim + 3);
    const int fhi1 = max(__ldg(P.fullres_lim + 1), tlo);                 // f == 1.0 from here on (and above the alias threshold)
    size_t i[K];
    bool ok[K];
    int b[K], d[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        ok[k] = t + k * half < np;
        i[k] = ok[k] ? t + k * half : t;
        b[k] = (int)bright[i[k]]; d[k] = (int)dark[i[k]];
    }
    double kk[K], fc[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        kk[k] = curve_at_nb(P.mix_curve, mlo, mhi, b[k] & 0xFFFFF);
        // fullres_above_thr: a plain comparison when the curve crosses the threshold once, else the table value
        fc[k] = tlo == thi ? 0.0 : __ldg(P.fullres_curve + (b[k] & 0xFFFFF));
    }
    double evb[K], evd[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        // a term with weight exactly 0.0 contributes exactly 0 (the table values are finite): fetched from index 0
        evb[k] = i2d(__ldg(P.raw2ev + (kk[k] < 1.0 ? b[k] : 0)));
        evd[k] = i2d(__ldg(P.raw2ev + (kk[k] > 0.0 ? d[k] : 0)));
    }
    int mixed[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double eb = kk[k] < 1.0 ? evb[k] : 0.0, ed = kk[k] > 0.0 ? evd[k] : 0.0;
        mixed[k] = (int)(eb * (1.0 - kk[k]) + ed * kk[k]);
    }
    uint32_t hr[K];
#pragma unroll
    for (int k = 0; k < K; k++) hr[k] = (uint32_t)__ldg(P.ev2raw + mixed[k]);
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!ok[k]) continue;
        halfres[i[k]] = hr[k];
        over[i[k]] = (b[k] >= P.white_darkened || d[k] >= P.white) ? 100 : 0;
        const bool sk = tlo == thi ? ((b[k] & 0xFFFFF) >= tlo) : (fc[k] > FULLRES_THR);
        // 3: besides, the final blend takes its full-resolution weight f as exactly 1 here (fullres_curve == 1.0 from
        // fullres_lim[1] on, and the dark-area limit (sig - black) / (4 DARK_NOISE) does not cut it): it never reads the
        // smoothed half-resolution sample, and the smoothed full-resolution one only where the blurred overexposure
        // flag is set -- the chroma smoothing skips such sites (chroma.cu: dead_mode)
        const bool full = (b[k] & 0xFFFFF) >= fhi1 && (int)(((uint32_t)d[k] + (uint32_t)b[k]) / 2) - P.black >= 4 * DARK_NOISE;
        skip[i[k]] = sk ? (full ? 3 : 1) : 0;
    }
}

// alias map, pass 1 (hdr.c:1397-1415)
__global__ void diso_alias1_kernel(const uint32_t *__restrict__ frs, const uint32_t *__restrict__ hrs, const uint8_t *__restrict__ skip,
                                   uint16_t *__restrict__ amap, uint16_t *__restrict__ aux, const PixParams P)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, np = (size_t)P.w * P.h;
    if (i >= np) return;
    uint16_t v = 0;
    if (!skip[i]) {
        const int f = (int)frs[i], hh = (int)hrs[i];
        const int e_lin = max(abs(f - hh) - DARK_NOISE * 3 / 2, 0), e_log = abs(__ldg(P.raw2ev + f) - __ldg(P.raw2ev + hh));
        v = (uint16_t)min(min(e_lin / 2, e_log / 16), 65530);
    }
    amap[i] = v;
    aux[i] = v;
}

// pass 2: 6th largest of the 37-point ring (hdr.c:1420-1440); writes aux
__global__ void diso_alias2_kernel(const uint16_t *__restrict__ amap, const uint8_t *__restrict__ skip, uint16_t *__restrict__ aux, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < 6 || x >= w - 6 || y < 6 || y >= h - 6) return;
    const size_t i = x + (size_t)y * w;
    if (skip[i]) return;
    int top[6] = {-1, -1, -1, -1, -1, -1};                               // six largest, descending
    auto push = [&](int v) {
        if (v <= top[5]) return;
        top[5] = v;
#pragma unroll
        for (int k = 5; k > 0; k--) if (top[k] > top[k - 1]) { int t = top[k]; top[k] = top[k - 1]; top[k - 1] = t; }
    };
#pragma unroll
    for (int dy = -6; dy <= 6; dy += 2) {
        const int reach = (dy == -6 || dy == 6) ? 2 : ((dy == -4 || dy == 4) ? 4 : 6);
        for (int dx = -reach; dx <= reach; dx += 2) push((int)amap[(x + dx) + (size_t)(y + dy) * w]);
    }
    aux[i] = (uint16_t)top[5];
}

// pass 3: the 13-tap "gaussian" with its duplicated terms (hdr.c:1443-1464), uint16 wrap -- and pass 4: 2x2 max + clamp
// (hdr.c:1466-1483), one thread per 2x2 block.  Pass 3 reads `aux` only and pass 4 stays inside its own block, so the
// two run as one kernel: pixels outside pass 3's domain (border, skip mask) keep their pass-1 value in amap.
__device__ __forceinline__ int alias_gauss(const uint16_t *__restrict__ aux, int x, int y, int w)
{
#define A(dx, dy) ((int)aux[(x + (dx)) + (size_t)(y + (dy)) * w])
    const int cross = A(0, -2) + A(-2, 0) + A(2, 0) + A(0, 2), diag = A(-2, -2) + A(2, -2) + A(-2, 2) + A(2, 2);
    const int far4 = A(0, -6) + A(-6, 0) + A(6, 0) + A(0, 6);
    const int knight = A(-2, -6) + A(2, -6) + A(-6, -2) + A(6, -2) + A(-6, 2) + A(6, 2) + A(-2, 6) + A(2, 6);
    const int c = A(0, 0) + cross * 820 / 1024 + diag * 657 / 1024 + cross * 421 / 1024 + (2 * diag) * 337 / 1024 + diag * 173 / 1024 +
                  far4 * 139 / 1024 + knight * 111 / 1024 + knight * 57 / 1024;
#undef A
    return (int)(uint16_t)c;
}

__global__ void diso_alias34_kernel(const uint16_t *__restrict__ aux, const uint8_t *__restrict__ skip, uint16_t *__restrict__ amap, int w, int h)
{
    // blocks at even (x, y); pass 4 covers 2 <= x < w - 2, 2 <= y < h - 2, pass 3 covers 6 <= x < w - 6, 6 <= y < h - 6
    const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x), y = 2 * blockIdx.y;
    if (x >= w || y >= h) return;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = x + (k & 1), yy = y + (k >> 1);
        v[k] = -1;
        if (xx < w && yy < h) {
            const size_t i = xx + (size_t)yy * w;
            const bool p3 = xx >= 6 && xx < w - 6 && yy >= 6 && yy < h - 6 && !skip[i];
            v[k] = p3 ? alias_gauss(aux, xx, yy, w) : (int)amap[i];
        }
    }
    const bool p4 = x >= 2 && x < w - 2 && y >= 2 && y < h - 2;         // then x + 1 < w and y + 1 < h as well (w, h even or not)
    if (p4 && x + 1 < w && y + 1 < h) {
        const int c = min(max(max(v[0], v[1]), max(v[2], v[3])), ALIAS_MAP_MAX);
        v[0] = v[1] = v[2] = v[3] = c;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = x + (k & 1), yy = y + (k >> 1);
        if (xx < w && yy < h) amap[xx + (size_t)yy * w] = (uint16_t)v[k];
    }
}

// 3x3 "blur" of the overexposure flags (hdr.c:1636-1655): in -> out.  The flags are 0 / 100 and final_blend only
// uses min(blurred / 200, 1): both planes are bytes, the blurred value saturates at 200.
// Four pixels per thread when the rows are word aligned (w % 4 == 0): per row one 32-bit load and the two bytes next
// to it instead of twelve byte loads, one 32-bit store.
// Also folds "the blurred flag is 0" into the site flags of the mix kernel: flags 3 -> 7 (chroma smoothing of the
// full-resolution plane skips sites with all three bits, of the half-resolution plane those with the low two).
__global__ void diso_over_blur_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, uint8_t *__restrict__ flags, int w, int h)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= w) return;
    const size_t i0 = x0 + (size_t)y * w;
    const bool rows_in = y >= 3 && y < h - 3;
    if ((w & 3) == 0 && rows_in && x0 >= 4 && x0 + 8 <= w) {
        int r[3][6];                                                      // rows y-1 .. y+1, columns x0-1 .. x0+4
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const uint8_t *p = in + i0 + (ptrdiff_t)(j - 1) * w;
            const uint32_t m = *reinterpret_cast<const uint32_t *>(p);
            r[j][0] = p[-1]; r[j][5] = p[4];
            r[j][1] = m & 0xFF; r[j][2] = (m >> 8) & 0xFF; r[j][3] = (m >> 16) & 0xFF; r[j][4] = m >> 24;
        }
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int v = r[1][k + 1];
            if (x0 + k < w - 3)                                           // x0 + k >= 3 holds (x0 >= 4)
                v = r[1][k + 1] + (r[0][k + 1] + r[1][k] + r[1][k + 2] + r[2][k + 1]) * 820 / 1024 +
                    (r[0][k] + r[0][k + 2] + r[2][k] + r[2][k + 2]) * 657 / 1024;
            o |= (uint32_t)min(v, 200) << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(out + i0) = o;
        uint32_t f = *reinterpret_cast<const uint32_t *>(flags + i0), g = f;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (((f >> (8 * k)) & 3u) == 3u && ((o >> (8 * k)) & 0xFFu) == 0u) g |= 4u << (8 * k);
        if (g != f) *reinterpret_cast<uint32_t *>(flags + i0) = g;
        return;
    }
    for (int k = 0; k < 4 && x0 + k < w; k++) {
        const int x = x0 + k;
        const size_t i = i0 + k;
        int v = in[i];
        if (x >= 3 && x < w - 3 && rows_in) {
#define O(dx, dy) ((int)in[(x + (dx)) + (size_t)(y + (dy)) * w])
            v = O(0, 0) + (O(0, -1) + O(-1, 0) + O(1, 0) + O(0, 1)) * 820 / 1024 + (O(-1, -1) + O(1, -1) + O(-1, 1) + O(1, 1)) * 657 / 1024;
#undef O
        }
        out[i] = (uint8_t)min(v, 200);
        if ((flags[i] & 3) == 3 && min(v, 200) == 0) flags[i] |= 4;
    }
}

// final_blend (hdr.c:1691-1752) + convert_20_to_16bit (hdr.c:1760-1772)
__global__ void diso_final_kernel(const uint32_t *__restrict__ dark, const uint32_t *__restrict__ bright, const uint32_t *__restrict__ fullres,
                                  const uint32_t *__restrict__ frs, const uint32_t *__restrict__ hrs, const uint8_t *__restrict__ over,
                                  const uint16_t *__restrict__ amap, uint16_t *__restrict__ out16, const PixParams P)
{
    // Two rows per thread, every stage written for both pixels before the next one: the kernel is a chain of four
    // dependent memory accesses per pixel (sample -> curve -> EV tables -> inverse table) and waits on their latency,
    // so a thread keeps two chains in flight.  Gathers whose weight is exactly 0.0 are fetched from index 0 instead of
    // being branched around (a finite value times 0.0 adds nothing, as before).
    constexpr int K = 2;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * K;
    if (x >= P.w) return;
    const int flo = __ldg(P.fullres_lim), fhi = __ldg(P.fullres_lim + 1);
    size_t i[K];
    bool ok[K];
    uint32_t b[K], d[K], fr[K], fs[K], hs[K];
    int ov[K], am[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        ok[k] = y0 + k < P.h;
        i[k] = x + (size_t)(ok[k] ? y0 + k : y0) * P.w;
        b[k] = bright[i[k]]; d[k] = dark[i[k]]; fr[k] = fullres[i[k]]; fs[k] = frs[i[k]]; hs[k] = hrs[i[k]];
        ov[k] = over[i[k]];
        am[k] = P.use_alias ? (int)amap[i[k]] : 0;
    }
    double f[K], noo[K];
#pragma unroll
    for (int k = 0; k < K; k++) f[k] = curve_at_nb(P.fullres_curve, flo, fhi, (int)b[k] & 0xFFFFF);
#pragma unroll
    for (int k = 0; k < K; k++) {
        // both quotients are >= 0 (integer numerators), so COERCE(x, 0, 1) is min(x, 1)
        double c = 0.0;
        if (P.use_alias)
            c = fmin(P.fast_div ? div_const(u2d((unsigned)am[k]), (double)ALIAS_MAP_MAX, 1.0 / (double)ALIAS_MAP_MAX)
                                : (double)am[k] / (double)ALIAS_MAP_MAX, 1.0);
        const double ovf = fmin(P.fast_div ? div_const(u2d((unsigned)ov[k]), 200.0, 1.0 / 200.0) : (double)ov[k] / 200.0, 1.0);
        c = fmax(c, ovf);
        noo[k] = fmax(ovf, 1.0 - f[k]);
        f[k] = fmax(f[k], c);
        const int sig = (int)((d[k] + b[k]) / 2);
        f[k] = fmax(0.0, fmin(f[k], i2d(sig - P.black) * (1.0 / (double)(4 * DARK_NOISE))));     // 4 * DARK_NOISE = 2^11: exact
    }
    // the three EV gathers are weighted by (1 - f), f * noo and f * (1 - noo)
    double hrev[K], frsev[K], frev[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        hrev[k] = i2d(__ldg(P.raw2ev + (f[k] < 1.0 ? hs[k] : 0u)));
        frsev[k] = i2d(__ldg(P.raw2ev + ((f[k] > 0.0 && noo[k] > 0.0) ? fs[k] : 0u)));
        frev[k] = i2d(__ldg(P.raw2ev + ((f[k] > 0.0 && noo[k] < 1.0) ? fr[k] : 0u)));
    }
    int o[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double h0 = f[k] < 1.0 ? hrev[k] : 0.0, s0 = (f[k] > 0.0 && noo[k] > 0.0) ? frsev[k] : 0.0;
        const double r0 = (f[k] > 0.0 && noo[k] < 1.0) ? frev[k] : 0.0;
        const double fev = noo[k] * s0 + (1.0 - noo[k]) * r0;
        o[k] = (int)(h0 * (1.0 - f[k]) + fev * f[k]);
        o[k] = min(max(o[k], -10 * EVR), 14 * EVR - 1);
    }
    uint32_t v20[K];
#pragma unroll
    for (int k = 0; k < K; k++) v20[k] = (uint32_t)__ldg(P.ev2raw + o[k]);
#pragma unroll
    for (int k = 0; k < K; k++)
        if (ok[k]) out16[i[k]] = (uint16_t)min(d2i_floor(u2d(v20[k]) * 0.0625 + 0.5), 0xFFFF);          // >= 0.5: floor = truncation
}

// ------------------------------------------------------------------ host: tables and scalar epilogues

// hdr.c:839-874, built with the host libm so every entry equals the reference's
void build_luts20(int black, int white, std::vector<int> &raw2ev, std::vector<int> &ev2raw_0)
{
    raw2ev.resize(N20);
    ev2raw_0.assign(24 * EVR, 0);
    int *ev2raw = ev2raw_0.data() + 10 * EVR;
    for (int i = 0; i < N20; i++) {
        const double signal = std::max(i / 64.0 - black / 64.0, -1023.0);
        raw2ev[i] = signal > 0 ? (int)round(log2(1 + signal) * EVR) : -(int)round(log2(1 - signal) * EVR);
    }
    for (int i = -10 * EVR; i < 0; i++)
        ev2raw[i] = (int)std::max(std::min(black + 64 - round(64 * pow(2, ((double)-i / EVR))), (double)black), 0.0);
    for (int i = 0; i < 14 * EVR; i++) {
        ev2raw[i] = (int)std::max(std::min(black - 64 + round(64 * pow(2, ((double)i / EVR))), (double)(N20 - 1)), (double)black);
        if (i >= raw2ev[white]) ev2raw[i] = std::max(ev2raw[i], white);
    }
    ev2raw[raw2ev[0]] = 0;
}

// smallest r with prefix[r] >= ref, prefix[r] = sum of hist[0..r)
int cursor_at(const std::vector<long long> &prefix, long long ref)
{
    return (int)(std::lower_bound(prefix.begin(), prefix.end(), ref) - prefix.begin());
}

// identify_bright_and_dark_fields (hdr.c:497-636) from the four green histograms, closed form of the walk
bool fields_from_hist(const unsigned hist[4][16384], int black, int is_bright[4])
{
    const int white = 10000;
    std::vector<long long> pre[4];
    for (int i = 0; i < 4; i++) {
        pre[i].resize(16385);
        pre[i][0] = 0;
        for (int r = 0; r < 16384; r++) pre[i][r + 1] = pre[i][r] + hist[i][r];
    }
    const long long total = pre[0][16384];
    const int ref_max = (int)(total * 0.998), ref_off = (int)(total * 0.05);
    int raw[4] = {0, 0, 0, 0}, off[4] = {0, 0, 0, 0};
    if (ref_max > 0) {
        long long ref_break = ref_max;                       // first ref at which some cursor reaches `white`
        for (int i = 0; i < 4; i++) ref_break = std::min(ref_break, pre[i][white - 1] + 1);   // cursor >= white <=> prefix[white-1] < ref
        const long long ref_final = std::min<long long>(ref_break, ref_max - 1);
        for (int i = 0; i < 4; i++) raw[i] = cursor_at(pre[i], ref_final);
        const int T = black + (white - black) / 4;
        if (T > 0) {
            long long ref_o = std::min<long long>(ref_off - 1, ref_final);
            for (int i = 0; i < 4; i++) ref_o = std::min(ref_o, pre[i][std::min(T - 1, 16384)]);   // cursor < T  <=>  prefix[T-1] >= ref
            if (ref_o >= 0)
                for (int i = 0; i < 4; i++) off[i] = cursor_at(pre[i], ref_o);
        }
    }
    int s[4];
    for (int i = 0; i < 4; i++) { raw[i] -= off[i]; s[i] = raw[i]; }
    std::sort(s, s + 4);
    const double median_bright = (s[1] + s[2]) / 2;
    for (int i = 0; i < 4; i++) is_bright[i] = raw[i] > median_bright;
    if (is_bright[0] + is_bright[1] + is_bright[2] + is_bright[3] != 2) return false;
    if (is_bright[0] == is_bright[2] || is_bright[1] == is_bright[3]) return false;
    return true;
}

// k-th smallest (0-based) of the multiset described by hist (value = bin - bias)
int kth_from_hist(const unsigned *hist, int nbins, long long k, int bias)
{
    long long acc = 0;
    for (int i = 0; i < nbins; i++) {
        acc += hist[i];
        if (acc > k) return i - bias;
    }
    return 0;
}

struct DisoScratch {        // carved out of the slot's aux buffer
    StatsA *statsA;
    unsigned *hist_white;   // [2][65536]
    unsigned *hist_expo;    // [2][HB]
    int2 *pairs, *sel;
    unsigned *nsel, *scores;
    uint32_t *raw32, *dark, *bright, *fullres, *halfres, *frs, *hrs;
    uint8_t *over, *over2;
    uint16_t *amap, *aux;
    uint8_t *skip;
    double *mix_curve;      // [2^20]
    int *mix_lim;           // [4]
    AmazeScratch amz;
};

constexpr int HB = 65536 + 8;

size_t carve(uint8_t *base, int w, int h, int interp_method, DisoScratch *S)
{
    const size_t npix = (size_t)w * h, ngrid = (size_t)((w + 2) / 3 + 1) * ((h + 2) / 3 + 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { void *p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return p; };
    DisoScratch s;
    s.statsA = (StatsA *)take(sizeof(StatsA));
    s.hist_white = (unsigned *)take(2 * 65536 * sizeof(unsigned));
    s.hist_expo = (unsigned *)take(2 * HB * sizeof(unsigned));
    s.pairs = (int2 *)take(ngrid * sizeof(int2));
    s.sel = (int2 *)take(ngrid * sizeof(int2));
    s.nsel = (unsigned *)take(256);
    s.scores = (unsigned *)take(4096 * sizeof(unsigned));
    s.raw32 = (uint32_t *)take(npix * 4); s.dark = (uint32_t *)take(npix * 4); s.bright = (uint32_t *)take(npix * 4);
    s.fullres = (uint32_t *)take(npix * 4); s.halfres = (uint32_t *)take(npix * 4);
    s.frs = (uint32_t *)take(npix * 4); s.hrs = (uint32_t *)take(npix * 4);
    s.over = (uint8_t *)take(npix); s.over2 = (uint8_t *)take(npix);
    s.amap = (uint16_t *)take(npix * 2); s.aux = (uint16_t *)take(npix * 2);
    s.skip = (uint8_t *)take(npix);
    s.mix_curve = (double *)take((size_t)N20 * sizeof(double));
    s.mix_lim = (int *)take(4 * sizeof(int));
    memset(&s.amz, 0, sizeof(s.amz));
    if (interp_method == 0) o += amaze_scratch_bytes(w, h, &s.amz, base ? base + o : nullptr);
    if (S) *S = s;
    return o;
}

}  // namespace

size_t dual_iso_scratch_bytes(int w, int h, int interp_method)
{
    return carve(nullptr, w, h, interp_method, nullptr);
}

// Per-context dual-ISO tables: the 20-bit EV LUTs keyed by black like the reference's statics (built
// from the first converted frame's white level, hdr.c:1089-1093), the fp64 log table of hdr_check and
// the 3000 candidate slopes of match_exposures.
// One immutable set of 20-bit tables for a black level.  Frames in flight on any stream keep reading the set they
// started with: a set is never rewritten, a black-level change publishes a NEW set and retires the old one, and
// retired sets are only freed after a device-wide synchronisation (or with the context).
struct Lut20Set {
    int black = -1, white = 0;
    int *d_raw2ev = nullptr, *d_ev2raw_0 = nullptr;
    double *d_fullres_curve = nullptr;  // 2^20 doubles
    int *d_fullres_lim = nullptr;       // 4 ints, see PixParams::fullres_lim
    ~Lut20Set()
    {
        if (d_raw2ev) cudaFree(d_raw2ev);
        if (d_ev2raw_0) cudaFree(d_ev2raw_0);
        if (d_fullres_curve) cudaFree(d_fullres_curve);
        if (d_fullres_lim) cudaFree(d_fullres_lim);
    }
};

struct DualIsoTables {
    std::mutex mu;
    std::vector<void *> pinned_free;    // host staging for the statistics read-backs (pinned: async, full PCIe rate)
    std::shared_ptr<Lut20Set> current;  // keyed by black, built with the white level of the frame that created it
    std::vector<std::shared_ptr<Lut20Set>> retired;
    double *d_raw2evf = nullptr;        // 16384 + MAX_BLACK doubles
    double *d_test_a = nullptr;
    std::vector<double> test_a;
};

static std::mutex g_tabs_mu;
static std::map<mlvb_context *, DualIsoTables *> g_tabs;

static DualIsoTables *tables_of(mlvb_context *ctx)
{
    std::lock_guard<std::mutex> lk(g_tabs_mu);
    auto it = g_tabs.find(ctx);
    if (it != g_tabs.end()) return it->second;
    DualIsoTables *t = new DualIsoTables();
    g_tabs[ctx] = t;
    return t;
}

// called by mlvb_context_destroy (the context's device is current, all work has finished)
void dual_iso_free_tables(mlvb_context *ctx)
{
    DualIsoTables *t = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_tabs_mu);
        auto it = g_tabs.find(ctx);
        if (it == g_tabs.end()) return;
        t = it->second;
        g_tabs.erase(it);
    }
    for (void *p : t->pinned_free) cudaFreeHost(p);
    t->current.reset();
    t->retired.clear();
    if (t->d_raw2evf) cudaFree(t->d_raw2evf);
    if (t->d_test_a) cudaFree(t->d_test_a);
    delete t;
}

constexpr size_t PINNED_STAGE_BYTES = sizeof(StatsA) + 2 * 65536 * sizeof(unsigned) + 2 * (65536 + 8) * sizeof(unsigned) +
                                      4096 * sizeof(unsigned) + 256;

// Statistics read-back without the copy engines: a kernel stores the bytes straight into the pinned host buffer
// (pinned memory is device-addressable under unified addressing).  A frame waits four times for such a read-back, and
// behind the bulk frame copies of a host batch a cudaMemcpyAsync of a few hundred KB queues for milliseconds
// (measured: 16 C3 frames took 10.5 ms on the lanes while another batch's 189 MB went back to the host, 6.5 ms alone).
__global__ void diso_readback_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void diso_readback_tail_kernel(unsigned char *__restrict__ dst, const unsigned char *__restrict__ src, size_t n)
{
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
static cudaError_t readback(void *host_pinned, const void *dev, size_t bytes, cudaStream_t st)
{
    static const bool by_kernel = getenv("MLVB_READBACK_KERNEL") != nullptr;
    if (!by_kernel || (((uintptr_t)host_pinned | (uintptr_t)dev) & 15)) return cudaMemcpyAsync(host_pinned, dev, bytes, cudaMemcpyDeviceToHost, st);
    const size_t n16 = bytes / 16, tail = bytes - n16 * 16;
    if (n16) diso_readback_kernel<<<(unsigned)std::min<size_t>((n16 + 255) / 256, 296), 256, 0, st>>>((uint4 *)host_pinned, (const uint4 *)dev, n16);
    if (tail) diso_readback_tail_kernel<<<1, 32, 0, st>>>((unsigned char *)host_pinned + n16 * 16, (const unsigned char *)dev + n16 * 16, tail);
    return cudaGetLastError();
}

struct PinnedLease {
    DualIsoTables *T;
    void *p = nullptr;
    explicit PinnedLease(DualIsoTables *t) : T(t)
    {
        if (!T) return;                                   // an empty lease (the caller brings its own)
        {
            std::lock_guard<std::mutex> lk(T->mu);
            if (!T->pinned_free.empty()) { p = T->pinned_free.back(); T->pinned_free.pop_back(); }
        }
        if (!p && cudaHostAlloc(&p, PINNED_STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
    }
    void release()
    {
        if (!p) return;
        std::lock_guard<std::mutex> lk(T->mu);
        T->pinned_free.push_back(p);
        p = nullptr;
    }
    ~PinnedLease() { release(); }
};

void dual_iso_reset_tables(mlvb_context *ctx)
{
    DualIsoTables *t = tables_of(ctx);
    std::lock_guard<std::mutex> lk(t->mu);
    if (t->current) t->retired.push_back(std::move(t->current));      // frames in flight may still read it
    t->current.reset();
}

// hdr_interpolate on a device-resident 14-bit frame (in place -> 16-bit).  Returns 1 converted,
// 0 not dual ISO / failed (frame keeps its 14-bit content), < 0 MLVB_ERR_*.
static int hdr_interpolate_impl(mlvb_context *ctx, uint16_t *d_img, int w, int h, int black14, int interp_method, int use_fullres,
                                int use_alias_map, int cs_method, void *d_aux, cudaStream_t st, PinnedLease *stats_on_host);

int run_hdr_interpolate(mlvb_context *ctx, uint16_t *d_img, int w, int h, int black14, int interp_method, int use_fullres,
                        int use_alias_map, int cs_method, void *d_aux, cudaStream_t st)
{
    return hdr_interpolate_impl(ctx, d_img, w, h, black14, interp_method, use_fullres, use_alias_map, cs_method, d_aux, st, nullptr);
}

// stats_on_host: phase A's statistics of this very frame are already in that pinned stage (run_cr2hdr20 took them in
// the same pass as hdr_check because no pixel repair runs in between): no second pass over the frame, no second wait
static int hdr_interpolate_impl(mlvb_context *ctx, uint16_t *d_img, int w, int h, int black14, int interp_method, int use_fullres,
                                int use_alias_map, int cs_method, void *d_aux, cudaStream_t st, PinnedLease *stats_on_host)
{
    if (w <= 0 || h <= 0) return 0;
    if (w < 16 || h < 16 || (w & 1)) return MLVB_ERR_UNSUPPORTED;
    if (interp_method == 0 && (w & 3)) {
        fprintf(stderr, "libmlvfs_b200: dual ISO --amaze-edge needs a width that is a multiple of 4 (use --mean23)\n");
        return MLVB_ERR_UNSUPPORTED;
    }
    DualIsoTables *T = tables_of(ctx);
    {
        std::lock_guard<std::mutex> lk(T->mu);
        if (!T->d_raw2evf) {
            const size_t n = 16384 + MLVB_MAX_BLACK;
            MLVB_CUDA_OK(cudaMalloc(&T->d_raw2evf, n * sizeof(double)));
            MLVB_CUDA_OK(cudaMemcpy(T->d_raw2evf, host_raw2evf_base(), n * sizeof(double), cudaMemcpyHostToDevice));
            for (double ev = 0; ev < 6; ev += 0.002) T->test_a.push_back(pow(2, -ev));          // hdr.c:753-755
            MLVB_CUDA_OK(cudaMalloc(&T->d_test_a, T->test_a.size() * sizeof(double)));
            MLVB_CUDA_OK(cudaMemcpy(T->d_test_a, T->test_a.data(), T->test_a.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    if (black14 > MLVB_MAX_BLACK) return 0;

    const size_t npix_full = (size_t)w * h;
    DisoScratch D;
    carve((uint8_t *)d_aux, w, h, interp_method, &D);
    (void)npix_full;

    // ---------------- phase A
    PinnedLease own_stage(stats_on_host ? nullptr : T);
    PinnedLease &stage = stats_on_host ? *stats_on_host : own_stage;
    if (!stage.p) return MLVB_ERR_CUDA;
    uint8_t *hostA = (uint8_t *)stage.p;
    unsigned *hw = (unsigned *)(hostA + sizeof(StatsA)), *he = hw + 2 * 65536, *scores = he + 2 * HB;
    if (!stats_on_host) {
        MLVB_CUDA_OK(cudaMemsetAsync(D.statsA, 0, sizeof(StatsA), st));
        if (launch_stats_a(true, d_img, w, h, black14, 0x7FFFFFFF, T->d_raw2evf + (MLVB_MAX_BLACK - black14), D.statsA, ctx->sm_count, st))
            return MLVB_ERR_UNSUPPORTED;
        ctx->launches += 1;
        MLVB_CUDA_OK(readback(hostA, D.statsA, sizeof(StatsA), st));
        MLVB_CUDA_OK(stream_wait(ctx, st));
    }
    const StatsA *A = (const StatsA *)hostA;

    // identify_rggb_or_gbrg (hdr.c:467-494)
    double d_rggb = 0, d_gbrg = 0;
    {
        long long acc[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16384; i++) {
            for (int k = 0; k < 4; k++) acc[k] += A->hist_cfa[k][i];
            d_rggb += (double)llabs(acc[1] - acc[2]);
            d_gbrg += (double)llabs(acc[0] - acc[3]);
        }
    }
    const int rggb = d_rggb < d_gbrg;
    FieldInfo F;
    F.y1 = rggb ? 0 : 1;
    if (!rggb) { d_img += w; h--; }                                   // hdr.c:1784-1791
    if (!fields_from_hist(A->hist_field[rggb ? 0 : 1], black14, F.is_bright)) return 0;

    const int black = black14 * 64;
    // ---------------- phase B: white levels
    const int max_pix = w * h / 2 / 9;
    unsigned n_class[2] = {0, 0};
    const int nx = (w + 2) / 3;
    for (int y = F.y1; y < h; y += 3) n_class[F.is_bright[y % 4]] += nx;
    MLVB_CUDA_OK(cudaMemsetAsync(D.hist_white, 0, 2 * 65536 * sizeof(unsigned), st));
    const int ny_w = (h - F.y1 + 2) / 3;
    if (ny_w > 0) diso_white_kernel<<<dim3(ceil_div(nx, 128), ny_w), 128, 0, st>>>(d_img, w, h, F, max_pix, n_class[0], n_class[1], D.hist_white);
    ctx->launches += 1;
    MLVB_CUDA_OK(readback(hw, D.hist_white, 2 * 65536 * sizeof(unsigned), st));
    MLVB_CUDA_OK(stream_wait(ctx, st));
    int whites[2];
    for (int c = 0; c < 2; c++) {
        const long long kept = std::min<long long>(n_class[c], std::max(max_pix, 0));
        const int k = c ? 50 : 10, margin = c ? 1500 : 100;
        int kth_largest = 0;                                           // kth_smallest_int of the negated values (hdr.c:286-287)
        if (kept > 0) {
            long long acc = 0;
            for (int v = 65535; v >= 0; v--) { acc += hw[c * 65536 + v]; if (acc > k) { kth_largest = v; break; } }
        }
        whites[c] = kth_largest - margin;
    }
    const int white = std::min(std::max(whites[0], 10000), 16383) * 64;
    const int white_bright = std::min(std::max(whites[1], 5000), 16383) * 64;

    // ---------------- phase C: exposure matching (hdr.c:638-823)
    int white_darkened = white_bright;
    const int white20 = std::min(white, white_darkened);
    ExpoParams E;
    E.black16 = black / 16;
    const int white16 = white20 / 16;
    E.clip0 = white16 - E.black16;
    E.clip = (int)(E.clip0 * 0.95);
    E.y0 = F.y1 + 2;
    const int ny_e = (h - 2 - E.y0 + 2) / 3;
    const unsigned ngrid = (unsigned)std::max(ny_e, 0) * nx;
    MLVB_CUDA_OK(cudaMemsetAsync(D.hist_expo, 0, 2 * HB * sizeof(unsigned), st));
    if (ngrid) diso_expo_pairs_kernel<<<dim3(ceil_div(nx, 128), ny_e), 128, 0, st>>>(d_img, w, h, F, E, D.pairs, D.hist_expo, HB);
    ctx->launches += 1;
    MLVB_CUDA_OK(readback(he, D.hist_expo, 2 * HB * sizeof(unsigned), st));
    MLVB_CUDA_OK(stream_wait(ctx, st));
    long long n = 0;
    for (int i = 0; i < HB; i++) n += he[i];
    auto med_idx = [](long long m) { return (m & 1) ? m / 2 : m / 2 - 1; };
    int bmed = 0, b_lo = 0, b_hi = 0, dmed = 0;
    if (n > 0) {
        bmed = kth_from_hist(he, HB, med_idx(n), E.black16);
        b_lo = kth_from_hist(he, HB, n * 98 / 100, E.black16);
        b_hi = kth_from_hist(he, HB, (long long)(int)(n * 99.9 / 100), E.black16);
        dmed = kth_from_hist(he + HB, HB, med_idx(n), E.black16);
    }
    const int nmax = (w + 2) * (h + 2) / 9;
    const unsigned hi_cap = (unsigned)std::max(nmax / 50, 0);
    MLVB_CUDA_OK(cudaMemsetAsync(D.nsel, 0, sizeof(unsigned), st));
    const unsigned ncand = (unsigned)T->test_a.size();
    if (ngrid) diso_highlights_kernel<<<ceil_div(ngrid, 256), 256, 0, st>>>(D.pairs, ngrid, b_lo, b_hi, D.sel, D.nsel, ngrid);
    diso_score_kernel<<<ceil_div(ncand, SCORE_CAND), 256, 0, st>>>(D.sel, D.nsel, ngrid, T->d_test_a, ncand, dmed, bmed, D.scores);
    ctx->launches += 2;
    if (ncand > 4095) return MLVB_ERR_ARG;
    MLVB_CUDA_OK(readback(scores, D.scores, ncand * sizeof(unsigned), st));
    MLVB_CUDA_OK(readback(scores + 4096, D.nsel, 16, st));               // nsel sits in its own 256-byte slot; ncand <= 4095
    MLVB_CUDA_OK(stream_wait(ctx, st));
    const unsigned nsel = scores[4096];
    if (nsel >= hi_cap && hi_cap > 0) {
        // the reference truncates its highlight list in raster order here (hdr.c:727-745); not reproduced
        fprintf(stderr, "libmlvfs_b200: dual ISO: highlight sample cap reached (%u >= %u), frame not converted\n", nsel, hi_cap);
        return MLVB_ERR_UNSUPPORTED;
    }
    double a = 0, b = 0;
    unsigned best = 0;
    for (unsigned c = 0; c < ncand; c++)
        if (scores[c] > best) { best = scores[c]; a = T->test_a[c]; b = dmed - bmed * a; }

    PixParams P;
    memset(&P, 0, sizeof(P));
    P.w = w; P.h = h;
    memcpy(P.is_bright, F.is_bright, sizeof(P.is_bright));
    P.black = black; P.white = white;
    P.a = a; P.b20 = b * 16;
    white_darkened = (int)((white20 - black + P.b20) * a + black);
    P.white_darkened = white_darkened;
    const double factor = 1 / a;
    if (factor < 1.2 || !std::isfinite(factor)) return 0;               // "Doesn't look like interlaced ISO"
    P.corr_ev = log2(factor);
    const double lowiso_dr = log2(white - black) - DARK_NOISE_EV;
    double overlap = lowiso_dr - P.corr_ev;
    overlap -= std::min(3.0, overlap - 3);
    if (overlap < 0.5) return 0;                                         // "Overlap error"
    P.overlap = overlap;
    P.max_ev = log2(white / 64 - black / 64);
    P.use_fullres = use_fullres; P.use_alias = use_alias_map;
    P.fast_div = fast_div_ok() ? 1 : 0;

    // 20-bit EV tables, rebuilt only when black changes (with this frame's white), hdr.c:1089-1093
    {
        std::lock_guard<std::mutex> lk(T->mu);
        if (!T->current || T->current->black != black) {
            if (T->retired.size() >= 4) {                               // rare: clips with different black levels alternating
                MLVB_CUDA_OK(cudaDeviceSynchronize());
                T->retired.clear();
            }
            auto S = std::make_shared<Lut20Set>();
            std::vector<int> r2e, e2r;
            build_luts20(black, white, r2e, e2r);
            MLVB_CUDA_OK(cudaMalloc(&S->d_raw2ev, N20 * sizeof(int)));
            MLVB_CUDA_OK(cudaMalloc(&S->d_ev2raw_0, 24 * EVR * sizeof(int)));
            MLVB_CUDA_OK(cudaMemcpy(S->d_raw2ev, r2e.data(), N20 * sizeof(int), cudaMemcpyHostToDevice));
            MLVB_CUDA_OK(cudaMemcpy(S->d_ev2raw_0, e2r.data(), 24 * EVR * sizeof(int), cudaMemcpyHostToDevice));
            MLVB_CUDA_OK(cudaMalloc(&S->d_fullres_curve, (size_t)N20 * sizeof(double)));
            MLVB_CUDA_OK(cudaMalloc(&S->d_fullres_lim, 4 * sizeof(int)));
            {
                // fullres_curve depends on black only (hdr.c:890-913): built once per black level on the HOST with glibc's
                // log2 / cos, so every entry is the reference's own value (the per-frame mix_curve stays device-built)
                std::vector<double> fc((size_t)N20);
                int lim[4] = {N20, 0, N20, 0};
                for (int i = 0; i < N20; i++) {
                    const double ev2 = log2(std::max((double)i / 64.0 - (double)black / 64.0, 1.0));
                    const double c2 = -cos(std::min(std::max(ev2 - 4.0, 0.0), 4.0) * M_PI / 4.0);
                    const double f = (c2 + 1.0) / 2.0;
                    fc[i] = f;
                    if (f != 0.0 && i < lim[0]) lim[0] = i;
                    if (f != 1.0) lim[1] = i + 1;
                    if (f > FULLRES_THR && i < lim[2]) lim[2] = i;
                    if (!(f > FULLRES_THR)) lim[3] = i + 1;
                }
                MLVB_CUDA_OK(cudaMemcpy(S->d_fullres_curve, fc.data(), (size_t)N20 * sizeof(double), cudaMemcpyHostToDevice));
                MLVB_CUDA_OK(cudaMemcpy(S->d_fullres_lim, lim, sizeof(lim), cudaMemcpyHostToDevice));
            }
            S->black = black; S->white = white;
            if (T->current) T->retired.push_back(std::move(T->current));
            T->current = S;
        }
        const Lut20Set &S = *T->current;                                // stays alive (current or retired) until a device-wide sync
        P.raw2ev = S.d_raw2ev;
        P.ev2raw = S.d_ev2raw_0 + 10 * EVR;
        P.fullres_curve = S.d_fullres_curve;
        P.fullres_lim = S.d_fullres_lim;
    }
    P.mix_curve = D.mix_curve;
    P.mix_lim = D.mix_lim;
    diso_curve_lim_init_kernel<<<1, 1, 0, st>>>(D.mix_lim);
    {
        // where the mixing curve is not flat (hdr.c:1562-1571): signal x = i / 64 - black / 64 between 2^(max_ev - overlap -
        // corr_ev) and 2^(max_ev - corr_ev); below one sample unit of signal the curve is constant (x is clamped to 1)
        const double x_lo = exp2(P.max_ev - P.overlap - P.corr_ev), x_hi = exp2(P.max_ev - P.corr_ev);
        long long i0 = x_lo > 1.0 ? (long long)floor(black + 64.0 * x_lo) - 128 : 0;
        long long i1 = (long long)ceil(black + 64.0 * std::max(x_hi, 1.0)) + 128;
        i0 = std::min<long long>(std::max<long long>(i0, 0), N20);
        i1 = std::min<long long>(std::max<long long>(i1, i0), N20);
        if (!(x_lo == x_lo) || !(x_hi == x_hi)) { i0 = 0; i1 = N20; }   // NaN parameters: evaluate everything
        if (i1 > i0)
            diso_mix_curve_kernel<<<ceil_div(i1 - i0, 256), 256, 0, st>>>(D.mix_curve, D.mix_lim, black, P.corr_ev, P.max_ev, P.overlap,
                                                                         (int)i0, (int)i1);
    }
    ctx->launches += 2;

    // ---------------- phase D: per-pixel pipeline
    const size_t np = (size_t)w * h;
    const dim3 g2(ceil_div(w, 256), h);
    const int g1 = ceil_div(np, 256);
    P.method = interp_method ? 1 : 0;
    if (P.method == 0) {
        diso_to20_kernel<<<g2, 256, 0, st>>>(d_img, D.raw32, P);
        ctx->launches += 1;
        // the GBRG row skip changed h after the scratch was carved for the full frame: re-carve the AMaZE part
        // for this geometry inside the same region (never larger than the full-frame request)
        AmazeScratch A;
        amaze_scratch_bytes(w, h, &A, (uint8_t *)D.amz.rawf);
        int nl = 0;
        const int rc = launch_amaze_stage(D.raw32, w, h, black, white_darkened, F.is_bright, P.raw2ev, P.fullres_curve, P.fullres_lim, A, st, &nl);
        if (rc) return rc;
        ctx->launches += nl;
        P.amz.red = A.red; P.amz.green = A.green; P.amz.blue = A.blue; P.amz.squeezed = A.squeezed; P.amz.edir = A.edir;
        P.amz.ws = w + 16;
    }
    if (P.method == 0) diso_interp_kernel<false><<<g2, 256, 0, st>>>(nullptr, D.raw32, D.dark, D.bright, D.fullres, P);
    else {
        // the EV plane lives in the half-resolution plane's memory, which the mix kernel below writes after the last read
        int *ev32 = reinterpret_cast<int *>(D.halfres);
        diso_to20ev_kernel<<<dim3(ceil_div(ceil_div(w, 4), 128), h), 128, 0, st>>>(d_img, D.raw32, ev32, P);
        diso_interp_kernel<true><<<g2, 256, 0, st>>>(ev32, D.raw32, D.dark, D.bright, D.fullres, P);
        ctx->launches += 1;
    }
    diso_mix_kernel<<<ceil_div((np + 1) / 2, 256), 256, 0, st>>>(D.dark, D.bright, D.halfres, D.over, D.skip, P);
    ctx->launches += 2;
    // the blurred overexposure flags first: they tell the chroma smoothing where the final blend reads its output
    diso_over_blur_kernel<<<dim3(ceil_div(ceil_div(w, 4), 128), h), 128, 0, st>>>(D.over, D.over2, D.skip, w, h);
    const uint32_t *frs = D.fullres, *hrs = D.halfres;
    if (cs_method == 2 || cs_method == 3 || cs_method == 5) {
        int rc;
        if (use_fullres) {
            rc = launch_chroma_smooth_u32(D.fullres, D.frs, w, h, cs_method, P.raw2ev, P.ev2raw, st, D.skip, 7);
            if (rc) return rc;
            frs = D.frs;
            ctx->launches += 1;
        }
        rc = launch_chroma_smooth_u32(D.halfres, D.hrs, w, h, cs_method, P.raw2ev, P.ev2raw, st, D.skip, 3);
        if (rc) return rc;
        hrs = D.hrs;
        ctx->launches += 1;
    } else if (cs_method) {
        fprintf(stderr, "Unsupported chroma smooth method\n");          // hdr.c:1519
    }
    if (use_alias_map) {
        diso_alias1_kernel<<<g1, 256, 0, st>>>(frs, hrs, D.skip, D.amap, D.aux, P);
        diso_alias2_kernel<<<g2, 256, 0, st>>>(D.amap, D.skip, D.aux, w, h);
        diso_alias34_kernel<<<dim3(ceil_div((w + 1) / 2, 128), (h + 1) / 2), 128, 0, st>>>(D.aux, D.skip, D.amap, w, h);
        ctx->launches += 3;
    }
    diso_final_kernel<<<dim3(ceil_div(w, 256), (h + 1) / 2), 256, 0, st>>>(D.dark, D.bright, D.fullres, frs, hrs, D.over2, D.amap, d_img, P);
    ctx->launches += 2;
    MLVB_CUDA_OK(cudaGetLastError());
    return 1;
}

// cr2hdr20_convert_data (hdr.c:1932-1957) on a device frame: hdr_check, focus / bad pixels with the
// horizontal interpolator, hdr_interpolate.  Returns like run_hdr_interpolate.
int run_cr2hdr20(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, uint16_t *d_img, int interp_method,
                 int use_fullres, int use_alias_map, int cs_method, int fix_bad_pixels_mode, void *d_aux, cudaStream_t st)
{
    if (g.black > MLVB_MAX_BLACK) return 0;
    DualIsoTables *T = tables_of(ctx);
    (void)T;
    // hdr_check needs the statistics of the untouched frame: run phase A's reduction once here
    DisoScratch D;
    carve((uint8_t *)d_aux, g.w, g.h, interp_method, &D);
    {
        std::lock_guard<std::mutex> lk(T->mu);
        if (!T->d_raw2evf) {
            const size_t n = 16384 + MLVB_MAX_BLACK;
            MLVB_CUDA_OK(cudaMalloc(&T->d_raw2evf, n * sizeof(double)));
            MLVB_CUDA_OK(cudaMemcpy(T->d_raw2evf, host_raw2evf_base(), n * sizeof(double), cudaMemcpyHostToDevice));
            for (double ev = 0; ev < 6; ev += 0.002) T->test_a.push_back(pow(2, -ev));
            MLVB_CUDA_OK(cudaMalloc(&T->d_test_a, T->test_a.size() * sizeof(double)));
            MLVB_CUDA_OK(cudaMemcpy(T->d_test_a, T->test_a.data(), T->test_a.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    // fix_focus_pixels(.., 1) and fix_bad_pixels(.., 1) run between hdr_check and the row-field statistics
    // (hdr.c:1944-1948).  When neither has anything to do, both statistics come from one pass over the frame.
    int rc;
    std::shared_ptr<PixelList> focus, bad;
    {
        std::lock_guard<std::mutex> lk(ctx->clip_mu);
        rc = get_focus_pixel_map(ctx, hdr, g, &focus);
        if (rc) return rc;
    }
    const bool one_pass = !(focus && focus->nlevels) && !fix_bad_pixels_mode;
    MLVB_CUDA_OK(cudaMemsetAsync(D.statsA, 0, sizeof(StatsA), st));
    if (launch_stats_a(one_pass, d_img, g.w, g.h, g.black, g.white, T->d_raw2evf + (MLVB_MAX_BLACK - g.black), D.statsA, ctx->sm_count, st))
        return MLVB_ERR_UNSUPPORTED;
    ctx->launches += 1;
    PinnedLease stage(T);
    if (!stage.p) return MLVB_ERR_CUDA;
    StatsA *hostA = (StatsA *)stage.p;
    if (one_pass) MLVB_CUDA_OK(readback(hostA, D.statsA, sizeof(StatsA), st));
    else MLVB_CUDA_OK(readback(hostA, D.statsA, 16, st));               // ev_sum, ev_num: the first 16 bytes
    MLVB_CUDA_OK(stream_wait(ctx, st));
    const double avg_ev = hostA->ev_sum / (double)hostA->ev_num;          // 0/0 -> NaN -> "not HDR" (hdr.c:435-438)
    if (!(avg_ev > 0.5)) return 0;
    if (one_pass)
        return hdr_interpolate_impl(ctx, d_img, g.w, g.h, g.black, interp_method, use_fullres, use_alias_map, cs_method, d_aux, st, &stage);

    if (focus && focus->nlevels) {
        rc = apply_pixel_list(ctx, *focus, d_img, g, g.npix, 1, 1, 1, st);
        if (rc) return rc;
    }
    if (fix_bad_pixels_mode) {
        {
            std::lock_guard<std::mutex> lk(ctx->clip_mu);
            rc = get_bad_pixel_map(ctx, hdr, g, fix_bad_pixels_mode == 2, d_img, st, &bad);
            if (rc) return rc;
        }
        if (bad && bad->nlevels) {
            rc = apply_pixel_list(ctx, *bad, d_img, g, g.npix, 1, 1, 0, st);
            if (rc) return rc;
        }
    }
    stage.release();                                                      // back to the pool before the next stage leases one
    return run_hdr_interpolate(ctx, d_img, g.w, g.h, g.black, interp_method, use_fullres, use_alias_map, cs_method, d_aux, st);
}

extern "C" {

int hdr_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, off_t offset, size_t max_size)
{
    (void)offset;                                                       // unused by the reference as well
    mlvb_context *ctx = mlvb_default_context();
    if (!ctx) { fprintf(stderr, "libmlvfs_b200: hdr_convert_data: no CUDA context (no CPU path)\n"); return 0; }
    const FrameGeom g = geom_from_headers(frame_headers);
    if (max_size < g.npix * 2) { fprintf(stderr, "libmlvfs_b200: hdr_convert_data needs the whole frame\n"); return 0; }
    Slot *s = acquire_slot(ctx);
    cudaSetDevice(ctx->device);
    int ret = 0;
    if (slot_reserve(*s, 16, g.npix * 2) == MLVB_OK &&
        reserve_device(&s->d_aux, &s->aux_cap, hdr_preview_scratch_bytes((uint16_t)g.white)) == MLVB_OK &&
        cudaMemcpyAsync(s->d_a, image_data, g.npix * 2, cudaMemcpyHostToDevice, s->stream) == cudaSuccess) {
        const int rc = run_hdr_preview(ctx, frame_headers, g, s->d_a, s->d_aux, s->stream);
        if (rc == 1 && cudaMemcpyAsync(image_data, s->d_a, g.npix * 2, cudaMemcpyDeviceToHost, s->stream) == cudaSuccess &&
            cudaStreamSynchronize(s->stream) == cudaSuccess) {
            frame_headers->rawi_hdr.raw_info.black_level *= 4;          // hdr.c:222-223
            frame_headers->rawi_hdr.raw_info.white_level *= 4;
            ret = 1;
        } else {
            cudaStreamSynchronize(s->stream);
        }
    }
    release_slot(ctx, s);
    return ret;
}

int cr2hdr20_convert_data(struct frame_headers *frame_headers, uint16_t *image_data, int interp_method, int fullres,
                          int use_alias_map, int chroma_smooth, int fix_bad_pixels_mode)
{
    mlvb_context *ctx = mlvb_default_context();
    if (!ctx) { fprintf(stderr, "libmlvfs_b200: cr2hdr20_convert_data: no CUDA context (no CPU path)\n"); return 0; }
    const FrameGeom g = geom_from_headers(frame_headers);
    Slot *s = acquire_slot(ctx);
    cudaSetDevice(ctx->device);
    int ret = 0;
    if (slot_reserve(*s, 16, g.npix * 2) == MLVB_OK &&
        reserve_device(&s->d_aux, &s->aux_cap, dual_iso_scratch_bytes(g.w, g.h, interp_method)) == MLVB_OK &&
        cudaMemcpyAsync(s->d_a, image_data, g.npix * 2, cudaMemcpyHostToDevice, s->stream) == cudaSuccess) {
        const int rc = run_cr2hdr20(ctx, frame_headers, g, s->d_a, interp_method, fullres, use_alias_map, chroma_smooth,
                                    fix_bad_pixels_mode, s->d_aux, s->stream);
        // the frame is copied back in both cases: a failed conversion still carries the focus / bad-pixel fixes
        if (rc >= 0 && cudaMemcpyAsync(image_data, s->d_a, g.npix * 2, cudaMemcpyDeviceToHost, s->stream) == cudaSuccess &&
            cudaStreamSynchronize(s->stream) == cudaSuccess && rc == 1) {
            frame_headers->rawi_hdr.raw_info.black_level *= 4;          // hdr.c:1951-1952
            frame_headers->rawi_hdr.raw_info.white_level *= 4;
            ret = 1;
        } else {
            cudaStreamSynchronize(s->stream);
        }
    }
    release_slot(ctx, s);
    return ret;
}

}  // extern "C"
