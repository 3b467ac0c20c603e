// pixfix.cu -- bad-pixel detection and bad/focus-pixel interpolation.
//
// Replaces reference cs.c:87-168 (interpolate_horizontal / _vertical / _pixel), cs.c:257-306 (the
// detection pass of fix_bad_pixels), cs.c:314-330 and cs.c:462-501 (the two apply loops).
//
// The reference applies its list sequentially IN PLACE, so an entry can read pixels that an
// earlier entry already rewrote.  We keep that exact semantics with a level
// schedule: level(m) = 1 + max level of the earlier entries inside m's +-3 cross stencil (0 if none).
// Because the stencil is symmetric, every later entry in m's stencil has a strictly higher level, so
// running levels in order, in place, reads exactly what the sequential loop would have read.
// Level 0 (virtually everything on real sensors) runs grid-wide; deeper levels run in one CTA that
// walks the levels with a barrier in between.
#include <algorithm>
#include <type_traits>

#include "kernels.cuh"
#include "scan.cuh"

namespace {

struct FixLut {
    const int *raw2ev;           // indexed by raw value
    const uint16_t *ev2raw;      // e in [0, 14 EV)
    int black;
};

__device__ __forceinline__ int ev_of(const FixLut &L, const uint16_t *im, long long i) { return __ldg(L.raw2ev + im[i]); }

__device__ __forceinline__ int ev_grad(const FixLut &L, const uint16_t *im, long long i, long long o1, long long o2)
{
    return wabs(wsub(ev_of(L, im, i + o1), ev_of(L, im, i + o2)));
}

// cs.c:87-109 (step 1) / cs.c:111-133 (step w)
__device__ void interp_line(uint16_t *im, long long i, long long step, const FixLut &L)
{
    const int d1 = ev_grad(L, im, i, 3 * step, step);
    const int d2 = ev_grad(L, im, i, -step, -3 * step);
    const int sum = wadd(d1, d2);
    if (sum == 0) { im[i] = im[i + 2 * step]; return; }
    const int c1 = ((sum - d1) << 8) / sum;
    const int c2 = ((sum - d2) << 8) / sum;
    const int ev = (wmul(ev_of(L, im, i + 2 * step), c1) >> 8) + (wmul(ev_of(L, im, i - 2 * step), c2) >> 8);
    im[i] = (uint16_t)(__ldg(L.ev2raw + clamp_ev(ev)) + L.black);
}

// cs.c:135-168
__device__ void interp_cross(uint16_t *im, long long i, long long w, const FixLut &L)
{
    const int dv1 = ev_grad(L, im, i, 3 * w, w);
    const int dv2 = ev_grad(L, im, i, -w, -3 * w);
    const int dh1 = ev_grad(L, im, i, 3, 1);
    const int dh2 = ev_grad(L, im, i, -1, -3);
    const int sum = wadd(wadd(dh1, dh2), wadd(dv1, dv2));
    if (sum == 0) { im[i] = im[i + 2]; return; }
    const int den = wmul(3, sum);
    const int cv1 = ((sum - dv1) << 8) / den;
    const int cv2 = ((sum - dv2) << 8) / den;
    const int ch1 = ((sum - dh1) << 8) / den;
    const int ch2 = ((sum - dh2) << 8) / den;
    const int ev = (wmul(ev_of(L, im, i + 2 * w), cv1) >> 8) + (wmul(ev_of(L, im, i - 2 * w), cv2) >> 8) +
                   (wmul(ev_of(L, im, i + 2), ch1) >> 8) + (wmul(ev_of(L, im, i - 2), ch2) >> 8);
    im[i] = (uint16_t)(__ldg(L.ev2raw + clamp_ev(ev)) + L.black);
}

// One list entry.  edge_rules = 0: bad-pixel semantics (cs.c:314-330, border entries ignored);
// edge_rules = 1: focus-pixel semantics (cs.c:462-501, border entries use 1-D / copy fallbacks).
__device__ void fix_entry(uint16_t *im, int w, int h, int x, int y, int dual_iso, int edge_rules, const FixLut &L)
{
    const long long i = (long long)x + (long long)y * w;
    if (x > 2 && x < w - 3 && y > 2 && y < h - 3) {
        if (dual_iso) interp_line(im, i, 1, L);
        else interp_cross(im, i, w, L);
    } else if (edge_rules && i > 0 && i < (long long)w * h) {
        const bool hedge = (x >= w - 3 && x < w) || (x >= 0 && x <= 3);
        const bool vedge = (y >= h - 3 && y < h) || (y >= 0 && y <= 3);
        if (hedge && !vedge && !dual_iso) interp_line(im, i, w, L);
        else if (vedge && !hedge) interp_line(im, i, 1, L);
        else if (x >= 0 && x <= 3) im[i] = im[i + 2];
        else if (x >= w - 3 && x < w) im[i] = im[i - 2];
    }
}

struct FixArgs {
    uint16_t *img;
    size_t frame_stride;
    int w, h, crop_x, crop_y, dual_iso, edge_rules;
    const PixelXY *list;         // sorted by level, list order preserved inside a level
    FixLut lut;
};

__global__ void fix_level0_kernel(const FixArgs A, unsigned count)
{
    const unsigned m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    const PixelXY p = A.list[m];
    fix_entry(A.img + (size_t)blockIdx.y * A.frame_stride, A.w, A.h, p.x - A.crop_x, p.y - A.crop_y, A.dual_iso,
              A.edge_rules, A.lut);
}

// levels 1..nlevels-1: one CTA per frame, barrier between levels (global writes of a CTA are visible
// to the whole CTA after __syncthreads)
__global__ void fix_deep_levels_kernel(const FixArgs A, const unsigned *__restrict__ level_start, unsigned nlevels)
{
    uint16_t *im = A.img + (size_t)blockIdx.y * A.frame_stride;
    for (unsigned l = 1; l < nlevels; l++) {
        const unsigned lo = level_start[l], hi = level_start[l + 1];
        for (unsigned m = lo + threadIdx.x; m < hi; m += blockDim.x) {
            const PixelXY p = A.list[m];
            fix_entry(im, A.w, A.h, p.x - A.crop_x, p.y - A.crop_y, A.dual_iso, A.edge_rules, A.lut);
        }
        __threadfence_block();
        __syncthreads();
    }
}

// dual-ISO form: every entry uses the horizontal interpolator (or a same-row copy), so a thread applies one
// independent run of a row's entries in list order (PixelList::upload builds the runs).  With --really-bad-pix
// every bright-row pixel of a dual-ISO frame is in the list (its third-largest same-colour neighbour sits in a
// dark row, cs.c:289-304), so a run is a whole row and each entry reads the three values just written before
// it: a serial chain of thousands of steps.  The chain is kept short per step: both EV tables in shared memory
// (persistent blocks), and a 7-pixel window of raw values and their EVs carried in registers, so that a step
// loads only the pixels that enter the window and never waits for its own stores.
constexpr int RUN_THREADS = 256;
constexpr size_t RUN_SMEM = 16384 * sizeof(int) + 32768 * sizeof(uint16_t);

template <bool SMEM_EV2RAW>
__global__ void __launch_bounds__(RUN_THREADS)
fix_row_runs_kernel(const FixArgs A, const unsigned *__restrict__ seg_start, unsigned nseg, int nframes)
{
    extern __shared__ __align__(16) unsigned char run_smem[];
    int *s_r2e = reinterpret_cast<int *>(run_smem);
    uint16_t *s_m13 = reinterpret_cast<uint16_t *>(run_smem + 16384 * sizeof(int));
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s_r2e[i] = __ldg(A.lut.raw2ev + i);
    if (SMEM_EV2RAW) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.lut.ev2raw + 13 * MLVB_EV_RES);
        for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t *>(s_m13)[i] = __ldg(src + i);
    }
    __syncthreads();
    const int w = A.w, h = A.h, black = A.lut.black;
    auto ev_of_raw = [&](int v) { return v < 16384 ? s_r2e[v] : __ldg(A.lut.raw2ev + v); };
    auto raw_of_ev = [&](int e) {
        const int c = clamp_ev(e);
        return SMEM_EV2RAW ? (int)(s_m13[c & (MLVB_EV_RES - 1)] >> (13 - (c >> 15))) : (int)__ldg(A.lut.ev2raw + c);
    };
    const unsigned total = nseg * (unsigned)nframes;
    for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const unsigned sidx = t % nseg, frame = t / nseg;
        uint16_t *im = A.img + (size_t)frame * A.frame_stride;
        int R[7], E[7];                    // raw values and EVs at x-3 .. x+3 of the previous entry (same row)
        int lastx = -100, lasty = -1;
        for (unsigned m = seg_start[2 * sidx], hi = seg_start[2 * sidx + 1]; m < hi; m++) {
            const PixelXY p = A.list[m];
            const int x = p.x - A.crop_x, y = p.y - A.crop_y;
            if (!(x > 2 && x < w - 3 && y > 2 && y < h - 3)) {                  // border rules: generic path
                fix_entry(im, w, h, x, y, 1, A.edge_rules, A.lut);
                lastx = -100;
                continue;
            }
            uint16_t *row = im + (size_t)y * w;
            const int dx = x - lastx;
            if (y == lasty && dx >= 1 && dx <= 3) {
                for (int sft = 0; sft < dx; sft++) {
#pragma unroll
                    for (int k = 0; k < 6; k++) { R[k] = R[k + 1]; E[k] = E[k + 1]; }
                    R[6] = row[lastx + sft + 4];
                    E[6] = ev_of_raw(R[6]);
                }
            } else if (!(y == lasty && dx == 0)) {
#pragma unroll
                for (int k = 0; k < 7; k++) { R[k] = row[x - 3 + k]; E[k] = ev_of_raw(R[k]); }
            }
            lastx = x; lasty = y;
            // cs.c:87-109 (interpolate_horizontal), same expressions as interp_line
            const int d1 = wabs(wsub(E[6], E[4])), d2 = wabs(wsub(E[2], E[0]));
            const int sum = wadd(d1, d2);
            int v;
            if (sum == 0) v = R[5];
            else {
                const int c1 = ((sum - d1) << 8) / sum, c2 = ((sum - d2) << 8) / sum;
                const int ev = (wmul(E[5], c1) >> 8) + (wmul(E[1], c2) >> 8);
                v = (raw_of_ev(ev) + black) & 0xFFFF;
            }
            row[x] = (uint16_t)v;
            R[3] = v;
            E[3] = ev_of_raw(v);
        }
    }
}

// Rows with many entries: one warp stages the whole image row in shared memory and applies the row's entries there
// (every dual-ISO-mode rule of fix_entry stays inside the row: horizontal interpolation or a copy from x +- 2) --
// independent runs of entries on different lanes, each run in list order -- then writes the row back.  The chains
// never wait for global memory.
constexpr int LONG_WARPS_MAX = 16;
constexpr int LONG_CHUNK = 256;            // list entries staged per warp at a time
constexpr int LONG_WIN = 4096;             // pixels of a row staged per warp at a time
#ifndef LONG_WARM_CFG
#define LONG_WARM_CFG 128
#endif
constexpr int LONG_WARM = LONG_WARM_CFG;    // warm-up steps of a speculative piece of a dense run (walk_dense_warp)
constexpr int LONG_DENSE_MIN = 192;        // shortest stretch of consecutive entries handed to walk_dense_warp
constexpr size_t LONG_WARP_BYTES = (LONG_WIN + 8 + LONG_CHUNK) * sizeof(uint16_t);   // window, entry columns

template <bool SMEM_EV2RAW>
__global__ void __launch_bounds__(LONG_WARPS_MAX * 32)
fix_long_rows_kernel(const FixArgs A, const unsigned *__restrict__ long_rows, unsigned nlong, int nframes, int warps_per_block)
{
    extern __shared__ __align__(16) unsigned char run_smem[];
    int *s_r2e = reinterpret_cast<int *>(run_smem);
    // with SMEM_EV2RAW the exp table's top octave follows (64 KiB), else the row buffers start right here and the
    // exp table is read through L1: more rows per SM, so that every long row of a frame is walked concurrently
    uint16_t *s_m13 = reinterpret_cast<uint16_t *>(run_smem + 16384 * sizeof(int));
    uint16_t *s_rows = reinterpret_cast<uint16_t *>(run_smem + (SMEM_EV2RAW ? RUN_SMEM : 16384 * sizeof(int)));
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s_r2e[i] = __ldg(A.lut.raw2ev + i);
    if (SMEM_EV2RAW) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(A.lut.ev2raw + 13 * MLVB_EV_RES);
        for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t *>(s_m13)[i] = __ldg(src + i);
    }
    __syncthreads();
    const int w = A.w, h = A.h, black = A.lut.black;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_block) return;
    // per warp: a window of LONG_WIN pixels of the row being repaired (+ 8 spare) and LONG_CHUNK staged entry columns
    uint16_t *win = s_rows + (size_t)warp * (LONG_WARP_BYTES / sizeof(uint16_t));
    uint16_t *xs = win + LONG_WIN + 8;
    uint16_t *row = win;                                                        // row[x] = win[x - wx0], set per window
    auto ev_of_raw = [&](int v) { return v < 16384 ? s_r2e[v] : __ldg(A.lut.raw2ev + v); };
    auto raw_of_ev = [&](int e) {
        const int c = clamp_ev(e);
        return SMEM_EV2RAW ? (int)(s_m13[c & (MLVB_EV_RES - 1)] >> (13 - (c >> 15))) : (int)__ldg(A.lut.ev2raw + c);
    };
    // (num << 8) / sum for 0 <= num <= sum < 2^22 (quotient 0..256): float estimate + exact correction, ~10
    // dependent instructions instead of the ~35 of a 32-bit integer division
    auto div8 = [&](int num, int sum) {
        const int n8 = num << 8;
        int q = __float2int_rz(__fdividef((float)n8, (float)sum));
        const int r = n8 - q * sum;
        q += (r >= sum) - (r < 0);
        return q;
    };
    // window of EVs at x-3 .. x+3 around the previous entry (valid while entries advance by 1..3 columns)
    int E[7];
    int lastx = -100;
    auto interp_h = [&](int x) {                                                // cs.c:87-109, as interp_line(step 1)
        const int dx = x - lastx;
        if (dx >= 1 && dx <= 3) {
            for (int sft = 0; sft < dx; sft++) {
#pragma unroll
                for (int k = 0; k < 6; k++) E[k] = E[k + 1];
                E[6] = ev_of_raw(row[lastx + sft + 4]);
            }
        } else if (dx != 0) {
#pragma unroll
            for (int k = 0; k < 7; k++) E[k] = ev_of_raw(row[x - 3 + k]);
        }
        lastx = x;
        const int d1 = wabs(wsub(E[6], E[4])), d2 = wabs(wsub(E[2], E[0]));
        const int sum = wadd(d1, d2);
        int v;
        if (sum == 0) v = row[x + 2];
        else {
            int c1, c2;
            if (d1 >= 0 && d2 >= 0 && sum > 0 && sum < (1 << 22)) { c1 = div8(sum - d1, sum); c2 = div8(sum - d2, sum); }
            else { c1 = ((sum - d1) << 8) / sum; c2 = ((sum - d2) << 8) / sum; }   // wrap-around cases (pixels at black)
            const int ev = (wmul(E[5], c1) >> 8) + (wmul(E[1], c2) >> 8);
            v = (raw_of_ev(ev) + black) & 0xFFFF;
        }
        row[x] = (uint16_t)v;
        E[3] = ev_of_raw(v);
    };
    // ---- dense stretches: `cnt` entries at consecutive interior columns x0, x0 + 1, .. -- what --really-bad-pix makes
    // of every bright row of a dual-ISO frame (cs.c:289-304 flags each of its pixels): one recurrence thousands of steps
    // long.  One step (cs.c:87-109, the arithmetic of interp_h) is a function of the EVs of the three repaired pixels
    // to the left (e0, e1, e2), the EVs of the three originals to the right (o1, o2, o3) and the original at x + 2.
    // A single warp issues the whole step as one dependent chain, so the step is written for few instructions: tables
    // and window through 32-bit shared addresses, one reciprocal for both weights, rare cases behind branches.
    const uint32_t sa_r2e = (uint32_t)__cvta_generic_to_shared(s_r2e), sa_m13 = (uint32_t)__cvta_generic_to_shared(s_m13);
    auto lds_s32 = [](uint32_t a) { int v; asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; };
    auto lds_u16 = [](uint32_t a) { unsigned short v; asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (int)v; };
    auto win_u16 = [](uint32_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return (int)v; };
    auto ev_fast = [&](int v) {
        if (__builtin_expect(v >= 16384, 0)) return (int)__ldg(A.lut.raw2ev + v);
        return lds_s32(sa_r2e + 4u * (unsigned)v);
    };
    auto raw_fast = [&](int e) {
        const int c = clamp_ev(e);
        if (!SMEM_EV2RAW) return (int)__ldg(A.lut.ev2raw + c);
        return lds_u16(sa_m13 + 2u * (unsigned)(c & (MLVB_EV_RES - 1))) >> (13 - (c >> 15));
    };
    // v = the repaired sample at x; sa_x = shared address of the window's pixel x
    auto dense_step = [&](uint32_t sa_x, int e0, int e1, int e2, int o1, int o2, int o3) {
        const int d1 = wabs(wsub(o3, o1)), d2 = wabs(wsub(e2, e0));
        const int sum = wadd(d1, d2);
        if (__builtin_expect(sum == 0, 0)) return win_u16(sa_x + 4);             // row[x + 2] (cs.c:96-99)
        int c1, c2;
        if (__builtin_expect((unsigned)(d1 | d2) < (1u << 21), 1)) {
            // 0 <= d1, d2 < 2^21: one division gives both weights.  c1 = floor(256 d2 / sum) from a float estimate
            // (within 1 of the quotient: 24-bit operands, approximate reciprocal) corrected by its remainder;
            // d1 + d2 = sum, so c2 = floor(256 d1 / sum) = 256 - ceil(256 d2 / sum)
            const int n8 = d2 << 8;
            float rs;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(__int2float_rn(sum)));
            int q = __float2int_rz(__int2float_rn(n8) * rs);
            int r = n8 - q * sum;
            if (r >= sum) { q++; r -= sum; } else if (r < 0) { q--; r += sum; }
            c1 = q;
            c2 = 256 - q - (r != 0);
        } else { c1 = ((sum - d1) << 8) / sum; c2 = ((sum - d2) << 8) / sum; }   // wrap-around cases (pixels at black)
        const int ev = (wmul(o2, c1) >> 8) + (wmul(e1, c2) >> 8);
        return (raw_fast(ev) + black) & 0xFFFF;
    };
    // one lane, in the window (the generic path's in-place semantics: results replace the window's pixels)
    auto walk_dense = [&](int x0, int cnt) {
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(row + x0);
        int e0 = ev_fast(win_u16(sa - 6)), e1 = ev_fast(win_u16(sa - 4)), e2 = ev_fast(win_u16(sa - 2));
        int o1 = ev_fast(win_u16(sa + 2)), o2 = ev_fast(win_u16(sa + 4)), o3 = ev_fast(win_u16(sa + 6));
        for (int x = x0; x < x0 + cnt; x++, sa += 2) {
            const int rn = x + 1 < x0 + cnt ? win_u16(sa + 8) : 0;              // the next original, needed one step later
            const int v = dense_step(sa, e0, e1, e2, o1, o2, o3);
            row[x] = (uint16_t)v;
            e0 = e1; e1 = e2; e2 = ev_fast(v);
            o1 = o2; o2 = o3; o3 = ev_fast(rn);
        }
        lastx = -100;                                                           // the next entry rebuilds its EV window from the row
    };
    // The same stretch walked by the whole warp.  The recurrence forgets: a repaired pixel is a blend of its two
    // neighbours at x - 2 / x + 2 rounded to a sample value, so a walk started some dozens of columns early from the
    // wrong state (the unrepaired pixels) runs into the true trajectory and stays on it (median ~45 columns on the
    // synthetic dual-ISO frames, > 128 for ~5 % of the starts).  Lane k therefore walks piece k of the stretch
    // speculatively after LONG_WARM warm-up steps, all lanes at once; results go straight to the frame in global
    // memory `grow`, the staged window keeps the originals every lane reads to its right.  Then the pieces are
    // checked in order: piece k is the reference's result iff the state it assumed at its first column (the EVs of the
    // three samples to the left) equals the state piece k - 1 ended in; a piece that fails is walked again from the
    // true state by its lane until it meets its own earlier results three columns in a row (same state: the rest of
    // the piece stands), and the check moves on.  Exact by construction; the speculation only decides how much is
    // redone -- in the worst case the stretch is walked serially once more.
    struct EvState { int e0, e1, e2; };
    auto spec_piece = [&](uint16_t *grow, int xb, int p0, int p1, EvState &at_p0, EvState &at_end) {
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(row + xb);
        int e0 = ev_fast(win_u16(sa - 6)), e1 = ev_fast(win_u16(sa - 4)), e2 = ev_fast(win_u16(sa - 2));
        int o1 = ev_fast(win_u16(sa + 2)), o2 = ev_fast(win_u16(sa + 4)), o3 = ev_fast(win_u16(sa + 6));
        for (int x = xb; x < p0; x++, sa += 2) {                                // warm-up: nothing is stored
            const int rn = win_u16(sa + 8);
            const int v = dense_step(sa, e0, e1, e2, o1, o2, o3);
            e0 = e1; e1 = e2; e2 = ev_fast(v);
            o1 = o2; o2 = o3; o3 = ev_fast(rn);
        }
        at_p0 = EvState{e0, e1, e2};
        uint16_t *gp = grow + p0;
        for (int x = p0; x < p1; x++, sa += 2, gp++) {
            const int rn = win_u16(sa + 8);                                     // past the stretch: the zeroed spare of the window
            const int v = dense_step(sa, e0, e1, e2, o1, o2, o3);
            *gp = (uint16_t)v;
            e0 = e1; e1 = e2; e2 = ev_fast(v);
            o1 = o2; o2 = o3; o3 = ev_fast(rn);
        }
        at_end = EvState{e0, e1, e2};
    };
    auto redo_piece = [&](uint16_t *grow, int p0, int p1, EvState st, EvState &at_end) {
        uint32_t sa = (uint32_t)__cvta_generic_to_shared(row + p0);
        int e0 = st.e0, e1 = st.e1, e2 = st.e2;
        int o1 = ev_fast(win_u16(sa + 2)), o2 = ev_fast(win_u16(sa + 4)), o3 = ev_fast(win_u16(sa + 6));
        uint16_t *gp = grow + p0;
        int same = 0, prev = *gp;                                               // this lane's earlier result at x
        for (int x = p0; x < p1; x++, sa += 2, gp++) {
            const int rn = win_u16(sa + 8);
            const int pn = x + 1 < p1 ? (int)gp[1] : 0;
            const int v = dense_step(sa, e0, e1, e2, o1, o2, o3);
            same = v == prev ? same + 1 : 0;
            if (same >= 3) return;                                              // back on the earlier trajectory: the end state stands
            if (v != prev) *gp = (uint16_t)v;
            prev = pn;
            e0 = e1; e1 = e2; e2 = ev_fast(v);
            o1 = o2; o2 = o3; o3 = ev_fast(rn);
        }
        at_end = EvState{e0, e1, e2};
    };
    auto walk_dense_warp = [&](uint16_t *grow, int x0, int cnt) {               // cnt <= LONG_WIN, whole warp
        const int L = (cnt + 31) >> 5;
        const int p0 = x0 + lane * L, p1 = min(p0 + L, x0 + cnt);
        const bool has = p0 < x0 + cnt;
        const int xb = max(x0, p0 - LONG_WARM);
        EvState a = {0, 0, 0}, b = {0, 0, 0};
        if (has) spec_piece(grow, xb, p0, p1, a, b);
        __syncwarp();
        // piece k is checked against piece k - 1; pieces that started at x0 itself started from the true state
        auto from_lane = [&](const EvState &s, int src) {
            return EvState{__shfl_sync(0xFFFFFFFFu, s.e0, src), __shfl_sync(0xFFFFFFFFu, s.e1, src), __shfl_sync(0xFFFFFFFFu, s.e2, src)};
        };
        auto differs = [](const EvState &u, const EvState &v) { return u.e0 != v.e0 || u.e1 != v.e1 || u.e2 != v.e2; };
        EvState pb = from_lane(b, max(lane - 1, 0));
        bool wrong = has && xb > x0 && differs(a, pb);
        unsigned bad = __ballot_sync(0xFFFFFFFFu, wrong);
        while (bad) {
            const int j = __ffs(bad) - 1;                                       // j >= 1: lane 0 starts at x0
            pb = from_lane(b, j - 1);
            if (lane == j) {
                redo_piece(grow, p0, p1, pb, b);
                wrong = false;
            }
            __syncwarp();
            pb = from_lane(b, j);
            if (lane == j + 1) wrong = has && differs(a, pb);
            bad = __ballot_sync(0xFFFFFFFFu, wrong);
        }
        __syncwarp();
        lastx = -100;
    };
    const unsigned total = nlong * (unsigned)nframes;
    for (unsigned t = (unsigned)warp * gridDim.x + blockIdx.x; t < total; t += gridDim.x * warps_per_block) {   // rows spread evenly over the blocks
        const unsigned ridx = t % nlong, frame = t / nlong;
        const unsigned m0 = long_rows[2 * ridx], m1 = long_rows[2 * ridx + 1];
        const int y = A.list[m0].y - A.crop_y;
        if (y <= 3 || y >= h - 3) {
            // top / bottom edge rows (and rows outside the frame): the reference's edge rules address pixels by
            // linear index there (cs.c:467, 479-500), which can leave the row; apply them in place, entry by entry
            if (lane == 0)
                for (unsigned m = m0; m < m1; m++)
                    fix_entry(A.img + (size_t)frame * A.frame_stride, w, h, A.list[m].x - A.crop_x, y, 1, A.edge_rules, A.lut);
            __syncwarp();
            continue;
        }
        uint16_t *grow = A.img + (size_t)frame * A.frame_stride + (size_t)y * w;
        auto one_entry = [&](int x) {                                           // rows 4 .. h-4: fix_entry with dual_iso = 1
            if (x > 2 && x < w - 3) interp_h(x);
            else if (A.edge_rules && x >= 0 && x < w) {
                if (x <= 3) row[x] = row[x + 2];
                else row[x] = row[x - 2];
                lastx = -100;
            }
        };
        // ascending rows (a detected bad-pixel list is in raster order) are repaired window by window in shared memory;
        // anything else (a focus-pixel map in file order) by one lane directly on the frame
        bool ascending, strict;                                               // strict: no column listed twice
        {
            bool bad = false, dup = false;                                    // one vote after the loop: the loads pipeline
            for (unsigned mb = m0; mb < m1; mb += 32) {
                const unsigned m = mb + lane;
                if (m + 1 < m1) {
                    const int xa = A.list[m].x, xb = A.list[m + 1].x;
                    bad |= xb < xa;
                    dup |= xb == xa;
                }
            }
            ascending = !__any_sync(0xFFFFFFFFu, bad);
            strict = !__any_sync(0xFFFFFFFFu, dup);
        }
        if (!ascending) {
            if (lane == 0) {
                row = grow;
                lastx = -100;
                for (unsigned m = m0; m < m1; m++) one_entry(A.list[m].x - A.crop_x);
            }
            __syncwarp();
            continue;
        }
        // Windows of LONG_WIN pixels slide along the row: a window starts 3 pixels left of the first entry that is still
        // to do and takes every entry whose stencil (x-3 .. x+4) lies inside it.  The entries' columns are staged
        // LONG_CHUNK at a time (coalesced) so that no walk waits for the list.  Entries more than 3 columns apart cannot
        // see each other (an entry reads x-3 .. x+3 and writes x), so a chunk splits into independent runs: the lane
        // that holds the first entry of a run walks that run in list order, all runs of the chunk at once; a run that
        // continues from the previous chunk or window is picked up by lane 0 (the EV window is only a cache of the
        // staged pixels).
        // A stretch of at least LONG_DENSE_MIN entries at consecutive interior columns is walked by the whole warp
        // (walk_dense_warp).  Columns are strictly increasing, so entries m .. m + k - 1 are consecutive iff the last
        // one sits k - 1 columns right of the first: the longest such stretch inside the window by bisection.
        auto dense_stretch = [&](unsigned m, int xfirst, int xlim) {
            if (!strict || xfirst <= 2) return 0;
            int K = 1, khi = min((int)(m1 - m), min(xlim, w - 4) - xfirst + 1);
            if (khi < LONG_DENSE_MIN) return 0;
            // invariant: a stretch of K holds, one of khi + 1 does not; every lane probes one length per round
            while (K < khi) {
                const int span = khi - K, step = (span + 31) >> 5;              // probes K + step, K + 2 step, ... (<= khi)
                const int probe = min(K + (lane + 1) * step, khi);
                const bool ok = A.list[m + probe - 1].x - A.crop_x == xfirst + probe - 1;
                const unsigned okm = __ballot_sync(0xFFFFFFFFu, ok);
                // consecutive-ness is monotone in the length: the lanes that succeed are a prefix
                const int nok = __popc(okm);
                const int newK = nok ? min(K + nok * step, khi) : K;
                const int newhi = nok == 32 ? khi : min(K + (nok + 1) * step, khi) - 1;
                K = newK; khi = max(newhi, K);
            }
            return K;
        };
        const bool vec_ok = (((uintptr_t)grow) & 15) == 0;                      // eight pixels as one 128-bit word
        auto stage = [&](int wx0, int wlen) {
            if (vec_ok && !(wx0 & 7)) {
                const uint4 *src = reinterpret_cast<const uint4 *>(grow + wx0);
                uint4 *dst = reinterpret_cast<uint4 *>(win);
#pragma unroll 4
                for (int i = lane; i < (wlen >> 3); i += 32) dst[i] = src[i];
                for (int i = (wlen & ~7) + lane; i < wlen; i += 32) win[i] = grow[wx0 + i];
            } else
                for (int i = lane; i < wlen; i += 32) win[i] = grow[wx0 + i];
            if (lane < 8) win[wlen + lane] = 0;                                 // the spare a walk may read ahead into
        };
        unsigned m = m0;
        while (m < m1) {
            const int xraw = A.list[m].x - A.crop_x;
            const int xf = min(max(xraw, 0), w - 1);
            const int wx0 = max((xf - 3) & ~7, 0);                              // 8-pixel aligned for the 128-bit staging
            const int wmax = min(LONG_WIN, w - wx0);
            const int xlim = wx0 + wmax >= w ? 0x3FFFFFFF : wx0 + wmax - 5;     // last column whose stencil fits
            const int K = dense_stretch(m, xraw, xlim);
            if (K >= LONG_DENSE_MIN) {
                stage(wx0, min(wmax, xraw + K + 5 - wx0));                      // the stretch's stencils: x - 3 .. x + 4
                row = win - wx0;
                __syncwarp();
                walk_dense_warp(grow, xraw, K);                                 // writes the frame itself; the window is dropped
                m += (unsigned)K;
                continue;
            }
            const int wlen = wmax;
            stage(wx0, wlen);
            row = win - wx0;
            __syncwarp();
            bool window_full = false;
            while (m < m1 && !window_full) {
                const int n = (int)min((unsigned)LONG_CHUNK, m1 - m);
                int nin = 0;                                                    // staged entries that fit this window (a prefix)
                for (int i = lane; i < n; i += 32) {
                    const int x = A.list[m + i].x - A.crop_x;
                    xs[i] = (uint16_t)(min(max(x, -1), 0xFFFE) + 1);            // biased by 1: 0 = left of the frame
                    nin += x <= xlim;
                }
                for (int o = 16; o; o >>= 1) nin += __shfl_xor_sync(0xFFFFFFFFu, nin, o);
                __syncwarp();
                for (int i = lane; i < nin; i += 32) {
                    const int x = (int)xs[i] - 1;
                    if (i == 0 || x - ((int)xs[i - 1] - 1) > 3) {               // first entry of a run (or of the chunk)
                        lastx = -100;                                           // (re)build the EV window from the staged pixels
                        int j = i, xx = x;
                        while (true) {
                            // strictly increasing columns: entries j .. j + len - 1 are consecutive iff the last one is
                            // len - 1 columns to the right of the first -- the longest such stretch by halving
                            int len = 1;
                            if (strict && xx > 2) {
                                len = nin - j;
                                while (len > 1 && ((int)xs[j + len - 1] - 1 - xx != len - 1 || xx + len - 1 >= w - 3)) len >>= 1;
                            }
                            if (len >= 8) { walk_dense(xx, len); j += len - 1; xx += len - 1; }
                            else one_entry(xx);
                            if (++j >= nin) break;
                            const int xnext = (int)xs[j] - 1;
                            if (xnext - xx > 3) break;
                            xx = xnext;
                        }
                    }
                }
                __syncwarp();
                m += (unsigned)nin;
                if (nin == 0) m++;                                              // cannot happen (a window takes its first entry); never spin
                window_full = nin < n;
                // a long stretch begins here: hand it to the warp-wide walk (new window)
                if (m < m1 && !window_full && dense_stretch(m, A.list[m].x - A.crop_x, 0x3FFFFFFF) >= LONG_DENSE_MIN) break;
            }
            for (int i = lane; i < wlen; i += 32) grow[wx0 + i] = win[i];
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------- detection (cs.c:257-306) ----

constexpr int DET_THREADS = 256;
constexpr int DET_PX = 8;       // consecutive pixels per thread (one flag byte)

__device__ __forceinline__ bool is_bad(const uint16_t *im, int w, int x, int y, const int *raw2ev, int black,
                                       int aggressive)
{
    const int p = im[x + (size_t)y * w];
    // three largest (with multiplicity) of the eight same-colour neighbours at +-2
    int m1 = -1, m2 = -1, m3 = -1;
#pragma unroll
    for (int dy = -2; dy <= 2; dy += 2)
#pragma unroll
        for (int dx = -2; dx <= 2; dx += 2) {
            if (dx == 0 && dy == 0) continue;
            const int q = im[(x + dx) + (size_t)(y + dy) * w];
            if (q >= m1) { m3 = m2; m2 = m1; m1 = q; }
            else if (q >= m2) { m3 = m2; m2 = q; }
            else if (q > m3) m3 = q;
        }
    const int dark_min = black - 96, dark_max = black + 96;          // dark_noise 12 * 8 (cs.c:257-259)
    if (p < dark_min) return true;
    const int ep = __ldg(raw2ev + p);
    if (wsub(ep, __ldg(raw2ev + m2)) > 2 * MLVB_EV_RES && p > dark_max) return true;
    if (aggressive && (wsub(ep, __ldg(raw2ev + m2)) > MLVB_EV_RES || wsub(ep, __ldg(raw2ev + m3)) > MLVB_EV_RES) &&
        p > dark_max)
        return true;
    return false;
}

__global__ void __launch_bounds__(DET_THREADS)
badpix_flag_kernel(const uint16_t *__restrict__ im, int w, int h, const int *__restrict__ raw2ev, int black,
                   int aggressive, uint8_t *__restrict__ flags, unsigned long long *__restrict__ cta_counts)
{
    const size_t npix = (size_t)w * h;
    const size_t t = (size_t)blockIdx.x * DET_THREADS + threadIdx.x;
    unsigned bits = 0;
#pragma unroll
    for (int k = 0; k < DET_PX; k++) {
        const size_t i = t * DET_PX + k;
        if (i < npix) {
            const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            if (x >= 6 && x < w - 6 && y >= 6 && y < h - 6 && is_bad(im, w, x, y, raw2ev, black, aggressive)) bits |= 1u << k;
        }
    }
    if (t * DET_PX < npix) flags[t] = (uint8_t)bits;
    unsigned total;
    block_exclusive_scan(__popc(bits), total);
    if (threadIdx.x == 0) cta_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(DET_THREADS)
badpix_scatter_kernel(const uint8_t *__restrict__ flags, int w, size_t npix, int crop_x, int crop_y,
                      const unsigned long long *__restrict__ cta_offsets, PixelXY *__restrict__ list)
{
    const size_t t = (size_t)blockIdx.x * DET_THREADS + threadIdx.x;
    const unsigned bits = (t * DET_PX < npix) ? flags[t] : 0u;
    unsigned total;
    unsigned pos = block_exclusive_scan(__popc(bits), total);
    if (!bits) return;
    size_t o = cta_offsets[blockIdx.x] + pos;
    for (int k = 0; k < DET_PX; k++)
        if (bits & (1u << k)) {
            const size_t i = t * DET_PX + k;
            const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            list[o++] = PixelXY{x + crop_x, y + crop_y};
        }
}

}  // namespace

int badpix_detect_scratch_bytes(int w, int h, size_t *flag_bytes, size_t *count_bytes)
{
    const size_t npix = (size_t)w * h;
    const size_t nthreads = (npix + DET_PX - 1) / DET_PX;
    const size_t nctas = (nthreads + DET_THREADS - 1) / DET_THREADS;
    *flag_bytes = (nthreads + 255) & ~(size_t)255;
    *count_bytes = (nctas + 1) * sizeof(unsigned long long);
    return (int)nctas;
}

// Phase 1: flags + per-CTA counts + scan.  d_counts[nctas] receives the total (read it back, allocate
// the list, then call phase 2).
int launch_badpix_detect_count(const uint16_t *d_img, int w, int h, int black, int aggressive, const EvLuts &luts,
                               uint8_t *d_flags, unsigned long long *d_counts, cudaStream_t st)
{
    if (black > MLVB_MAX_BLACK) return MLVB_ERR_ARG;
    size_t fb, cb;
    const int nctas = badpix_detect_scratch_bytes(w, h, &fb, &cb);
    badpix_flag_kernel<<<nctas, DET_THREADS, 0, st>>>(d_img, w, h, luts.raw2ev_base + (MLVB_MAX_BLACK - black), black,
                                                      aggressive, d_flags, d_counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(d_counts, (unsigned)nctas, d_counts + nctas);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_badpix_detect_scatter(const uint8_t *d_flags, const unsigned long long *d_offsets, int w, int h, int crop_x,
                                 int crop_y, PixelXY *d_list, cudaStream_t st)
{
    size_t fb, cb;
    const int nctas = badpix_detect_scratch_bytes(w, h, &fb, &cb);
    badpix_scatter_kernel<<<nctas, DET_THREADS, 0, st>>>(d_flags, w, (size_t)w * h, crop_x, crop_y, d_offsets, d_list);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_pixel_fix(uint16_t *d_img, int w, int h, size_t frame_stride, int nframes, int black, int crop_x, int crop_y,
                     int dual_iso, int edge_rules, const PixelXY *d_list_by_level, const unsigned *d_level_start,
                     const unsigned *h_level_start, unsigned nlevels, const EvLuts &luts, cudaStream_t st)
{
    if (nlevels == 0 || h_level_start[nlevels] == 0) return MLVB_OK;
    if (black > MLVB_MAX_BLACK) return MLVB_ERR_ARG;
    FixArgs A;
    A.img = d_img; A.frame_stride = frame_stride; A.w = w; A.h = h; A.crop_x = crop_x; A.crop_y = crop_y;
    A.dual_iso = dual_iso; A.edge_rules = edge_rules; A.list = d_list_by_level;
    A.lut.raw2ev = luts.raw2ev_base + (MLVB_MAX_BLACK - black);
    A.lut.ev2raw = luts.ev2raw_pos;
    A.lut.black = black;
    const unsigned n0 = h_level_start[1];
    if (n0) fix_level0_kernel<<<dim3(ceil_div(n0, 128), nframes), 128, 0, st>>>(A, n0);
    if (nlevels > 1) fix_deep_levels_kernel<<<dim3(1, nframes), 1024, 0, st>>>(A, d_level_start, nlevels);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_pixel_fix_rows(uint16_t *d_img, int w, int h, size_t frame_stride, int nframes, int black, int crop_x, int crop_y,
                          int edge_rules, const PixelXY *d_list_by_row, const unsigned *d_seg_start, unsigned nseg,
                          const unsigned *d_long_rows, unsigned nlong, const EvLuts &luts, int ev2raw_octaves_ok, int sm_count,
                          cudaStream_t st)
{
    if (nseg == 0 && nlong == 0) return MLVB_OK;
    if (black > MLVB_MAX_BLACK) return MLVB_ERR_ARG;
    FixArgs A;
    A.img = d_img; A.frame_stride = frame_stride; A.w = w; A.h = h; A.crop_x = crop_x; A.crop_y = crop_y;
    A.dual_iso = 1; A.edge_rules = edge_rules; A.list = d_list_by_row;
    A.lut.raw2ev = luts.raw2ev_base + (MLVB_MAX_BLACK - black);
    A.lut.ev2raw = luts.ev2raw_pos;
    A.lut.black = black;
    const int sms = sm_count > 0 ? sm_count : 148;
    // long rows first?  No: rows are independent of each other in this mode, any order is the reference's result
    const size_t row_bytes = (size_t)((w + 7) & ~7) * sizeof(uint16_t);
    // rows per SM with both tables in shared memory; if that cannot hold all long rows of the batch at once, keep
    // only the log table there (the exp table is then read through L1) to avoid a second round of serial chains
    (void)row_bytes;
    const size_t warp_bytes = LONG_WARP_BYTES;                                    // staged pixel window + staged entry columns + results
    int long_warps = (int)std::min<size_t>(LONG_WARPS_MAX, (227 * 1024 - RUN_SMEM - 1024) / warp_bytes);
    bool long_ev2raw_smem = ev2raw_octaves_ok != 0;
    const char *force_smem = getenv("MLVB_LONG_SMEM");
    if (!(force_smem && *force_smem == '1' && long_warps >= 1) &&
        ((long long)nlong * nframes > (long long)long_warps * sms || long_warps < 1)) {
        const int alt = (int)std::min<size_t>(LONG_WARPS_MAX, (227 * 1024 - 16384 * sizeof(int) - 1024) / warp_bytes);
        if (alt > long_warps) { long_warps = alt; long_ev2raw_smem = false; }
    }
    const unsigned *segs = d_seg_start;
    unsigned nshort = nseg;
    if (nlong && long_warps < 1) {                                            // rows too wide to stage: thread-per-row walk
        segs = d_long_rows; nshort = nlong;
        if (nseg) {
            const int b0 = (int)std::min<long long>(sms, ceil_div((long long)nseg * nframes, RUN_THREADS));
            MLVB_CUDA_OK(cudaFuncSetAttribute(fix_row_runs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM));
            fix_row_runs_kernel<false><<<b0, RUN_THREADS, RUN_SMEM, st>>>(A, d_seg_start, nseg, nframes);
        }
    } else if (nlong) {
        const size_t smem = (long_ev2raw_smem ? RUN_SMEM : 16384 * sizeof(int)) + (size_t)long_warps * warp_bytes;
        const int blocks = (int)std::min<long long>(sms, (long long)nlong * nframes);
        if (long_ev2raw_smem) {
            MLVB_CUDA_OK(cudaFuncSetAttribute(fix_long_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fix_long_rows_kernel<true><<<blocks, LONG_WARPS_MAX * 32, smem, st>>>(A, d_long_rows, nlong, nframes, long_warps);
        } else {
            MLVB_CUDA_OK(cudaFuncSetAttribute(fix_long_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fix_long_rows_kernel<false><<<blocks, LONG_WARPS_MAX * 32, smem, st>>>(A, d_long_rows, nlong, nframes, long_warps);
        }
    }
    if (nshort) {
        const int blocks = (int)std::min<long long>(sms, ceil_div((long long)nshort * nframes, RUN_THREADS));
        if (ev2raw_octaves_ok) {
            MLVB_CUDA_OK(cudaFuncSetAttribute(fix_row_runs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM));
            fix_row_runs_kernel<true><<<blocks, RUN_THREADS, RUN_SMEM, st>>>(A, segs, nshort, nframes);
        } else {
            MLVB_CUDA_OK(cudaFuncSetAttribute(fix_row_runs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RUN_SMEM));
            fix_row_runs_kernel<false><<<blocks, RUN_THREADS, RUN_SMEM, st>>>(A, segs, nshort, nframes);
        }
    }
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
