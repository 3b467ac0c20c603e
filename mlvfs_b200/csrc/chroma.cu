// chroma.cu -- 2x2 / 3x3 / 5x5 median chroma smoothing in EV space, with the vertical-stripe gain
// fused in as a store epilogue.
//
// Replaces reference cs.c:49-84 + the chroma_smooth.c:22-71 template (uint16 instance) and its
// uint32 twin in hdr.c:1488-1522, and stripes.c:250-266 (stripes_apply_correction) when fused.
//
// Work unit: an RGGB quad.  For every quad of the tile plus a halo ring the kernel computes, once,
//     ge = (ev(g1) + ev(g2)) / 2,   dr = ev(r) - ge,   db = ev(b) - ge
// into shared memory; every interior quad then takes the median of dr / db over its 5 / 9 / 25
// neighbouring quads and rewrites its R and B sites as ev2raw[ge + median] + black.  G sites and
// every skipped quad pass through unchanged (the reference smooths from an untouched copy, so the
// kernel is out of place: in -> out).  All arithmetic is the reference's wrapping int32 math.
#include "kernels.cuh"
#include "median_networks.cuh"

namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_TQX = 64;   // tile width in quads (128 px)
constexpr int CS_TQY = 8;    // tile height in quads (16 px)

template <int METHOD> struct CsGeom {
    static constexpr int R = (METHOD == 5) ? 2 : 1;                    // halo in quads
    static constexpr int N = (METHOD == 5) ? 25 : (METHOD == 3 ? 9 : 5);
    static constexpr int PW = CS_TQX + 2 * R;                          // staged region, quads
    static constexpr int PH = CS_TQY + 2 * R;
};

template <typename T> struct Quad { T r, g1, g2, b; };

template <typename T>
__device__ __forceinline__ Quad<T> load_quad(const T *img, int w, int h, int x, int y, bool even_w)
{
    Quad<T> q = {0, 0, 0, 0};
    if (x < 0 || y < 0 || x >= w || y >= h) return q;
    const T *p0 = img + (size_t)y * w + x;
    if (even_w && x + 1 < w && y + 1 < h) {
        if (sizeof(T) == 2) {
            uint32_t a = *reinterpret_cast<const uint32_t *>(p0), b = *reinterpret_cast<const uint32_t *>(p0 + w);
            q.r = (T)(a & 0xFFFF); q.g1 = (T)(a >> 16); q.g2 = (T)(b & 0xFFFF); q.b = (T)(b >> 16);
        } else {
            uint2 a = *reinterpret_cast<const uint2 *>(p0), b = *reinterpret_cast<const uint2 *>(p0 + w);
            q.r = (T)a.x; q.g1 = (T)a.y; q.g2 = (T)b.x; q.b = (T)b.y;
        }
        return q;
    }
    q.r = p0[0];
    if (x + 1 < w) q.g1 = p0[1];
    if (y + 1 < h) {
        q.g2 = p0[w];
        if (x + 1 < w) q.b = p0[w + 1];
    }
    return q;
}

template <typename T>
__device__ __forceinline__ void store_quad(T *img, int w, int h, int x, int y, bool even_w, Quad<T> q)
{
    if (x >= w || y >= h) return;
    T *p0 = img + (size_t)y * w + x;
    if (even_w && x + 1 < w && y + 1 < h) {
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint32_t *>(p0) = (uint32_t)q.r | ((uint32_t)q.g1 << 16);
            *reinterpret_cast<uint32_t *>(p0 + w) = (uint32_t)q.g2 | ((uint32_t)q.b << 16);
        } else {
            *reinterpret_cast<uint2 *>(p0) = make_uint2((uint32_t)q.r, (uint32_t)q.g1);
            *reinterpret_cast<uint2 *>(p0 + w) = make_uint2((uint32_t)q.g2, (uint32_t)q.b);
        }
        return;
    }
    p0[0] = q.r;
    if (x + 1 < w) p0[1] = q.g1;
    if (y + 1 < h) {
        p0[w] = q.g2;
        if (x + 1 < w) p0[w + 1] = q.b;
    }
}

// stripes.c:250-266: v > black+64 ? min(white, (v-black)*coef/65536 + black) : v   (exact in integers:
// the product is < 2^53 and the divisor a power of two, so the double expression truncates to >>16)
__device__ __forceinline__ uint16_t stripe_gain(uint16_t v, int coef, int black16, int white16)
{
    if (coef != 0 && (int)v > black16 + 64) {
        long long t = ((long long)((int)v - black16) * coef >> 16) + black16;
        return (uint16_t)min((long long)white16, t);
    }
    return v;
}

struct CsParams {
    int w, h, black;
    const int *raw2ev;            // indexed by raw value (pointer already offset by MAX_BLACK - black)
    const uint16_t *ev2raw_u16;   // 14-bit path: e in [0, 14 EV)
    const int *ev2raw_i32;        // 20-bit path: full table, pointer pre-offset
    size_t frame_stride;          // in elements
    int stripes;                  // fused stripes epilogue (uint16 only)
    int black16, white16;
    int coef[8];
    // Optional (dual-ISO planes): where the consumer provably never reads the smoothed sample: pixel i is dead iff
    // (dead_flags[i] & dead_mask) == dead_mask (dualiso.cu: final blend).  A quad whose R and B sites are both dead skips
    // its medians and table lookups (its samples pass through).
    const uint8_t *dead_flags;
    int dead_mask;
};

template <typename T, int METHOD>
__global__ void __launch_bounds__(CS_THREADS)
chroma_smooth_kernel(const T *__restrict__ in_base, T *__restrict__ out_base, const CsParams P)
{
    using G = CsGeom<METHOD>;
    __shared__ int s_ge[G::PH][G::PW];
    __shared__ int s_dr[G::PH][G::PW];
    __shared__ int s_db[G::PH][G::PW];

    const T *in = in_base + (size_t)blockIdx.z * P.frame_stride;
    T *out = out_base + (size_t)blockIdx.z * P.frame_stride;
    const int w = P.w, h = P.h;
    const bool even_w = (w & 1) == 0;
    const int qx0 = blockIdx.x * CS_TQX - G::R, qy0 = blockIdx.y * CS_TQY - G::R;

    auto quad_dead = [&](int x, int y) {                       // R at (x, y), B at (x + 1, y + 1)
        if (x + 1 >= w || y + 1 >= h) return false;
        const int m = P.dead_mask;
        return (P.dead_flags[x + (size_t)y * w] & m) == m && (P.dead_flags[x + 1 + (size_t)(y + 1) * w] & m) == m;
    };
    unsigned dead_bits = 0;                                     // bit k: this thread's k-th quad of pass 2 is dead
    if (P.dead_mask) {                                          // whole tile dead: copy through, no gathers, no medians
        int live = 0, k = 0;
        for (int i = threadIdx.x; i < CS_TQX * CS_TQY; i += CS_THREADS, k++) {
            const int ty = i / CS_TQX, tx = i - ty * CS_TQX;
            const int x = 2 * (blockIdx.x * CS_TQX + tx), y = 2 * (blockIdx.y * CS_TQY + ty);
            if (x < w && y < h) { if (quad_dead(x, y)) dead_bits |= 1u << k; else live = 1; }
        }
        if (!__syncthreads_or(live)) {
            for (int i = threadIdx.x; i < CS_TQX * CS_TQY; i += CS_THREADS) {
                const int ty = i / CS_TQX, tx = i - ty * CS_TQX;
                const int x = 2 * (blockIdx.x * CS_TQX + tx), y = 2 * (blockIdx.y * CS_TQY + ty);
                if (x < w && y < h) store_quad(out, w, h, x, y, even_w, load_quad(in, w, h, x, y, even_w));
            }
            return;
        }
    }

    // pass 1: EV triplets for tile + halo
    for (int i = threadIdx.x; i < G::PW * G::PH; i += CS_THREADS) {
        const int ly = i / G::PW, lx = i - ly * G::PW;
        const int x = 2 * (qx0 + lx), y = 2 * (qy0 + ly);
        int ge = 0, dr = 0, db = 0;
        if (x >= 0 && y >= 0 && x + 1 < w && y + 1 < h) {
            Quad<T> q = load_quad(in, w, h, x, y, even_w);
            ge = wadd(__ldg(P.raw2ev + q.g1), __ldg(P.raw2ev + q.g2)) / 2;
            dr = wsub(__ldg(P.raw2ev + q.r), ge);
            db = wsub(__ldg(P.raw2ev + q.b), ge);
        }
        s_ge[ly][lx] = ge; s_dr[ly][lx] = dr; s_db[ly][lx] = db;
    }
    __syncthreads();

    // pass 2: medians + rewrite
    for (int i = threadIdx.x, k2 = 0; i < CS_TQX * CS_TQY; i += CS_THREADS, k2++) {
        const int ty = i / CS_TQX, tx = i - ty * CS_TQX;
        const int x = 2 * (blockIdx.x * CS_TQX + tx), y = 2 * (blockIdx.y * CS_TQY + ty);
        if (x >= w || y >= h) continue;
        Quad<T> q = load_quad(in, w, h, x, y, even_w);
        // chroma_smooth.c:26-28 loop bounds
        if (y >= 4 && y < h - 5 && x >= 4 && x < w - 4 && !(dead_bits >> k2 & 1)) {
            const int ly = ty + G::R, lx = tx + G::R;
            const int ge = s_ge[ly][lx];
            if (ge >= 2 * MLVB_EV_RES) {
                int mr[G::N], mb[G::N];
                int k = 0;
#pragma unroll
                for (int dj = -G::R; dj <= G::R; dj++)
#pragma unroll
                    for (int di = -G::R; di <= G::R; di++) {
                        if (METHOD == 2 && di != 0 && dj != 0) continue;   // plus-shaped 5-tap (chroma_smooth.c:45-48)
                        mr[k] = s_dr[ly + dj][lx + di];
                        mb[k] = s_db[ly + dj][lx + di];
                        k++;
                    }
                int dr, db;
                if constexpr (METHOD == 2) { dr = median5(reinterpret_cast<int(&)[5]>(mr)); db = median5(reinterpret_cast<int(&)[5]>(mb)); }
                else if constexpr (METHOD == 3) { dr = median9(reinterpret_cast<int(&)[9]>(mr)); db = median9(reinterpret_cast<int(&)[9]>(mb)); }
                else { dr = median25(reinterpret_cast<int(&)[25]>(mr)); db = median25(reinterpret_cast<int(&)[25]>(mb)); }
                const int er = wadd(ge, dr), eb = wadd(ge, db);
                if (er > MLVB_EV_RES && eb > MLVB_EV_RES) {
                    if (sizeof(T) == 2) {
                        q.r = (T)(__ldg(P.ev2raw_u16 + clamp_ev(er)) + P.black);
                        q.b = (T)(__ldg(P.ev2raw_u16 + clamp_ev(eb)) + P.black);
                    } else {
                        q.r = (T)(__ldg(P.ev2raw_i32 + clamp_ev(er)) + P.black);
                        q.b = (T)(__ldg(P.ev2raw_i32 + clamp_ev(eb)) + P.black);
                    }
                }
            }
        }
        if (sizeof(T) == 2 && P.stripes) {
            const int c0 = P.coef[x & 7], c1 = P.coef[(x + 1) & 7];
            q.r = (T)stripe_gain((uint16_t)q.r, c0, P.black16, P.white16);
            q.g1 = (T)stripe_gain((uint16_t)q.g1, c1, P.black16, P.white16);
            q.g2 = (T)stripe_gain((uint16_t)q.g2, c0, P.black16, P.white16);
            q.b = (T)stripe_gain((uint16_t)q.b, c1, P.black16, P.white16);
        }
        store_quad(out, w, h, x, y, even_w, q);
    }
}

// stripes only (no chroma smoothing requested): elementwise, 8 pixels (one coefficient period) per thread
__global__ void stripes_apply_kernel(uint16_t *__restrict__ img, size_t n, size_t frame_stride, int black16, int white16,
                                     const int c0, const int c1, const int c2, const int c3, const int c4, const int c5,
                                     const int c6, const int c7)
{
    const int coef[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
    uint16_t *p = img + (size_t)blockIdx.y * frame_stride;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t i0 = g * 8;
    if (i0 >= n) return;
    if (i0 + 8 <= n && ((uintptr_t)p % 16 == 0)) {
        uint4 v = reinterpret_cast<uint4 *>(p)[g];
        uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint16_t lo = stripe_gain((uint16_t)(wv[k] & 0xFFFF), coef[2 * k], black16, white16);
            uint16_t hi = stripe_gain((uint16_t)(wv[k] >> 16), coef[2 * k + 1], black16, white16);
            wv[k] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        reinterpret_cast<uint4 *>(p)[g] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    } else {
        for (size_t i = i0; i < n && i < i0 + 8; i++) p[i] = stripe_gain(p[i], coef[i & 7], black16, white16);
    }
}

template <typename T, int METHOD>
void launch_cs(const T *in, T *out, const CsParams &P, int nframes, cudaStream_t st)
{
    const int qw = (P.w + 1) / 2, qh = (P.h + 1) / 2;
    dim3 grid(ceil_div(qw, CS_TQX), ceil_div(qh, CS_TQY), nframes);
    chroma_smooth_kernel<T, METHOD><<<grid, CS_THREADS, 0, st>>>(in, out, P);
}

// ------------------------------------------------------------------------------------------------
// 3x3 fast path (uint16, even w and h): register-only "warp strip" kernel.
//
// A warp owns 32 adjacent quad columns (lanes 1..30 produce output, lanes 0 and 31 are the halo
// shared with the neighbouring strips) and walks down CS3_ROWS quad rows.  Every lane keeps the
// (ge, dr, db) triplets of its own column for the rows y-1, y, y+1 in registers, sorts that
// 3-element column once per step and obtains the neighbours' sorted columns with shuffles; the
// median of 9 is then  med3( max(lows), med3(mids), min(highs) ).  No shared memory, no halo
// recomputation beyond 2/32 lanes and 2/(CS3_ROWS+2) rows, no per-pixel index arithmetic.
constexpr int CS3_ROWS = 34;         // output quad rows per warp
constexpr int CS3_WARPS = 4;

struct RowQ { uint32_t top, bot; int ge, dr, db; };      // top = r | g1<<16, bot = g2 | b<<16

__device__ __forceinline__ void sort3(int &a, int &b, int &c)
{
    int t = min(a, b); b = max(a, b); a = t;
    t = min(b, c); c = max(b, c); b = t;
    t = min(a, b); b = max(a, b); a = t;
}
__device__ __forceinline__ int med3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

__device__ __forceinline__ RowQ load_rowq(const uint16_t *img, int w, int ph, int x, int qr, bool col_ok, const int *raw2ev)
{
    RowQ q = {0u, 0u, 0, 0, 0};
    if (col_ok && qr >= 0 && qr < ph) {
        const uint16_t *p = img + (size_t)(2 * qr) * w + x;
        q.top = *reinterpret_cast<const uint32_t *>(p);
        q.bot = *reinterpret_cast<const uint32_t *>(p + w);
        q.ge = wadd(__ldg(raw2ev + (q.top >> 16)), __ldg(raw2ev + (q.bot & 0xFFFF))) / 2;
        q.dr = wsub(__ldg(raw2ev + (q.top & 0xFFFF)), q.ge);
        q.db = wsub(__ldg(raw2ev + (q.bot >> 16)), q.ge);
    }
    return q;
}

__device__ __forceinline__ int median9_columns(int a, int b, int c)
{
    sort3(a, b, c);                                               // own column: a <= b <= c
    const int al = __shfl_up_sync(0xFFFFFFFFu, a, 1), ar = __shfl_down_sync(0xFFFFFFFFu, a, 1);
    const int bl = __shfl_up_sync(0xFFFFFFFFu, b, 1), br = __shfl_down_sync(0xFFFFFFFFu, b, 1);
    const int cl = __shfl_up_sync(0xFFFFFFFFu, c, 1), cr = __shfl_down_sync(0xFFFFFFFFu, c, 1);
    return med3(max(max(al, a), ar), med3(bl, b, br), min(min(cl, c), cr));
}

// exact stripes.c:250-266 for 0 < coef: ((v - black) * coef >> 16) + black, clamped at white
__device__ __forceinline__ uint32_t stripe_gain_fast(uint32_t v, uint32_t coef, int black16, int white16)
{
    if ((int)v > black16 + 64) {
        const uint32_t t = __umulhi((v - (uint32_t)black16) << 16, coef) + (uint32_t)black16;
        return min(t, (uint32_t)white16);
    }
    return v;
}

template <bool STRIPES>
__global__ void __launch_bounds__(CS3_WARPS * 32)
chroma3_strip_kernel(const uint16_t *__restrict__ in_base, uint16_t *__restrict__ out_base, const CsParams P)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = P.w, pw = w >> 1, ph = P.h >> 1;
    const int strip = blockIdx.x * CS3_WARPS + warp;
    const int qc = strip * 30 + lane - 1;
    if (strip * 30 >= pw) return;                                  // whole warp out of the frame
    const bool col_ok = qc >= 0 && qc < pw;
    const int x = 2 * qc;
    const uint16_t *in = in_base + (size_t)blockIdx.z * P.frame_stride;
    uint16_t *out = out_base + (size_t)blockIdx.z * P.frame_stride;
    const int *raw2ev = P.raw2ev;
    const int qr0 = blockIdx.y * CS3_ROWS, qr1 = min(qr0 + CS3_ROWS, ph);
    const bool x_inside = x >= 4 && x < w - 4;                     // chroma_smooth.c:28
    const bool writer = col_ok && lane >= 1 && lane <= 30;
    uint32_t c0 = 0, c1 = 0;
    if (STRIPES) { c0 = (uint32_t)P.coef[x & 7]; c1 = (uint32_t)P.coef[(x + 1) & 7]; }

    RowQ A = load_rowq(in, w, ph, x, qr0 - 1, col_ok, raw2ev);
    RowQ B = load_rowq(in, w, ph, x, qr0, col_ok, raw2ev);
    for (int qr = qr0; qr < qr1; qr++) {
        const RowQ C = load_rowq(in, w, ph, x, qr + 1, col_ok, raw2ev);
        const int mr = median9_columns(A.dr, B.dr, C.dr);
        const int mb = median9_columns(A.db, B.db, C.db);
        uint32_t top = B.top, bot = B.bot;
        const int y = 2 * qr;
        if (x_inside && y >= 4 && y < P.h - 5 && B.ge >= 2 * MLVB_EV_RES) {      // chroma_smooth.c:26,35
            const int er = wadd(B.ge, mr), eb = wadd(B.ge, mb);
            if (er > MLVB_EV_RES && eb > MLVB_EV_RES) {                          // chroma_smooth.c:63-64
                const uint32_t r = (uint32_t)(__ldg(P.ev2raw_u16 + clamp_ev(er)) + P.black) & 0xFFFFu;
                const uint32_t b = (uint32_t)(__ldg(P.ev2raw_u16 + clamp_ev(eb)) + P.black) & 0xFFFFu;
                top = (top & 0xFFFF0000u) | r;
                bot = (bot & 0x0000FFFFu) | (b << 16);
            }
        }
        if (STRIPES) {
            top = stripe_gain_fast(top & 0xFFFF, c0, P.black16, P.white16) | (stripe_gain_fast(top >> 16, c1, P.black16, P.white16) << 16);
            bot = stripe_gain_fast(bot & 0xFFFF, c0, P.black16, P.white16) | (stripe_gain_fast(bot >> 16, c1, P.black16, P.white16) << 16);
        }
        if (writer) {
            uint16_t *o = out + (size_t)y * w + x;
            *reinterpret_cast<uint32_t *>(o) = top;
            *reinterpret_cast<uint32_t *>(o + w) = bot;
        }
        A = B;
        B = C;
    }
}

void launch_cs3_fast(const uint16_t *in, uint16_t *out, const CsParams &P, int nframes, cudaStream_t st)
{
    const int pw = P.w / 2, ph = P.h / 2;
    dim3 grid(ceil_div(ceil_div(pw, 30), CS3_WARPS), ceil_div(ph, CS3_ROWS), nframes);
    if (P.stripes) chroma3_strip_kernel<true><<<grid, CS3_WARPS * 32, 0, st>>>(in, out, P);
    else chroma3_strip_kernel<false><<<grid, CS3_WARPS * 32, 0, st>>>(in, out, P);
}

}  // namespace

int launch_chroma_smooth_u16(const uint16_t *d_in, uint16_t *d_out, int w, int h, size_t frame_stride, int nframes,
                             int black, int method, const EvLuts &luts, const StripeCoef *stripes, int white,
                             cudaStream_t st)
{
    if (black > MLVB_MAX_BLACK) return MLVB_ERR_ARG;     // reference: get_raw2ev returns NULL (main.c:170-174)
    CsParams P = {};
    P.w = w; P.h = h; P.black = black;
    P.raw2ev = luts.raw2ev_base + (MLVB_MAX_BLACK - black);
    P.ev2raw_u16 = luts.ev2raw_pos;
    P.ev2raw_i32 = luts.ev2raw_full;
    P.frame_stride = frame_stride;
    P.stripes = 0;
    if (stripes && stripes->needed && (w % 8 == 0)) {
        P.stripes = 1;
        P.black16 = (uint16_t)black; P.white16 = (uint16_t)white;
        for (int i = 0; i < 8; i++) P.coef[i] = stripes->coef[i];
    }
    switch (method) {
    case 2: launch_cs<uint16_t, 2>(d_in, d_out, P, nframes, st); break;
    case 3: {
        bool fast = (w % 2 == 0) && (h % 2 == 0) && w >= 8 && h >= 8 && ((uintptr_t)d_in % 4 == 0) && ((uintptr_t)d_out % 4 == 0) &&
                    (frame_stride % 2 == 0);
        for (int i = 0; i < 8 && P.stripes; i++) fast = fast && P.coef[i] > 0;     // fast gain needs positive coefficients
        if (fast) launch_cs3_fast(d_in, d_out, P, nframes, st);
        else launch_cs<uint16_t, 3>(d_in, d_out, P, nframes, st);
        break;
    }
    case 5: launch_cs<uint16_t, 5>(d_in, d_out, P, nframes, st); break;
    default: return MLVB_ERR_ARG;
    }
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_chroma_smooth_u32(const uint32_t *d_in, uint32_t *d_out, int w, int h, int method, const int *d_raw2ev,
                             const int *d_ev2raw, cudaStream_t st, const uint8_t *d_dead_flags, int dead_mask)
{
    CsParams P = {};
    P.w = w; P.h = h; P.black = 0;
    P.dead_flags = d_dead_flags; P.dead_mask = d_dead_flags ? dead_mask : 0;
    P.raw2ev = d_raw2ev; P.ev2raw_i32 = d_ev2raw; P.ev2raw_u16 = nullptr;
    P.frame_stride = (size_t)w * h;
    switch (method) {
    case 2: launch_cs<uint32_t, 2>(d_in, d_out, P, 1, st); break;
    case 3: launch_cs<uint32_t, 3>(d_in, d_out, P, 1, st); break;
    case 5: launch_cs<uint32_t, 5>(d_in, d_out, P, 1, st); break;
    default: return MLVB_ERR_ARG;
    }
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

int launch_stripes_apply(uint16_t *d_img, int w, size_t npix, size_t frame_stride, int nframes, int black, int white,
                         const StripeCoef *sc, cudaStream_t st)
{
    if (!sc || !sc->needed || (w % 8) != 0) return MLVB_OK;      // stripes.c:252-253
    dim3 grid(ceil_div((npix + 7) / 8, 256), nframes);
    stripes_apply_kernel<<<grid, 256, 0, st>>>(d_img, npix, frame_stride, (uint16_t)black, (uint16_t)white, sc->coef[0],
                                               sc->coef[1], sc->coef[2], sc->coef[3], sc->coef[4], sc->coef[5],
                                               sc->coef[6], sc->coef[7]);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
