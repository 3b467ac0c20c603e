// fused_wide.cuh -- batch form of the fused single-ISO chain (unpack + bad-pixel patches + 3x3 median
// chroma smoothing + stripe gains, one pass over HBM), written for the instruction roofline that bounds
// the strip kernel in fused.cu (profiles/r01d_fused3_ncu.md):
//
//   * one persistent CTA per SM; both EV tables live in shared memory: raw2ev for this black level
//     (16384 x int32, 64 KB) and the top octave of ev2raw (32768 x uint16, 64 KB) from which every other
//     octave is a right shift (ev2raw[e] == ev2raw[13 EV + e mod EV] >> (13 - e / EV), verified entry by
//     entry when the context is created) -- no table gather goes through the L1 tag stage;
//   * a lane owns FW_COLS adjacent RGGB quad columns (2 * FW_COLS pixels of a row), so every bit offset is a
//     compile-time constant (one shift + one mask per pixel), neighbouring sorted columns are mostly in the
//     lane's own registers (12 shuffles per FW_COLS quads instead of per quad) and every stripe gain has a
//     fixed column;
//   * the packed rows are staged by cp.async (16-byte chunks, L2 only) into a per-warp double buffer, one
//     quad row ahead; bad-pixel patches are written into the staged bytes before extraction.  -DFW_STAGE_TMA
//     stages the same bytes with cp.async.bulk + one mbarrier per warp and slot instead (bit-identical, measured
//     325 k against 360 k frames/s on C2: the single issuing lane's address / uniform-register code costs more
//     issue slots than the 32-lane LDGSTS it replaces, and issue slots are what this kernel is short of);
//   * sort3 / med3 are evaluated as (min3, max3, a + b + c - min3 - max3): the additions issue on the FMA
//     pipe (IMAD) while the min/max run on the ALU pipe, instead of 6 / 4 ALU-pipe min/max.
//
// Arithmetic is the same 32-bit wrap-around integer arithmetic as fused3_strip_kernel / the reference
// (cs.c:49-84, chroma_smooth.c:22-71, stripes.c:250-266); results are bit-identical.
#pragma once

#include <vector>

#ifdef FW_PLAIN_STORES
#define FW_STORE(p, v) (*(p) = (v))
#else
#define FW_STORE(p, v) __stcs(p, v)
#endif

namespace {

// Defaults, each measured on B200 (profiles/r01m_variants.md); -DFW_NO_<name> switches one off.
#if !defined(FW_NO_MIDPAIR) && !defined(FW_MIDPAIR)
#define FW_MIDPAIR
#endif
#if !defined(FW_NO_PACK_PRMT) && !defined(FW_PACK_PRMT)
#define FW_PACK_PRMT
#endif
#if !defined(FW_NO_EXTRACT_SHL) && !defined(FW_EXTRACT_SHL)
#define FW_EXTRACT_SHL
#endif
#if !defined(FW_NO_GAIN_X) && !defined(FW_GAIN_X) && defined(FW_PACK_PRMT) && !defined(FW_GAIN_HI) && !defined(FW_GAIN_MAD)
#define FW_GAIN_X
#endif
#if !defined(FW_NO_CLAMP_X2) && !defined(FW_CLAMP_X2) && defined(FW_GAIN_X)
#define FW_CLAMP_X2
#endif
#if !defined(FW_NO_ADD_FMA) && !defined(FW_ADD_FMA)
#define FW_ADD_FMA
#endif
#if defined(FW_CLAMP_X2) && !defined(FW_GAIN_X)
#error "FW_CLAMP_X2 needs FW_GAIN_X (samples are packed before the clamp)"
#endif
#if defined(FW_GAIN_X) && !defined(FW_PACK_PRMT)
#error "FW_GAIN_X needs FW_PACK_PRMT (the packing PRMT takes the high halves)"
#endif

#ifndef FW_COLS_CFG
#define FW_COLS_CFG 4
#endif
#ifndef FW_PATCHES
#define FW_PATCHES 1
#endif
#ifndef FW_WARPS_CFG
#define FW_WARPS_CFG 16
#endif
constexpr int FW_COLS = FW_COLS_CFG;             // quad columns per lane (4 or 8)
constexpr int FW_WARPS = FW_WARPS_CFG;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int FW_LANE_PX = 2 * FW_COLS;          // pixels of a row per lane
constexpr int FW_LANE_BYTES = FW_LANE_PX * 14 / 8;   // 14 or 28 stream bytes
constexpr int FW_STRIP_PX = 30 * FW_LANE_PX;     // pixels a warp finishes per row (lanes 0 and 31 are halo)
constexpr int FW_WINDOW_PX = 32 * FW_LANE_PX;
constexpr int FW_NW = FW_LANE_PX * 14 / 32 + ((FW_LANE_PX * 14) % 32 ? 1 : 0);   // stream words holding a lane's pixels (4 or 7)
constexpr int FW_ROWBYTES = FW_COLS == 4 ? 480 : 960;   // staged bytes per pixel row and warp (16-byte chunks)
constexpr int FW_STAGE_PER_WARP = 4 * FW_ROWBYTES;       // 2 slots x 2 pixel rows
constexpr int FW_SMEM_R2E = 16384 * 4;
constexpr int FW_SMEM_T13 = 32768 * 2;
constexpr int FW_SMEM_STAGE = FW_WARPS * FW_STAGE_PER_WARP;
constexpr int FW_SMEM_BYTES = FW_SMEM_R2E + FW_SMEM_T13 + FW_SMEM_STAGE + FW_WARPS * 2 * 8;   // + one mbarrier per warp and slot
static_assert(FW_COLS == 4 || FW_COLS == 8, "a lane owns 8 or 16 pixels of a row");

#ifdef MLVB_HOST_EMU
static thread_local uint8_t *fw_smem;              // tests/emu/wide_emu.cpp: the block's "shared memory", one host thread per lane
#else
extern __shared__ __align__(16) uint8_t fw_smem[];
#endif
#define FW_R2E(v) (reinterpret_cast<const int *>(fw_smem)[(v)])
#define FW_T13(f) (reinterpret_cast<const uint16_t *>(fw_smem + FW_SMEM_R2E)[(f)])

struct WideItem { unsigned short px_sub; unsigned short qrow; unsigned entry; };   // px_sub: pixel column inside the strip's window | (y & 1) << 13

// Host: the kernel's patch lists for n repaired pixels (list order; get(m, x, y) gives entry m's position inside the
// frame).  A strip is a window of 32 lanes starting one lane left of pixel FW_STRIP_PX * s (lanes 0 and 31 are halo), so
// an entry near a seam is listed for both strips.  items are grouped by strip, then quad row; row_start[s * (ph + 1) + q]
// is the first item of strip s at quad row q or below it.  Only interior entries are repaired (cs.c:317).
template <class GetXY>
inline void wide_build_items(size_t n, GetXY get, int w, int h, std::vector<WideItem> &items, std::vector<unsigned> &row_start)
{
    const int ph = h / 2, wstrips = (w + FW_STRIP_PX - 1) / FW_STRIP_PX;
    std::vector<std::vector<WideItem>> wb((size_t)wstrips * ph);
    for (size_t m = 0; m < n; m++) {
        int x, y;
        get(m, x, y);
        if (!(x > 2 && x < w - 3 && y > 2 && y < h - 3)) continue;
        for (int s = 0; s < wstrips; s++) {
            const int px = x - (FW_STRIP_PX * s - FW_LANE_PX);
            if (px < 0 || px >= FW_WINDOW_PX) continue;
            wb[(size_t)s * ph + (y >> 1)].push_back(WideItem{(unsigned short)(px | ((y & 1) << 13)), (unsigned short)(y >> 1), (unsigned)m});
        }
    }
    row_start.assign((size_t)wstrips * (ph + 1), 0);
    items.clear();
    for (int s = 0; s < wstrips; s++) {
        for (int q = 0; q < ph; q++) {
            row_start[(size_t)s * (ph + 1) + q] = (unsigned)items.size();
            auto &v = wb[(size_t)s * ph + q];
            items.insert(items.end(), v.begin(), v.end());
        }
        row_start[(size_t)s * (ph + 1) + ph] = (unsigned)items.size();
    }
}

struct WideParams {
    const uint8_t *packed; size_t payload_stride;
    uint16_t *out; size_t out_stride;
    int w, h, black;
    const int *raw2ev;                // indexed by raw value (16384 entries)
    const uint16_t *ev2raw13;         // ev2raw[13 EV ...], 32768 entries
    int black16, white16;
    struct { unsigned coef, k1; } gain[8];   // stripe gain and -black * gain (mod 2^32)
    unsigned coefh[8], k4;            // gain << 14 and -4 * black: ((v - black) * gain) >> 16 as one high multiply (-DFW_GAIN_HI)
    struct { unsigned coef, kx; } gainx[8];  // kx = (black << 16) - black * gain: v * gain + kx holds the corrected sample in its
    unsigned whitex;                  // high half (-DFW_GAIN_X); whitex = white << 16 | 0xFFFF clamps it there
    int one, mone;                    // +1 / -1: multipliers of the FMA-pipe additions (opaque to the compiler)
    unsigned shl[4];                  // 2^14, 2^10, 2^6, 2^2: left shifts done as FMA-pipe multiplies (-DFW_EXTRACT_SHL)
    unsigned shr[2];                  // 2^14, 2^17: right shifts by 18 / 15 as FMA-pipe high multiplies (-DFW_SHR_HI)
    const WideItem *items; const unsigned *row_start; const uint16_t *vals; unsigned n_entries;
    int nstrips, nseg, seg_rows, nframes;
};

#ifdef MLVB_HOST_EMU
// host emulation: the copy happens at once, groups and waits are no-ops (the __syncwarp after the wait orders it)
inline void cp_async16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait1() {}
inline void cp_async_wait0() {}
#else
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// TMA (bulk async copy) row staging: one elected lane arms the slot's mbarrier with the byte count and issues
// the copies; every lane of the warp waits on the barrier's phase before reading the staged bytes.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

#endif

__device__ __forceinline__ int imin3(int a, int b, int c) { return __vimin3_s32(a, b, c); }   // one VIMNMX3
__device__ __forceinline__ int imax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

struct WideConst { uint32_t black, thr, white, white2; uint32_t coef[8], k1[8], coefh[8], k4, kx[8], whitex; uint32_t shl[4], shr[2]; int one, mone; };

// a * m + b on the FMA pipe (IMAD with a register multiplier the compiler cannot fold): the min/max network
// keeps the ALU pipe busy, the bookkeeping additions go next door.  m is +1 or -1.
__device__ __forceinline__ int fma_add(int a, int m, int b)
{
#ifdef MLVB_HOST_EMU
    return (int)((unsigned)a * (unsigned)m + (unsigned)b);
#else
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b));
    return d;
#endif
}
// -DFW_ADD_FMA: the two-input additions of the chain (ev - ge, ge + median, + black) as one IMAD each on the FMA pipe
// instead of one IADD3 each on the ALU pipe -- the same number of issue slots, taken from the pipe that has room
__device__ __forceinline__ int add2(int a, int b, const WideConst &K)
{
#ifdef FW_ADD_FMA
    return fma_add(a, K.one, b);
#else
    return (int)((unsigned)a + (unsigned)b);
#endif
}
__device__ __forceinline__ int sub2(int a, int b, const WideConst &K)     // a - b
{
#ifdef FW_ADD_FMA
    return fma_add(b, K.mone, a);
#else
    return (int)((unsigned)a - (unsigned)b);
#endif
}
// the middle of three = a + b + c - min - max; exact in wrap-around arithmetic (a permutation of the three)
__device__ __forceinline__ int mid_of(int a, int b, int c, int lo, int hi, const WideConst &K)
{
#if defined(FW_SUMS_HALF)
    return fma_add(hi, K.mone, fma_add(lo, K.mone, (int)((unsigned)a + (unsigned)b + (unsigned)c)));   // one IADD3 + two IMAD
#elif !defined(FW_SUMS_ON_FMA)
    return (int)((unsigned)a + (unsigned)b + (unsigned)c - (unsigned)lo - (unsigned)hi);
#else
    return fma_add(hi, K.mone, fma_add(lo, K.mone, fma_add(c, K.one, fma_add(a, K.one, b))));
#endif
}
__device__ __forceinline__ int imed3(int a, int b, int c, const WideConst &K)
{
    return mid_of(a, b, c, imin3(a, b, c), imax3(a, b, c), K);
}

// pixel I of a lane's pixel group; W[] are the group's 32-bit words in stream order
template <int I>
__device__ __forceinline__ uint32_t wide_px(const uint32_t (&W)[FW_NW], const uint32_t (&shl)[4], uint32_t m14)
{
    constexpr int bit = 14 * I, j = bit >> 5, s = bit & 31;
    uint32_t v;
#if defined(FW_SHR_HI)
    // both pipes issue one instruction per two cycles and the ALU pipe is the full one: x >> 18 as the high word of
    // x * 2^14 (IMAD.HI with a kernel-parameter multiplier the compiler cannot turn back into a shift) runs next door
    if constexpr (s == 0) v = __umulhi(W[j], m14);
    else if constexpr (s == 14 || s == 10 || s == 6 || s == 2) v = __umulhi(W[j] * shl[(14 - s) / 4], m14);
    else if constexpr (s <= 18) v = (W[j] >> (18 - s)) & 0x3FFFu;
    else if constexpr (s > 18) v = __umulhi(__funnelshift_l(W[j + 1], W[j], s), m14);
#elif defined(FW_EXTRACT_SHL)
    // (W << s) >> 18 with the left shift as a multiply by a kernel parameter (2^s, opaque to the compiler): the
    // multiply issues on the FMA pipe and leaves one ALU-pipe shift instead of shift + mask
    if constexpr (s == 0) v = W[j] >> 18;
    else if constexpr (s == 14 || s == 10 || s == 6 || s == 2) v = (W[j] * shl[(14 - s) / 4]) >> 18;
    else if constexpr (s <= 18) v = (W[j] >> (18 - s)) & 0x3FFFu;
#else
    if constexpr (s <= 18) v = (W[j] >> (18 - s)) & 0x3FFFu;
#endif
#if !defined(FW_SHR_HI)
    else v = __funnelshift_l(W[j + 1], W[j], s) >> 18;
#endif
    asm("" : "+r"(v));          // keep the table address a plain v * 4 (one IMAD) instead of a second shift + mask of W
    return v;
}

struct WideRow {                                  // per-lane state of one quad row
    int dr[FW_COLS], db[FW_COLS];                 // ev(r) - ge, ev(b) - ge
    int ge[FW_COLS];
    uint32_t r[FW_COLS], b[FW_COLS];              // raw R / B samples
    uint32_t g1s[FW_COLS], g2[FW_COLS];           // finished G1 << 16, finished G2
};

// STRIPES: 0 = no stripe correction, 1 = eight general gains, 2 = gains 0 and 1 are exactly 1.0 (what
// stripes_compute_correction always produces, stripes.c:236-237) and white > black + 64: those two columns
// only need the clamp to white
template <int STRIPES, int IDX>
__device__ __forceinline__ uint32_t wide_gain(uint32_t v, const WideConst &K)
{
    if (STRIPES == 2 && IDX < 2) return min(v, K.white);
    if (STRIPES) {
        // ((v - black) * coef >> 16) + black as one multiply-add: k1 = -black * coef (mod 2^32); the product is exact
        // for v > black (coef < 2^18, host check).  stripes.c:258: only v > black + 64 is touched.
#if defined(FW_GAIN_HI)
        // (v - black) * 4 < 2^16 and gain << 14 < 2^32: the high word of their product is ((v - black) * gain) >> 16,
        // (IMAD.HI's addend is a 64-bit pair, so black is added by the clamp: VIADDMNMX) -- two FMA-pipe instructions
        // and one ALU-pipe instruction under the predicate
        if (v > K.thr) {
            // clamp before adding black: the addition stays out of the IMAD.HI (its addend would need a register pair)
            return min(__umulhi(v * 4u + K.k4, K.coefh[IDX]), K.white - K.black) + K.black;
        }
#elif defined(FW_GAIN_MAD)
        const uint32_t t = min(((v * K.coef[IDX] + K.k1[IDX]) >> 16) + K.black, K.white);
        return v > K.thr ? t : v;
#else
        if (v > K.thr) return min((((v - K.black) * K.coef[IDX]) >> 16) + K.black, K.white);
#endif
    }
    return v;
}

// -DFW_GAIN_X: a corrected sample is left in the HIGH half of its register: X = v * gain + ((black << 16) - black * gain)
// = (v - black) * gain + (black << 16), so X >> 16 is the reference's ((v - black) * gain >> 16) + black and the clamp to
// white is min(X, white << 16 | 0xFFFF).  The shift is free: the PRMT that packs two samples into a word picks the
// high bytes.  One ALU-pipe instruction less per corrected pixel (the shift-and-add LEA.HI).  The host checks that X
// cannot overflow 32 bits.
template <int STRIPES, int IDX>
__host__ __device__ constexpr bool wide_in_hi()
{
#ifdef FW_GAIN_X
    return STRIPES == 1 || (STRIPES == 2 && IDX >= 2);
#else
    return false;
#endif
}
// -DFW_CLAMP_X2 (with FW_GAIN_X): the clamp to white is taken after the packing PRMT, on both 16-bit halves of the
// output word at once (one packed unsigned min per two samples instead of one min per sample).  Equal results:
// min(X, white << 16 | 0xFFFF) >> 16 == min(X >> 16, white), and samples that are not corrected are below white.
template <int STRIPES, int IDX>
__device__ __forceinline__ uint32_t wide_gain_x(uint32_t v, const WideConst &K)
{
    if constexpr (!wide_in_hi<STRIPES, IDX>()) {
#ifdef FW_CLAMP_X2
        if (STRIPES == 2 && IDX < 2) return v;            // clamp only: done on the packed word
#endif
        return wide_gain<STRIPES, IDX>(v, K);
    } else {
        uint32_t x = v << 16;
#ifdef FW_CLAMP_X2
        if (v > K.thr) x = v * K.coef[IDX] + K.kx[IDX];
#else
        if (v > K.thr) x = min(v * K.coef[IDX] + K.kx[IDX], K.whitex);
#endif
        return x;
    }
}
template <int STRIPES>
__device__ __forceinline__ uint32_t wide_clamp2(uint32_t packed, const WideConst &K)
{
#ifdef FW_CLAMP_X2
    if (STRIPES) return __vminu2(packed, K.white2);
#endif
    return packed;
}
// low half of the result = sample A, high half = sample B; AH / BH: the sample sits in the high half of its register
template <bool AH, bool BH>
__device__ __forceinline__ uint32_t wide_pack(uint32_t a, uint32_t b)
{
    return __byte_perm(a, b, (AH ? 0x32u : 0x10u) | ((BH ? 0x76u : 0x54u) << 8));
}

template <int STRIPES, int C>
__device__ __forceinline__ void wide_ingest_col(const uint32_t (&T)[FW_NW], const uint32_t (&B)[FW_NW], const WideConst &K, WideRow &R)
{
    const uint32_t r = wide_px<2 * C>(T, K.shl, K.shr[0]), g1 = wide_px<2 * C + 1>(T, K.shl, K.shr[0]);
    const uint32_t g2 = wide_px<2 * C>(B, K.shl, K.shr[0]), b = wide_px<2 * C + 1>(B, K.shl, K.shr[0]);
    const int ge = wadd(FW_R2E(g1), FW_R2E(g2)) / 2;
    R.ge[C] = ge;
    R.dr[C] = sub2(FW_R2E(r), ge, K);
    R.db[C] = sub2(FW_R2E(b), ge, K);
    R.r[C] = r;
    R.b[C] = b;
#ifdef FW_PACK_PRMT
    R.g1s[C] = wide_gain_x<STRIPES, (2 * C + 1) & 7>(g1, K);        // packed next to R by one PRMT in wide_finish_col
    R.g2[C] = wide_gain_x<STRIPES, (2 * C) & 7>(g2, K);
#else
    R.g1s[C] = wide_gain<STRIPES, (2 * C + 1) & 7>(g1, K) << 16;
    R.g2[C] = wide_gain<STRIPES, (2 * C) & 7>(g2, K);
#endif
}

struct Tri { int lo, mid, hi; };

__device__ __forceinline__ Tri wide_sort3(int a, int b, int c, const WideConst &K)
{
    Tri t;
    t.lo = imin3(a, b, c);
    t.hi = imax3(a, b, c);
    t.mid = mid_of(a, b, c, t.lo, t.hi, K);
    return t;
}
__device__ __forceinline__ int wide_med9(const Tri &l, const Tri &m, const Tri &r, const WideConst &K)
{
    return imed3(imax3(l.lo, m.lo, r.lo), imed3(l.mid, m.mid, r.mid, K), imin3(l.hi, m.hi, r.hi), K);
}
// -DFW_MIDPAIR: the median of the three column medians as max(min(a, b), min(max(a, b), c)) with (a, b) the pair of
// columns two neighbouring windows share (columns 2k, 2k + 1 of a lane): 6 two-input min/max per two windows instead
// of 2 x (2 three-input min/max + 2 three-input adds)
struct MidPair { int mn, mx; };
__device__ __forceinline__ MidPair mid_pair(const Tri &a, const Tri &b) { MidPair p; p.mn = min(a.mid, b.mid); p.mx = max(a.mid, b.mid); return p; }
__device__ __forceinline__ int wide_med9_pair(const Tri &l, const Tri &m, const Tri &r, const MidPair &p, int third, const WideConst &K)
{
    return imed3(imax3(l.lo, m.lo, r.lo), max(p.mn, min(p.mx, third)), imin3(l.hi, m.hi, r.hi), K);
}
__device__ __forceinline__ Tri tri_shfl_up(const Tri &t)
{
    Tri o;
    o.lo = __shfl_up_sync(0xFFFFFFFFu, t.lo, 1); o.mid = __shfl_up_sync(0xFFFFFFFFu, t.mid, 1); o.hi = __shfl_up_sync(0xFFFFFFFFu, t.hi, 1);
    return o;
}
__device__ __forceinline__ Tri tri_shfl_down(const Tri &t)
{
    Tri o;
    o.lo = __shfl_down_sync(0xFFFFFFFFu, t.lo, 1); o.mid = __shfl_down_sync(0xFFFFFFFFu, t.mid, 1); o.hi = __shfl_down_sync(0xFFFFFFFFu, t.hi, 1);
    return o;
}

// ev2raw[min(e, 14 EV - 1)] for e > 0 from the top-octave table.  With t = (14 EV - 1) - e clamped at 0:
// octave shift 13 - (e >> 15) == t >> 15 and the table index e & (EV - 1) == ~t & (EV - 1).
__device__ __forceinline__ uint32_t wide_ev2raw(int e, uint32_t m17)
{
    const int t = max(MLVB_EV_MAX - e, 0);
    uint32_t f = ~(uint32_t)t & (uint32_t)(MLVB_EV_RES - 1);
    asm("" : "+r"(f));          // index * 2 + table base as one IMAD
#ifdef FW_SHR_HI
    return (uint32_t)FW_T13(f) >> (__umulhi((uint32_t)t, m17) & 31);      // t >> 15 on the FMA pipe
#else
    return (uint32_t)FW_T13(f) >> (((uint32_t)t >> 15) & 31);
#endif
}

// finish quad column C of the middle row: smoothed R/B (chroma_smooth.c:30-68), stripe gains, packed words
template <int STRIPES, int C>
__device__ __forceinline__ void wide_finish_col(const WideRow &M, int mr, int mb, int ge_thr, bool edge_first, bool edge_last,
                                                const WideConst &K, uint32_t &top, uint32_t &bot)
{
    uint32_t r = M.r[C], b = M.b[C];
    const int ge = M.ge[C];
    bool go = ge >= ge_thr;
    if (C < 2) go = go && !edge_first;
    if (C >= FW_COLS - 2) go = go && !edge_last;
    const int er = add2(mr, ge, K), eb = add2(mb, ge, K);
    go = go && er > MLVB_EV_RES && eb > MLVB_EV_RES;
    // branch-free: both lookups always run (indices clamped into the table), the result is selected -- no
    // reconvergence barrier between the columns of a lane, so their instruction streams interleave freely
    const uint32_t nr = (uint32_t)add2((int)wide_ev2raw(er, K.shr[1]), (int)K.black, K), nb = (uint32_t)add2((int)wide_ev2raw(eb, K.shr[1]), (int)K.black, K);
    r = go ? nr : r;
    b = go ? nb : b;
#ifdef FW_PACK_PRMT
    constexpr bool EH = wide_in_hi<STRIPES, (2 * C) & 7>(), OH = wide_in_hi<STRIPES, (2 * C + 1) & 7>();   // even / odd pixel column
    r = wide_gain_x<STRIPES, (2 * C) & 7>(r, K);
    b = wide_gain_x<STRIPES, (2 * C + 1) & 7>(b, K);
    top = wide_clamp2<STRIPES>(wide_pack<EH, OH>(r, M.g1s[C]), K);           // all four samples are < 2^16
    bot = wide_clamp2<STRIPES>(wide_pack<EH, OH>(M.g2[C], b), K);
#else
    r = wide_gain<STRIPES, (2 * C) & 7>(r, K);
    b = wide_gain<STRIPES, (2 * C + 1) & 7>(b, K);
    top = r | M.g1s[C];
    bot = M.g2[C] | (b << 16);
#endif
}

template <int C> struct ColTag { static constexpr int value = C; };

// One row step: take a new quad row from the staged bytes into N (overwriting the row two above, whose
// dr / db are still read for the column sorts), and write the finished middle row M when `emit`.
//   lane_off : byte offset of the lane's first word inside a staged row (4-byte aligned)
//   perm     : byte-permute selector turning two loaded words into one stream-order word
template <int STRIPES>
__device__ __forceinline__ void wide_step(WideRow &N, const WideRow &M, const uint8_t *stage_rows, int lane_off, uint32_t perm,
                                          bool emit, int ge_thr, bool edge_first, bool edge_last, bool writer, uint16_t *orow, int w,
                                          const WideConst &K)
{
    uint32_t T[FW_NW], B[FW_NW];
    if constexpr (FW_COLS == 8) {                   // 28 bytes per lane: always word aligned
#pragma unroll
        for (int j = 0; j < FW_NW; j++) {
            T[j] = __byte_perm(*reinterpret_cast<const uint32_t *>(stage_rows + lane_off + 4 * j), 0, 0x1032);
            B[j] = __byte_perm(*reinterpret_cast<const uint32_t *>(stage_rows + FW_ROWBYTES + lane_off + 4 * j), 0, 0x1032);
        }
    } else {                                        // 14 bytes per lane: every other lane starts in the middle of a word
        uint32_t lt[FW_NW + 1], lb[FW_NW + 1];
#pragma unroll
        for (int j = 0; j <= FW_NW; j++) {
            lt[j] = *reinterpret_cast<const uint32_t *>(stage_rows + lane_off + 4 * j);
            lb[j] = *reinterpret_cast<const uint32_t *>(stage_rows + FW_ROWBYTES + lane_off + 4 * j);
        }
#pragma unroll
        for (int j = 0; j < FW_NW; j++) {
            T[j] = __byte_perm(lt[j], lt[j + 1], perm);
            B[j] = __byte_perm(lb[j], lb[j + 1], perm);
        }
    }
    // sorted columns of (row above = N's old contents, M, new row); N is replaced column by column
    Tri tr[FW_COLS + 2], tb[FW_COLS + 2];            // index c + 1
    auto do_col = [&](auto cc) {
        constexpr int C = decltype(cc)::value;
        const int adr = N.dr[C], adb = N.db[C];
        wide_ingest_col<STRIPES, C>(T, B, K, N);
        tr[C + 1] = wide_sort3(adr, M.dr[C], N.dr[C], K);
        tb[C + 1] = wide_sort3(adb, M.db[C], N.db[C], K);
    };
    uint32_t top[FW_COLS], bot[FW_COLS];
    auto fin_col = [&](auto cc) {
        constexpr int C = decltype(cc)::value;
#ifdef FW_MIDPAIR
        // window C holds columns C - 1 .. C + 1 = tr[C .. C + 2]; the shared pair is tr[P], tr[P + 1] with P = (C | 1)
        constexpr int PP = C | 1, TH = (C & 1) ? C + 2 : C;
        const int mr = wide_med9_pair(tr[C], tr[C + 1], tr[C + 2], mid_pair(tr[PP], tr[PP + 1]), tr[TH].mid, K);
        const int mb = wide_med9_pair(tb[C], tb[C + 1], tb[C + 2], mid_pair(tb[PP], tb[PP + 1]), tb[TH].mid, K);
#else
        const int mr = wide_med9(tr[C], tr[C + 1], tr[C + 2], K);
        const int mb = wide_med9(tb[C], tb[C + 1], tb[C + 2], K);
#endif
        wide_finish_col<STRIPES, C>(M, mr, mb, ge_thr, edge_first, edge_last, K, top[C], bot[C]);
    };
    do_col(ColTag<0>{});
    do_col(ColTag<FW_COLS - 1>{});
    tr[0] = tri_shfl_up(tr[FW_COLS]);           tb[0] = tri_shfl_up(tb[FW_COLS]);
    tr[FW_COLS + 1] = tri_shfl_down(tr[1]);     tb[FW_COLS + 1] = tri_shfl_down(tb[1]);
    do_col(ColTag<1>{}); fin_col(ColTag<0>{});
    do_col(ColTag<2>{}); fin_col(ColTag<1>{});
    if constexpr (FW_COLS == 8) {
        do_col(ColTag<3>{}); fin_col(ColTag<2>{});
        do_col(ColTag<4>{}); fin_col(ColTag<3>{});
        if (emit && writer) {
            FW_STORE(reinterpret_cast<uint4 *>(orow), make_uint4(top[0], top[1], top[2], top[3]));
            FW_STORE(reinterpret_cast<uint4 *>(orow + w), make_uint4(bot[0], bot[1], bot[2], bot[3]));
        }
        do_col(ColTag<5>{}); fin_col(ColTag<4>{});
        do_col(ColTag<6>{}); fin_col(ColTag<5>{});
    }
    fin_col(ColTag<FW_COLS - 2>{});
    fin_col(ColTag<FW_COLS - 1>{});
    if (emit && writer) {
        constexpr int o = FW_COLS - 4;
        FW_STORE(reinterpret_cast<uint4 *>(orow + 2 * o), make_uint4(top[o], top[o + 1], top[o + 2], top[o + 3]));
        FW_STORE(reinterpret_cast<uint4 *>(orow + w + 2 * o), make_uint4(bot[o], bot[o + 1], bot[o + 2], bot[o + 3]));
    }
}

// write a repaired 14-bit sample into the staged stream bytes of one pixel row (16-bit LE words, MSB first)
__device__ __forceinline__ void wide_patch(uint8_t *row, int bit, uint32_t v)
{
    v &= 0x3FFFu;
    uint16_t *wds = reinterpret_cast<uint16_t *>(row);
    const int k = bit >> 4, s = bit & 15;
    uint32_t pair = (uint32_t)wds[k] << 16;
    if (s > 2) pair |= wds[k + 1];
    const uint32_t m = 0x3FFFu << (18 - s);
    pair = (pair & ~m) | (v << (18 - s));
    wds[k] = (uint16_t)(pair >> 16);
    if (s > 2) wds[k + 1] = (uint16_t)pair;
}

// applies the items of quad row `qrow` starting at index a (items are sorted by quad row); returns the index of the
// first item of a later row
__device__ __noinline__ unsigned wide_apply_patches(const WideItem *items, unsigned a, unsigned e, int qrow, const uint16_t *vals,
                                                    uint8_t *slot_rows, int bit0, bool writer_lane)
{
#pragma unroll 1
    for (; a < e; a++) {
        const WideItem it = items[a];
        if ((int)it.qrow != qrow) break;
        if (writer_lane) wide_patch(slot_rows + (it.px_sub >> 13) * FW_ROWBYTES, bit0 + (it.px_sub & 0xFFF) * 14, vals[it.entry]);
    }
    return a;
}

template <int STRIPES>
__global__ void __launch_bounds__(FW_THREADS, 1)
fused3_wide_kernel(const __grid_constant__ WideParams P)
{
    int *s_r2e = reinterpret_cast<int *>(fw_smem);
    uint16_t *s_t13 = reinterpret_cast<uint16_t *>(fw_smem + FW_SMEM_R2E);
    for (int i = threadIdx.x; i < 16384; i += FW_THREADS) s_r2e[i] = __ldg(P.raw2ev + i);
    for (int i = threadIdx.x; i < FW_SMEM_T13 / 16; i += FW_THREADS)
        reinterpret_cast<uint4 *>(s_t13)[i] = __ldg(reinterpret_cast<const uint4 *>(P.ev2raw13) + i);
#ifdef FW_STAGE_TMA
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(fw_smem + FW_SMEM_R2E + FW_SMEM_T13 + FW_SMEM_STAGE) + (threadIdx.x >> 5) * 16;
    if ((threadIdx.x & 31) == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    uint32_t phase = 0;                            // bit s = parity the next wait on slot s expects
#endif
    __syncthreads();

    WideConst K;
    K.black = (uint32_t)P.black16; K.thr = (uint32_t)P.black16 + 64u; K.white = (uint32_t)P.white16;
    K.white2 = K.white | (K.white << 16);
#pragma unroll
    for (int i = 0; i < 8; i++) { K.coef[i] = P.gain[i].coef; K.k1[i] = P.gain[i].k1; K.coefh[i] = P.coefh[i]; }
    K.k4 = P.k4; K.whitex = P.whitex;
#ifdef FW_GAIN_X
#pragma unroll
    for (int i = 0; i < 8; i++) { K.coef[i] = P.gainx[i].coef; K.kx[i] = P.gainx[i].kx; }
#endif
    K.one = P.one; K.mone = P.mone;
#pragma unroll
    for (int i = 0; i < 4; i++) K.shl[i] = P.shl[i];
    K.shr[0] = P.shr[0]; K.shr[1] = P.shr[1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stage = fw_smem + FW_SMEM_R2E + FW_SMEM_T13 + warp * FW_STAGE_PER_WARP;
    const int w = P.w, ph = P.h >> 1;
    const int rowbytes = (w * 7) >> 2;
    // Work is cut two ways.  nseg > 0: frames x strips x nseg equal segments dealt round-robin to the warps.
    // nseg == 0 (default): the strips are dealt to the warps round-robin (warp gw owns strip gw % nstrips, so the
    // warps of a CTA walk neighbouring strips in step and the bytes two strips share meet in L2), and the column
    // of nframes x ph quad rows of a strip is cut into one contiguous run per warp of that strip, split only at
    // frame ends -- every warp gets the same number of rows (no partial last round) and re-reads halo rows for two or
    // three pieces instead of one per segment.
    const bool runs = P.nseg == 0;
    const int gw = blockIdx.x * FW_WARPS + warp, nwarps = gridDim.x * FW_WARPS;
    int cur, end, run_strip = 0;
    if (runs) {
        run_strip = gw % P.nstrips;
        const int nw_s = (nwarps - run_strip + P.nstrips - 1) / P.nstrips;      // warps that own this strip
        const int col_rows = P.nframes * ph;
        const int per = (col_rows + nw_s - 1) / nw_s;
        cur = min((gw / P.nstrips) * per, col_rows);
        end = min(cur + per, col_rows);
    } else {
        cur = gw;
        end = P.nframes * P.nstrips * P.nseg;
    }

    while (cur < end) {
        int strip, frame_i, qr0, qr1;
        if (runs) {
            strip = run_strip; frame_i = cur / ph; qr0 = cur - frame_i * ph; qr1 = min(qr0 + (end - cur), ph);
            cur += qr1 - qr0;
        } else {
            const int t0 = cur / P.nseg;
            qr0 = (cur - t0 * P.nseg) * P.seg_rows; qr1 = min(qr0 + P.seg_rows, ph);
            strip = t0 % P.nstrips; frame_i = t0 / P.nstrips;
            cur += nwarps;
        }
        const uint8_t *frame = P.packed + (size_t)frame_i * P.payload_stride;
        uint16_t *out = P.out + (size_t)frame_i * P.out_stride;
        const int xl = strip * FW_STRIP_PX - FW_LANE_PX + FW_LANE_PX * lane;   // first pixel column of this lane
        const bool lane_ok = xl >= 0 && xl < w;
        const bool writer = lane_ok && lane >= 1 && lane <= 30;
        const bool edge_first = xl == 0, edge_last = xl + FW_LANE_PX == w;
        const int byte0 = (strip * FW_STRIP_PX - FW_LANE_PX) * 14 / 8;      // lane 0's first stream byte inside a row
        const int base16 = byte0 & ~15;
        const int lane_byte = (byte0 - base16) + FW_LANE_BYTES * lane;
        const int lane_off = lane_byte & ~3;
        const uint32_t perm = (lane_byte & 2) ? 0x3254u : 0x1032u;
        const unsigned *row_start = P.items ? P.row_start + (size_t)strip * (ph + 1) : nullptr;
        const uint16_t *vals = P.vals + (size_t)frame_i * P.n_entries;

#ifndef FW_STAGE_TMA
        // rows are requested in order (qr0-1, qr0, ...): a running source pointer per 16-byte chunk of the lane
        constexpr int NCH = (FW_ROWBYTES / 16 + 31) / 32;
        bool pf_ok[NCH];
        const uint8_t *pf_src[NCH];
#pragma unroll
        for (int k = 0; k < NCH; k++) {
            const int ch = lane + 32 * k, src = base16 + 16 * ch;
            pf_ok[k] = ch < FW_ROWBYTES / 16 && src >= 0 && src + 16 <= rowbytes;
            pf_src[k] = frame + (ptrdiff_t)(2 * (qr0 - 1)) * rowbytes + src;
        }
        auto prefetch = [&](int qr, int slot) {
            const bool row_ok = (unsigned)qr < (unsigned)ph && qr <= qr1;
#pragma unroll
            for (int k = 0; k < NCH; k++) {
                if (pf_ok[k] && row_ok) {
                    uint8_t *dst = stage + slot * 2 * FW_ROWBYTES + 16 * (lane + 32 * k);
                    cp_async16(dst, pf_src[k]);
                    cp_async16(dst + FW_ROWBYTES, pf_src[k] + rowbytes);
                }
                pf_src[k] += 2 * rowbytes;
            }
            cp_async_commit();
        };
        auto landed = [&](int qr, int slot) { cp_async_wait1(); __syncwarp(); };
#else
        // the strip's bytes of a row, cut to the row: [lo, hi) in 16-byte units relative to base16
        const int lo = max(-base16, 0), hi = min(FW_ROWBYTES, rowbytes - base16);
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
        auto prefetch = [&](int qr, int slot) {
            if (qr >= 0 && qr < ph && lane == 0) {
                const uint8_t *src = frame + (size_t)(2 * qr) * rowbytes + base16 + lo;
                const uint32_t dst = stage_s + slot * 2 * FW_ROWBYTES + lo, bar = bar0 + 8 * slot;
                mbar_expect_tx(bar, 2u * (uint32_t)(hi - lo));
                tma_bulk_g2s(dst, src, (uint32_t)(hi - lo), bar);
                tma_bulk_g2s(dst + FW_ROWBYTES, src + rowbytes, (uint32_t)(hi - lo), bar);
            }
        };
        auto landed = [&](int qr, int slot) {
            if (qr >= 0 && qr < ph) {
                mbar_wait(bar0 + 8 * slot, (phase >> slot) & 1u);
                phase ^= 1u << slot;
            }
        };
#endif
        // repaired pixels of this segment: items [pc, pe) sorted by quad row; only the row of the next pending item is
        // kept in a register, so a row step costs one compare unless it has a patch
        unsigned pc = 0, pe = 0;
        if (FW_PATCHES && row_start) { pc = __ldg(row_start + max(qr0 - 1, 0)); pe = __ldg(row_start + min(qr1, ph - 1) + 1); }
        int next_patch_row = pc < pe ? (int)P.items[pc].qrow : 0x7FFFFFFF;
        auto arrive = [&](int qr, int slot) {                                // staged bytes of quad row qr are visible after this
            landed(qr, slot);
            if (qr == next_patch_row) {
                pc = wide_apply_patches(P.items, pc, pe, qr, vals, stage + slot * 2 * FW_ROWBYTES, (byte0 - base16) * 8, lane == 0);
                next_patch_row = pc < pe ? (int)P.items[pc].qrow : 0x7FFFFFFF;
#ifdef FW_STAGE_TMA
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");           // generic writes before the next bulk copy into this slot
#endif
                __syncwarp();
            }
        };

        // chroma_smooth.c:26: rows y = 2 * qr with 4 <= y < h - 5, i.e. quad rows 2 .. 2 + smooth_rows - 1
        const unsigned smooth_rows = (unsigned)max((P.h - 5 + 1) / 2 - 2, 0);
        WideRow R0, R1;
#pragma unroll
        for (int c = 0; c < FW_COLS; c++) { R0.dr[c] = R0.db[c] = R1.dr[c] = R1.db[c] = 0; }
        prefetch(qr0 - 1, 0);
        prefetch(qr0, 1);
        uint16_t *orow = out + (size_t)(2 * qr0) * w + xl;
        // rows enter in the order qr0-1, qr0, ..., qr1; row q is finished when row q+1 has entered
        int q = qr0 - 1;
        arrive(q, 0);
        wide_step<STRIPES>(R0, R1, stage, lane_off, perm, false, 0, false, false, false, orow, w, K);
        __syncwarp();
        prefetch(q + 2, 0);
        q++;
        arrive(q, 1);
        wide_step<STRIPES>(R1, R0, stage + 2 * FW_ROWBYTES, lane_off, perm, false, 0, false, false, false, orow, w, K);
        __syncwarp();
        prefetch(q + 2, 1);
        q++;
        // steady state: row q enters, row q-1 is written
        while (q <= qr1) {
            {
                const int thr = (unsigned)(q - 3) < smooth_rows ? 2 * MLVB_EV_RES : 0x7FFFFFFF;   // quad row q-1 is smoothed
                arrive(q, 0);
                wide_step<STRIPES>(R0, R1, stage, lane_off, perm, true, thr, edge_first, edge_last, writer, orow, w, K);
                __syncwarp();
                prefetch(q + 2, 0);
                orow += 2 * w;
                q++;
            }
            if (q > qr1) break;
            {
                const int thr = (unsigned)(q - 3) < smooth_rows ? 2 * MLVB_EV_RES : 0x7FFFFFFF;
                arrive(q, 1);
                wide_step<STRIPES>(R1, R0, stage + 2 * FW_ROWBYTES, lane_off, perm, true, thr, edge_first, edge_last, writer, orow, w, K);
                __syncwarp();
                prefetch(q + 2, 1);
                orow += 2 * w;
                q++;
            }
        }
#ifndef FW_STAGE_TMA
        cp_async_wait0();
#endif
        __syncwarp();
    }
}

}  // namespace
