// fused_wide.cuh -- batch form of the fused single-ISO chain (unpack + bad-pixel patches + 3x3 median
// chroma smoothing + stripe gains, one pass over HBM), written for the instruction roofline that bounds
// the strip kernel in fused.cu (profiles/r01d_fused3_ncu.md):
//
//   * one persistent 512-thread CTA per SM; both EV tables live in shared memory: raw2ev for this black
//     level (16384 x int32, 64 KB) and the top octave of ev2raw (32768 x uint16, 64 KB) from which every
//     other octave is a right shift (ev2raw[e] == ev2raw[13 EV + e mod EV] >> (13 - e / EV), verified
//     entry by entry when the context is created) -- no table gather goes through the L1 tag stage;
//   * a lane owns EIGHT adjacent RGGB quad columns (16 pixels = 28 stream bytes per row), so every bit
//     offset is a compile-time constant (one shift + one mask per pixel), neighbouring sorted columns are
//     already in the lane's registers (12 shuffles per 8 quads instead of per quad) and the stripe gains
//     are constant-bank operands;
//   * the packed rows are staged by cp.async (16-byte chunks, L2 only) into a per-warp double buffer, one
//     quad row ahead; bad-pixel patches are written into the staged bytes before extraction;
//   * sort3 / med3 are evaluated as (min3, max3, a + b + c - min3 - max3): the additions can issue on the
//     FMA pipe (IMAD) while the min/max run on the ALU pipe, instead of 6 / 4 ALU-pipe min/max.
//
// Arithmetic is the same 32-bit wrap-around integer arithmetic as fused3_strip_kernel / the reference
// (cs.c:49-84, chroma_smooth.c:22-71, stripes.c:250-266); results are bit-identical.
#pragma once

namespace {

constexpr int FW_WARPS = 12;
constexpr int FW_THREADS = FW_WARPS * 32;
constexpr int FW_COLS = 8;                       // quad columns per lane
constexpr int FW_ROWBYTES = 960;                 // staged bytes per pixel row and warp (60 chunks of 16 B)
constexpr int FW_STAGE_PER_WARP = 4 * FW_ROWBYTES;   // 2 slots x 2 pixel rows
constexpr int FW_SMEM_R2E = 16384 * 4;
constexpr int FW_SMEM_T13 = 32768 * 2;
constexpr int FW_SMEM_BYTES = FW_SMEM_R2E + FW_SMEM_T13 + FW_WARPS * FW_STAGE_PER_WARP;

extern __shared__ __align__(16) uint8_t fw_smem[];
#define FW_R2E(v) (reinterpret_cast<const int *>(fw_smem)[(v)])
#define FW_T13(f) (reinterpret_cast<const uint16_t *>(fw_smem + FW_SMEM_R2E)[(f)])

struct WideItem { unsigned short px; unsigned short sub; unsigned entry; };   // px: pixel column inside the strip's 512-pixel window

struct WideParams {
    const uint8_t *packed; size_t payload_stride;
    uint16_t *out; size_t out_stride;
    int w, h, black;
    const int *raw2ev;                // indexed by raw value (16384 entries)
    const uint16_t *ev2raw13;         // ev2raw[13 EV ...], 32768 entries
    int black16, white16;
    unsigned coef[8];
    const WideItem *items; const unsigned *row_start; const uint16_t *vals; unsigned n_entries;
    int nstrips, nseg, seg_rows, nframes;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ int imin3(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
// exact in wrap-around arithmetic: the three values are a permutation of (min, med, max)
__device__ __forceinline__ int imed3(int a, int b, int c)
{
    const unsigned s = (unsigned)a + (unsigned)b + (unsigned)c;
    return (int)(s - (unsigned)imin3(a, b, c) - (unsigned)imax3(a, b, c));
}

// pixel I (0..15) of a lane's 16-pixel group; W[] are the group's seven 32-bit words in stream order
template <int I>
__device__ __forceinline__ uint32_t wide_px(const uint32_t (&W)[7])
{
    constexpr int bit = 14 * I, j = bit >> 5, s = bit & 31;
    if constexpr (s <= 18) return (W[j] >> (18 - s)) & 0x3FFFu;
    else return __funnelshift_l(W[j + 1], W[j], s) >> 18;
}

struct WideRow {                                  // per-lane state of one quad row (8 quad columns)
    int dr[FW_COLS], db[FW_COLS];                 // ev(r) - ge, ev(b) - ge
    int ge[FW_COLS];
    uint32_t r[FW_COLS], b[FW_COLS];              // raw R / B samples
    uint32_t g1s[FW_COLS], g2[FW_COLS];           // finished G1 << 16, finished G2
};

struct WideConst { uint32_t black, thr, white; uint32_t coef[8]; };

template <bool STRIPES, int IDX>
__device__ __forceinline__ uint32_t wide_gain(uint32_t v, const WideConst &K)
{
    if (STRIPES) {
        if (v > K.thr) {                                                       // stripes.c:258: v > black + 64
            const uint32_t t = (((v - K.black) * K.coef[IDX]) >> 16) + K.black;   // product < 2^32: coef < 2^18 (host check)
            return min(t, K.white);
        }
    }
    return v;
}

template <bool STRIPES, int C>
__device__ __forceinline__ void wide_ingest_col(const uint32_t (&T)[7], const uint32_t (&B)[7], const int *s_r2e,
                                                const WideConst &P, WideRow &R)
{
    const uint32_t r = wide_px<2 * C>(T), g1 = wide_px<2 * C + 1>(T);
    const uint32_t g2 = wide_px<2 * C>(B), b = wide_px<2 * C + 1>(B);
    const int ge = wadd(FW_R2E(g1), FW_R2E(g2)) / 2;
    R.ge[C] = ge;
    R.dr[C] = wsub(FW_R2E(r), ge);
    R.db[C] = wsub(FW_R2E(b), ge);
    R.r[C] = r;
    R.b[C] = b;
    R.g1s[C] = wide_gain<STRIPES, (2 * C + 1) & 7>(g1, P) << 16;
    R.g2[C] = wide_gain<STRIPES, (2 * C) & 7>(g2, P);
}

struct Tri { int lo, mid, hi; };

__device__ __forceinline__ Tri wide_sort3(int a, int b, int c)
{
    Tri t;
    t.lo = imin3(a, b, c);
    t.hi = imax3(a, b, c);
    t.mid = (int)((unsigned)a + (unsigned)b + (unsigned)c - (unsigned)t.lo - (unsigned)t.hi);
    return t;
}
__device__ __forceinline__ int wide_med9(const Tri &l, const Tri &m, const Tri &r)
{
    return imed3(imax3(l.lo, m.lo, r.lo), imed3(l.mid, m.mid, r.mid), imin3(l.hi, m.hi, r.hi));
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

// finish quad column C of the middle row: smoothed R/B (chroma_smooth.c:30-68), stripe gains, packed words
template <bool STRIPES, int C>
__device__ __forceinline__ void wide_finish_col(const WideRow &M, int mr, int mb, int ge_thr, bool edge_first, bool edge_last,
                                                const uint16_t *s_t13, const WideConst &P, uint32_t &top, uint32_t &bot)
{
    uint32_t r = M.r[C], b = M.b[C];
    const int ge = M.ge[C];
    bool go = ge >= ge_thr;
    if (C < 2) go = go && !edge_first;
    if (C >= FW_COLS - 2) go = go && !edge_last;
    const int er = wadd(ge, mr), eb = wadd(ge, mb);
    if (go && er > MLVB_EV_RES && eb > MLVB_EV_RES) {
        const int cr = min(er, MLVB_EV_MAX), cb = min(eb, MLVB_EV_MAX);
        r = ((uint32_t)FW_T13(cr & (MLVB_EV_RES - 1)) >> (13 - (cr >> 15))) + P.black;
        b = ((uint32_t)FW_T13(cb & (MLVB_EV_RES - 1)) >> (13 - (cb >> 15))) + P.black;
    }
    r = wide_gain<STRIPES, (2 * C) & 7>(r, P);
    b = wide_gain<STRIPES, (2 * C + 1) & 7>(b, P);
    top = r | M.g1s[C];
    bot = M.g2[C] | (b << 16);
}

// One row step: take quad row `qr_new` from the staged bytes into N (overwriting the row two above), and
// write the finished middle row M (quad row qr_new - 1) when `emit`.
template <bool STRIPES>
__device__ __forceinline__ void wide_step(WideRow &N, const WideRow &M, const uint8_t *stage_rows, int lane_off, bool have_new,
                                          bool emit, int ge_thr, bool edge_first, bool edge_last, bool writer, uint16_t *orow, int w,
                                          const int *s_r2e, const uint16_t *s_t13, const WideConst &P)
{
    uint32_t T[7], B[7];
    if (have_new) {
#pragma unroll
        for (int j = 0; j < 7; j++) {
            T[j] = __byte_perm(*reinterpret_cast<const uint32_t *>(stage_rows + lane_off + 4 * j), 0, 0x1032);
            B[j] = __byte_perm(*reinterpret_cast<const uint32_t *>(stage_rows + FW_ROWBYTES + lane_off + 4 * j), 0, 0x1032);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 7; j++) T[j] = B[j] = 0u;
    }
    // sorted columns of (row above = N's old contents, M, new row); N is replaced column by column
    Tri tr[FW_COLS + 2], tb[FW_COLS + 2];            // index c + 1
    auto do_col = [&](auto cc) {
        constexpr int C = decltype(cc)::value;
        const int adr = N.dr[C], adb = N.db[C];
        wide_ingest_col<STRIPES, C>(T, B, s_r2e, P, N);
        tr[C + 1] = wide_sort3(adr, M.dr[C], N.dr[C]);
        tb[C + 1] = wide_sort3(adb, M.db[C], N.db[C]);
    };
    do_col(std::integral_constant<int, 0>{});
    do_col(std::integral_constant<int, 7>{});
    tr[0] = tri_shfl_up(tr[8]);   tb[0] = tri_shfl_up(tb[8]);
    tr[9] = tri_shfl_down(tr[1]); tb[9] = tri_shfl_down(tb[1]);
    uint32_t top[FW_COLS], bot[FW_COLS];
    auto fin_col = [&](auto cc) {
        constexpr int C = decltype(cc)::value;
        const int mr = wide_med9(tr[C], tr[C + 1], tr[C + 2]);
        const int mb = wide_med9(tb[C], tb[C + 1], tb[C + 2]);
        wide_finish_col<STRIPES, C>(M, mr, mb, ge_thr, edge_first, edge_last, s_t13, P, top[C], bot[C]);
    };
    do_col(std::integral_constant<int, 1>{}); fin_col(std::integral_constant<int, 0>{});
    do_col(std::integral_constant<int, 2>{}); fin_col(std::integral_constant<int, 1>{});
    do_col(std::integral_constant<int, 3>{}); fin_col(std::integral_constant<int, 2>{});
    do_col(std::integral_constant<int, 4>{}); fin_col(std::integral_constant<int, 3>{});
    if (emit && writer) {
        __stcs(reinterpret_cast<uint4 *>(orow), make_uint4(top[0], top[1], top[2], top[3]));
        __stcs(reinterpret_cast<uint4 *>(orow + w), make_uint4(bot[0], bot[1], bot[2], bot[3]));
    }
    do_col(std::integral_constant<int, 5>{}); fin_col(std::integral_constant<int, 4>{});
    do_col(std::integral_constant<int, 6>{}); fin_col(std::integral_constant<int, 5>{});
    fin_col(std::integral_constant<int, 6>{});
    fin_col(std::integral_constant<int, 7>{});
    if (emit && writer) {
        __stcs(reinterpret_cast<uint4 *>(orow) + 1, make_uint4(top[4], top[5], top[6], top[7]));
        __stcs(reinterpret_cast<uint4 *>(orow + w) + 1, make_uint4(bot[4], bot[5], bot[6], bot[7]));
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

__device__ __noinline__ void wide_apply_patches(const WideItem *items, unsigned a, unsigned e, const uint16_t *vals, uint8_t *slot_rows,
                                                int bit0)
{
#pragma unroll 1
    for (unsigned i = a; i < e; i++) {
        const WideItem it = items[i];
        wide_patch(slot_rows + (it.sub >> 1) * FW_ROWBYTES, bit0 + it.px * 14, vals[it.entry]);
    }
}

template <bool STRIPES>
__global__ void __launch_bounds__(FW_THREADS, 1)
fused3_wide_kernel(const __grid_constant__ WideParams P)
{
    int *s_r2e = reinterpret_cast<int *>(fw_smem);
    uint16_t *s_t13 = reinterpret_cast<uint16_t *>(fw_smem + FW_SMEM_R2E);
    for (int i = threadIdx.x; i < 16384; i += FW_THREADS) s_r2e[i] = __ldg(P.raw2ev + i);
    for (int i = threadIdx.x; i < FW_SMEM_T13 / 16; i += FW_THREADS)
        reinterpret_cast<uint4 *>(s_t13)[i] = __ldg(reinterpret_cast<const uint4 *>(P.ev2raw13) + i);
    __syncthreads();

    WideConst K;
    K.black = (uint32_t)P.black16; K.thr = (uint32_t)P.black16 + 64u; K.white = (uint32_t)P.white16;
#pragma unroll
    for (int i = 0; i < 8; i++) K.coef[i] = P.coef[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stage = fw_smem + FW_SMEM_R2E + FW_SMEM_T13 + warp * FW_STAGE_PER_WARP;
    const int w = P.w, ph = P.h >> 1;
    const int rowbytes = (w * 7) >> 2;
    const int total = P.nframes * P.nstrips * P.nseg;

    for (int item = blockIdx.x * FW_WARPS + warp; item < total; item += gridDim.x * FW_WARPS) {
        const int seg = item % P.nseg;
        const int t0 = item / P.nseg;
        const int strip = t0 % P.nstrips, frame_i = t0 / P.nstrips;
        const uint8_t *frame = P.packed + (size_t)frame_i * P.payload_stride;
        uint16_t *out = P.out + (size_t)frame_i * P.out_stride;
        const int xl = strip * 480 - 16 + 16 * lane;                       // first pixel column of this lane
        const bool lane_ok = xl >= 0 && xl < w;
        const bool writer = lane_ok && lane >= 1 && lane <= 30;
        const bool edge_first = xl == 0, edge_last = xl + 16 == w;
        const int byte0 = strip * 840 - 28;                                // lane 0's first stream byte inside a row
        const int base16 = byte0 & ~15;
        const int lane_off = (byte0 - base16) + 28 * lane;
        const int qr0 = seg * P.seg_rows, qr1 = min(qr0 + P.seg_rows, ph);
        const unsigned *row_start = P.items ? P.row_start + (size_t)strip * (ph + 1) : nullptr;
        const uint16_t *vals = P.vals + (size_t)frame_i * P.n_entries;

        auto prefetch = [&](int qr, int slot) {
            if (qr >= 0 && qr < ph) {
                const uint8_t *src_row = frame + (size_t)(2 * qr) * rowbytes;
                uint8_t *dst = stage + slot * 2 * FW_ROWBYTES;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int ch = lane + 32 * k;
                    const int src = base16 + 16 * ch;
                    if (ch < FW_ROWBYTES / 16 && src >= 0 && src + 16 <= rowbytes) {
                        cp_async16(dst + 16 * ch, src_row + src);
                        cp_async16(dst + FW_ROWBYTES + 16 * ch, src_row + rowbytes + src);
                    }
                }
            }
            cp_async_commit();
        };
        auto arrive = [&](int qr, int slot) {                                // staged bytes of quad row qr are visible after this
            cp_async_wait1();
            __syncwarp();
            if (row_start && qr >= 0 && qr < ph) {
                const unsigned a = __ldg(row_start + qr), e = __ldg(row_start + qr + 1);
                if (a < e) {
                    if (lane == 0) wide_apply_patches(P.items, a, e, vals, stage + slot * 2 * FW_ROWBYTES, (byte0 - base16) * 8);
                    __syncwarp();
                }
            }
        };

        WideRow R0, R1;
#pragma unroll
        for (int c = 0; c < FW_COLS; c++) { R0.dr[c] = R0.db[c] = R1.dr[c] = R1.db[c] = 0; }
        prefetch(qr0 - 1, 0);
        prefetch(qr0, 1);
        uint16_t *orow = out + (size_t)(2 * qr0) * w + xl;
        // rows enter in the order qr0-1, qr0, ..., qr1; row q is finished when row q+1 has entered
        int q = qr0 - 1;
        // prologue: two rows without output
        arrive(q, 0);
        wide_step<STRIPES>(R0, R1, stage, lane_off, q >= 0, false, 0, false, false, false, orow, w, s_r2e, s_t13, K);
        __syncwarp();
        prefetch(q + 2, 0);
        q++;
        arrive(q, 1);
        wide_step<STRIPES>(R1, R0, stage + 2 * FW_ROWBYTES, lane_off, true, false, 0, false, false, false, orow, w, s_r2e, s_t13, K);
        __syncwarp();
        prefetch(q + 2, 1);
        q++;
        // steady state: row q enters, row q-1 is written
        while (q <= qr1) {
            {
                const int y = 2 * (q - 1);
                const int thr = (y >= 4 && y < P.h - 5) ? 2 * MLVB_EV_RES : 0x7FFFFFFF;
                arrive(q, 0);
                wide_step<STRIPES>(R0, R1, stage, lane_off, q < ph, true, thr, edge_first, edge_last, writer, orow, w, s_r2e, s_t13, K);
                __syncwarp();
                prefetch(q + 2 <= qr1 ? q + 2 : -1, 0);
                orow += 2 * w;
                q++;
            }
            if (q > qr1) break;
            {
                const int y = 2 * (q - 1);
                const int thr = (y >= 4 && y < P.h - 5) ? 2 * MLVB_EV_RES : 0x7FFFFFFF;
                arrive(q, 1);
                wide_step<STRIPES>(R1, R0, stage + 2 * FW_ROWBYTES, lane_off, q < ph, true, thr, edge_first, edge_last, writer, orow, w, s_r2e,
                                   s_t13, K);
                __syncwarp();
                prefetch(q + 2 <= qr1 ? q + 2 : -1, 1);
                orow += 2 * w;
                q++;
            }
        }
        cp_async_wait0();
        __syncwarp();
    }
}

}  // namespace
