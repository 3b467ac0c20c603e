// lj92.cu -- lossless-JPEG ("LJ92") VIDF payload -> 16-bit frame, decoded in parallel INSIDE a frame.
//
// Replaces reference lj92.c:650-702 (lj92_open / lj92_decode) and the quadrant de-interleave loop of
// main.c:656-668.  The reference (and the first version of this file) walks a frame's scan serially;
// the kernels below run the cooperative programs of lj92_core.cuh (see the header there for the
// algorithm): unstuff -> self-synchronising Huffman decode over 1024-bit subsequences -> prefix sums ->
// difference write -> wavefront prediction -> untile.  Ten launches per batch of frames, every stage
// batched over the frames with blockIdx.y.
//
// HBM traffic per frame (bytes; P = payload, N = pixels): read P twice and write the clean stream (3P),
// read the clean stream three times (mostly L2 hits: 3P), differences write + read + sample write (6N),
// untile read + write (4N).  Algorithmic bytes are P + 2N.
#include <algorithm>

#include "kernels.cuh"
#include "lj92_core.cuh"

namespace {

using namespace lj92;

constexpr int PREDICT_WARPS = 8;

struct DevCtx {
    int tid, nthr;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ int sync_or(int p) { return __syncthreads_or(p); }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ int all(int p) { return __all_sync(0xFFFFFFFFu, p); }
    __device__ __forceinline__ uint32_t atomic_add(uint32_t *p, uint32_t v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void pause() { __nanosleep(32); }
    __device__ __forceinline__ uint32_t shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xFFFFFFFFu, v, d); }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
    __device__ __forceinline__ void set_status(int *p, int v) { atomicCAS(p, 0, v); }
    __device__ __forceinline__ void load64(const uint16_t *p, uint32_t nx[16])
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 v = __ldcs(q + i);
            nx[4 * i] = v.x; nx[4 * i + 1] = v.y; nx[4 * i + 2] = v.z; nx[4 * i + 3] = v.w;
        }
    }
    __device__ __forceinline__ void store64(uint16_t *p, const uint32_t o[16])
    {
        uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    __device__ __forceinline__ void store16(uint16_t *p, const uint32_t w[4])
    {
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ __forceinline__ uint32_t load_volatile32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
    __device__ __forceinline__ void store_volatile32(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }
};

// ---- 0. headers ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
lj92_parse_kernel(const uint8_t *__restrict__ payload_base, size_t payload_stride, size_t payload_bytes,
                  char *scratch, Layout L, int W, int H)
{
    const int frame = blockIdx.x, lane = threadIdx.x;
    FrameWork F = frame_work(scratch, L, frame);
    if (lane == 0) {
        // payload = uint32 stored size, then the JPEG stream (main.c:626-629)
        const uint8_t *data = payload_base + (size_t)frame * payload_stride + 4;
        const int len = (int)min(payload_bytes - 4, (size_t)0x1FFFFFF0);
        int rc = parse_headers(data, len, *F.T);
        if (rc == ST_OK && (long long)F.T->lw * F.T->lh != (long long)W * H) rc = ST_HEADER;
        F.T->status = rc;
    }
    __syncwarp();
    if (F.T->status != ST_OK) return;
    for (int i = lane; i < (1 << LUT_BITS); i += 32) F.T->lut[i] = lut_entry(*F.T, i);
    for (int i = lane; i < (1 << LUT1_BITS); i += 32) F.T->lut1[i] = lut_entry(*F.T, i, LUT1_BITS);
}

// ---- 1. unstuff ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned load16_mask(const uint8_t *pl, size_t payload_bytes, const Tables &T, size_t q, uint8_t b[16])
{
    if (q + 16 <= payload_bytes && (((uintptr_t)(pl + q)) & 15) == 0) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pl + q));
        memcpy(b, &v, 16);
    } else {
        for (int j = 0; j < 16; j++) b[j] = q + j < payload_bytes ? pl[q + j] : (uint8_t)0;
    }
    return keep_mask16(pl, b, (long long)q, 4ll + T.scan_off, (long long)payload_bytes);
}

__device__ __forceinline__ unsigned block_sum_u32(unsigned v, unsigned *s_warp)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned t = 0;
    for (unsigned i = 0; i < blockDim.x / 32; i++) t += s_warp[i];
    return t;
}

__global__ void __launch_bounds__(UNSTUFF_THREADS)
lj92_unstuff_count_kernel(const uint8_t *__restrict__ payload_base, size_t payload_stride, size_t payload_bytes,
                          char *scratch, Layout L)
{
    __shared__ unsigned s_warp[UNSTUFF_THREADS / 32];
    FrameWork F = frame_work(scratch, L, blockIdx.y);
    if (F.T->status != ST_OK) return;
    const uint8_t *pl = payload_base + (size_t)blockIdx.y * payload_stride;
    const size_t q = ((size_t)blockIdx.x * UNSTUFF_THREADS + threadIdx.x) * 16;
    uint8_t b[16];
    const unsigned m = q < payload_bytes ? load16_mask(pl, payload_bytes, *F.T, q, b) : 0u;
    const unsigned tot = block_sum_u32(__popc(m), s_warp);
    if (threadIdx.x == 0) F.chunk_cnt[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
lj92_unstuff_scan_kernel(char *scratch, Layout L)
{
    __shared__ unsigned s_scan[32];
    __shared__ unsigned s_carry;
    FrameWork F = frame_work(scratch, L, blockIdx.x);
    if (F.T->status != ST_OK) return;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (unsigned b0 = 0; b0 < L.raw_chunks; b0 += blockDim.x) {
        const unsigned i = b0 + threadIdx.x;
        const unsigned v = i < L.raw_chunks ? F.chunk_cnt[i] : 0u;
        unsigned inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned w = s_scan[lane];
            unsigned winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (lane >= o) winc += t;
            }
            s_scan[lane] = winc - w;
        }
        __syncthreads();
        const unsigned carry = s_carry;
        const unsigned ex = carry + s_scan[wid] + inc - v;
        if (i < L.raw_chunks) F.chunk_cnt[i] = ex;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { F.chunk_cnt[L.raw_chunks] = s_carry; F.T->clean_bytes = s_carry; }
}

__global__ void __launch_bounds__(UNSTUFF_THREADS)
lj92_unstuff_scatter_kernel(const uint8_t *__restrict__ payload_base, size_t payload_stride, size_t payload_bytes,
                            char *scratch, Layout L)
{
    __shared__ unsigned s_warp[UNSTUFF_THREADS / 32];
    __shared__ __align__(16) uint8_t s_out[UNSTUFF_CHUNK + 32];
    FrameWork F = frame_work(scratch, L, blockIdx.y);
    if (F.T->status != ST_OK) return;
    const uint8_t *pl = payload_base + (size_t)blockIdx.y * payload_stride;
    const size_t q = ((size_t)blockIdx.x * UNSTUFF_THREADS + threadIdx.x) * 16;
    uint8_t b[16];
    const unsigned m = q < payload_bytes ? load16_mask(pl, payload_bytes, *F.T, q, b) : 0u;
    const unsigned cnt = __popc(m), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    // the block's kept bytes are contiguous in the clean stream: stage them in shared memory at the
    // same 16-byte phase as their destination, then copy out with aligned 128-bit stores
    const unsigned base = F.chunk_cnt[blockIdx.x], a = base & 15u;
    unsigned local = a + inc - cnt, total = 0;
    for (unsigned i = 0; i < UNSTUFF_THREADS / 32; i++) { if (i < wid) local += s_warp[i]; total += s_warp[i]; }
    if (m == 0xFFFFu && (local & 3u) == 0) {
        uint32_t w[4];
        memcpy(w, b, 16);
        uint32_t *d = reinterpret_cast<uint32_t *>(s_out + local);
        d[0] = w[0]; d[1] = w[1]; d[2] = w[2]; d[3] = w[3];
    } else {
        uint8_t *d = s_out + local;
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (m >> j & 1) *d++ = b[j];
    }
    __syncthreads();
    uint8_t *dst = F.clean + (base - a);                                       // 16-byte aligned
    const unsigned lo_all = a, hi_all = a + total;
    for (unsigned v = threadIdx.x; v * 16 < hi_all; v += blockDim.x) {
        const unsigned lo = v * 16, hi = lo + 16;
        if (lo >= lo_all && hi <= hi_all) *reinterpret_cast<uint4 *>(dst + lo) = *reinterpret_cast<const uint4 *>(s_out + lo);
        else
            for (unsigned j = max(lo, lo_all); j < min(hi, hi_all); j++) dst[j] = s_out[j];
    }
}

// ---- 2-4. Huffman decode ------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(DEC_THREADS)
lj92_dec_kernel(char *scratch, Layout L, uint32_t npix)
{
    __shared__ DecShared S;
    DevCtx C{(int)threadIdx.x, (int)blockDim.x};
    FrameWork F = frame_work(scratch, L, blockIdx.y);
    if (MODE == 2) clear_boundary(C, F, L, blockIdx.x, gridDim.x);             // for the prediction launch that follows
    dec_body<MODE>(C, S, F, blockIdx.x, npix);
}

__global__ void __launch_bounds__(DEC_THREADS)
lj92_resolve_kernel(char *scratch, Layout L, uint32_t npix)
{
    __shared__ DecShared S;
    DevCtx C{(int)threadIdx.x, (int)blockDim.x};
    FrameWork F = frame_work(scratch, L, blockIdx.x);
    resolve_body(C, S, F, npix);
}

// ---- 5. prediction ------------------------------------------------------------------------------
__global__ void __launch_bounds__(PREDICT_WARPS * 32)
lj92_predict_kernel(char *scratch, Layout L, int W, int geo_ok)
{
    __shared__ uint16_t s_ring[PREDICT_WARPS * RING_ELEMS];
    __shared__ uint32_t s_ticket;
    DevCtx C{(int)threadIdx.x, (int)blockDim.x};
    FrameWork F = frame_work(scratch, L, blockIdx.y);
    predict_body(C, s_ring, &s_ticket, F, geo_ok != 0, W);
}

// ---- 5b. predictor 6, separated (lj92_core.cuh): row recurrence, then column prefix sums fused with the
// untile of main.c:656-668 -------------------------------------------------------------------------
constexpr int ROW_WARPS = 4;

__device__ __forceinline__ void row_block64(uint4 (&v)[8], int &U, bool row0)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            U = row_step(row0, U, (int)(int16_t)(w[j] & 0xFFFFu));
            const uint32_t lo = (uint32_t)U & 0xFFFFu;
            U = row_step(row0, U, (int)(int16_t)(w[j] >> 16));
            w[j] = lo | ((uint32_t)U << 16);
        }
        v[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// one lane per row: U[r][c] = (U[r][c-1] >> 1) + d[r][c] in place, 64 columns (one 128-byte line) a time
__global__ void __launch_bounds__(ROW_WARPS * 32)
lj92_row_kernel(char *scratch, Layout L, int W)
{
    FrameWork F = frame_work(scratch, L, blockIdx.y);
    const Tables &T = *F.T;
    if (T.status != ST_OK || !separable(T, W)) return;
    const int lw = T.lw, r = blockIdx.x * (ROW_WARPS * 32) + threadIdx.x;
    if (r >= T.lh) return;
    const bool row0 = r == 0;
    int U = row0 ? 1 << (T.bits - 1) : 0;
    uint16_t *row = F.tiled + (size_t)r * lw;
    if ((lw & 63) == 0) {
        uint4 *p = reinterpret_cast<uint4 *>(row);
        const int nblk = lw >> 6;
        uint4 cur[8], nxt[8];
#pragma unroll
        for (int i = 0; i < 8; i++) cur[i] = p[i];
        for (int k = 0; k < nblk; k++) {
            if (k + 1 < nblk) {
#pragma unroll
                for (int i = 0; i < 8; i++) nxt[i] = p[(k + 1) * 8 + i];
            }
            row_block64(cur, U, row0);
#pragma unroll
            for (int i = 0; i < 8; i++) { p[k * 8 + i] = cur[i]; cur[i] = nxt[i]; }
        }
    } else {
        for (int c = 0; c < lw; c++) {
            U = row_step(row0, U, (int)(int16_t)row[c]);
            row[c] = (uint16_t)U;
        }
    }
}

// thread q owns stream columns 2q, 2q+1 and W/2 + 2q, W/2 + 2q + 1, i.e. output columns 4q .. 4q+3
template <int PASS>      // 0: column sums of a 32-row chunk   1: exclusive scan over chunks   2: samples out
__global__ void __launch_bounds__(128)
lj92_col_kernel(char *scratch, Layout L, uint16_t *__restrict__ out_base, size_t out_stride_px, int W, int H)
{
    FrameWork F = frame_work(scratch, L, blockIdx.z);
    const Tables &T = *F.T;
    if (T.status != ST_OK || !separable(T, W)) return;
    const int nq = W >> 2, q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint2 *sums = reinterpret_cast<uint2 *>(F.colsum);
    if (PASS == 1) {
        const int nch = (H + CH_ROWS - 1) / CH_ROWS;
        uint2 run = make_uint2(0u, 0u);
        for (int c0 = 0; c0 < nch; c0 += 16) {                                  // 16 independent loads in flight
            uint2 v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = c0 + i < nch ? sums[(size_t)(c0 + i) * nq + q] : make_uint2(0u, 0u);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (c0 + i < nch) sums[(size_t)(c0 + i) * nq + q] = run;
                run.x = __vadd2(run.x, v[i].x);
                run.y = __vadd2(run.y, v[i].y);
            }
        }
        return;
    }
    const int ch = blockIdx.y, y0 = ch * CH_ROWS, y1 = min(H, y0 + CH_ROWS);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(F.tiled) + q;          // word = 2 samples
    const int wrow = W >> 1, half = W >> 2;                                         // words per row / per half row
    uint2 acc = PASS == 2 ? sums[(size_t)ch * nq + q] : make_uint2(0u, 0u);
    uint16_t *out = out_base + (size_t)blockIdx.z * out_stride_px + 4 * q;
#pragma unroll 8
    for (int y = y0; y < y1; y++) {
        acc.x = __vadd2(acc.x, __ldg(src + (size_t)y * wrow));
        acc.y = __vadd2(acc.y, __ldg(src + (size_t)y * wrow + half));
        if (PASS == 2) {
            const int dy = 2 * y < H ? 2 * y : 2 * y - H + 1;
            __stcs(reinterpret_cast<uint2 *>(out + (size_t)dy * W),
                   make_uint2(__byte_perm(acc.x, acc.y, 0x5410), __byte_perm(acc.x, acc.y, 0x7632)));
        }
    }
    if (PASS == 0) sums[(size_t)ch * nq + q] = acc;
}

// ---- 6. untile (main.c:656-668) -----------------------------------------------------------------
// dst[(2y mod H) + (2y div H)][(2x mod W) + (2x div W)] = src[y][x] over the decoded samples taken
// as a W x H raster.  For even W, H that is a bijection: output row dy comes from src row dy/2 (even)
// or H/2 + dy/2 (odd), and likewise for columns.  8 output samples per thread.
__global__ void __launch_bounds__(256)
lj92_untile_kernel(char *scratch, Layout L, uint16_t *__restrict__ out_base, size_t out_stride_px, int W, int H,
                   int *__restrict__ status, int geo_ok)
{
    FrameWork F = frame_work(scratch, L, blockIdx.z);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) status[blockIdx.z] = F.T->status;
    if (F.T->status != ST_OK || (geo_ok && separable(*F.T, W))) return;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;                        // g: group of 8 output columns
    if (g * 8 >= W) return;
    for (int dy = blockIdx.y; dy < H; dy += gridDim.y) {
        const int y = (dy >> 1) + (dy & 1) * (H >> 1);
        const uint16_t *src = F.tiled + (size_t)y * W;
        const uint2 a = __ldcs(reinterpret_cast<const uint2 *>(src + g * 4));                 // even output columns
        const uint2 b = __ldcs(reinterpret_cast<const uint2 *>(src + (W >> 1) + g * 4));     // odd output columns
        uint4 o;
        o.x = __byte_perm(a.x, b.x, 0x5410); o.y = __byte_perm(a.x, b.x, 0x7632);
        o.z = __byte_perm(a.y, b.y, 0x5410); o.w = __byte_perm(a.y, b.y, 0x7632);
        __stcs(reinterpret_cast<uint4 *>(out_base + (size_t)blockIdx.z * out_stride_px + (size_t)dy * W + g * 8), o);
    }
}

// any geometry (odd sizes collide exactly where the reference's loop overwrites; last writer differs)
__global__ void __launch_bounds__(256)
lj92_untile_scalar_kernel(char *scratch, Layout L, uint16_t *__restrict__ out_base, size_t out_stride_px, int W, int H,
                          int *__restrict__ status, int geo_ok)
{
    FrameWork F = frame_work(scratch, L, blockIdx.z);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) status[blockIdx.z] = F.T->status;
    if (F.T->status != ST_OK || (geo_ok && separable(*F.T, W))) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int dx = (2 * x) % W + (2 * x) / W;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const int dy = (2 * y) % H + (2 * y) / H;
        out_base[(size_t)blockIdx.z * out_stride_px + (size_t)dy * W + dx] = F.tiled[(size_t)y * W + x];
    }
}

}  // namespace

size_t lj92_scratch_bytes(size_t payload_bytes, size_t npix, int nframes)
{
    return lj92::make_layout(payload_bytes, npix).frame_stride * (size_t)nframes;
}

// returns the number of kernels launched (> 0) or a negative MLVB_ERR_* code
int launch_lj92_decode(const void *d_payload, size_t payload_stride, size_t payload_bytes, uint16_t *d_out,
                       size_t out_stride_px, int w, int h, int nframes, int *d_status, void *d_scratch,
                       size_t scratch_bytes, cudaStream_t st)
{
    if (payload_bytes < 8 || payload_bytes > 0x1FFFFFF0u || w <= 0 || h <= 0) return MLVB_ERR_ARG;
    const size_t npix = (size_t)w * h;
    const Layout L = make_layout(payload_bytes, npix);
    if (!d_scratch || scratch_bytes < L.frame_stride * (size_t)nframes) return MLVB_ERR_ARG;
    char *sc = (char *)d_scratch;
    const uint8_t *pl = (const uint8_t *)d_payload;
    // prediction: strips of 32 rows, PREDICT_WARPS per block, taken by ticket (any grid size is correct;
    // this one gives every strip its own warp when the stream's height is the frame's)
    const int predict_parts = std::max(1, ceil_div(ceil_div(h, 32), PREDICT_WARPS));
    const unsigned ncta = L.max_cta;
    lj92_parse_kernel<<<nframes, 32, 0, st>>>(pl, payload_stride, payload_bytes, sc, L, w, h);
    lj92_unstuff_count_kernel<<<dim3(L.raw_chunks, nframes), UNSTUFF_THREADS, 0, st>>>(pl, payload_stride, payload_bytes, sc, L);
    lj92_unstuff_scan_kernel<<<nframes, 1024, 0, st>>>(sc, L);
    lj92_unstuff_scatter_kernel<<<dim3(L.raw_chunks, nframes), UNSTUFF_THREADS, 0, st>>>(pl, payload_stride, payload_bytes, sc, L);
    lj92_dec_kernel<0><<<dim3(ncta, nframes), DEC_THREADS, 0, st>>>(sc, L, (uint32_t)npix);
    lj92_dec_kernel<1><<<dim3(ncta, nframes), DEC_THREADS, 0, st>>>(sc, L, (uint32_t)npix);
    lj92_resolve_kernel<<<nframes, DEC_THREADS, 0, st>>>(sc, L, (uint32_t)npix);
    lj92_dec_kernel<2><<<dim3(ncta, nframes), DEC_THREADS, 0, st>>>(sc, L, (uint32_t)npix);
    // predictor-6 streams (what MLV holds): row recurrence + column prefix sums fused with the untile
    const int geo_ok = (w % 4) == 0 && (h % 2) == 0 && (out_stride_px % 4) == 0 && (((uintptr_t)d_out) & 7) == 0;
    int launches = 10;
    if (geo_ok) {
        const dim3 cg(ceil_div(w / 4, 128), ceil_div(h, CH_ROWS), nframes);
        lj92_row_kernel<<<dim3(ceil_div(h, ROW_WARPS * 32), nframes), ROW_WARPS * 32, 0, st>>>(sc, L, w);
        lj92_col_kernel<0><<<cg, 128, 0, st>>>(sc, L, d_out, out_stride_px, w, h);
        lj92_col_kernel<1><<<dim3(cg.x, 1, nframes), 128, 0, st>>>(sc, L, d_out, out_stride_px, w, h);
        lj92_col_kernel<2><<<cg, 128, 0, st>>>(sc, L, d_out, out_stride_px, w, h);
        launches += 4;
    }
    // every other stream: wavefront prediction, then untile (both skip the frames handled above)
    lj92_predict_kernel<<<dim3(predict_parts, nframes), PREDICT_WARPS * 32, 0, st>>>(sc, L, w, geo_ok);
    if ((w % 8) == 0 && (h % 2) == 0 && (out_stride_px % 8) == 0 && (((uintptr_t)d_out) & 15) == 0)
        lj92_untile_kernel<<<dim3(ceil_div(w / 8, 256), std::min(h, 120), nframes), 256, 0, st>>>(sc, L, d_out, out_stride_px, w, h, d_status, geo_ok);
    else
        lj92_untile_scalar_kernel<<<dim3(ceil_div(w, 256), std::min(h, 120), nframes), 256, 0, st>>>(sc, L, d_out, out_stride_px, w, h, d_status, geo_ok);
    MLVB_CUDA_OK(cudaGetLastError());
    return launches;
}
