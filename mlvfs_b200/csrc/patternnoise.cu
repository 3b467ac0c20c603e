// patternnoise.cu -- row/column pattern-noise removal on the Bayer frame (as int16).
//
// Replaces reference patternnoise.c:357-380 (fix_pattern_noise) and everything under it:
// fix_column_noise_rggb :312-355, horizontal_edge_aware_blur_rggb :88-180, fix_column_noise :185-282.
//
// Per direction (columns first, then the same on the transposed frame):
//   pn_blur_kernel     one CTA per half-res row.  For every RGGB quad: the run of neighbours whose
//                      average green stays within 500 of its own (<= 25 left, <= 24 right), the exact
//                      LOWER medians (wirth.h:129) of G1, G2, R-G, B-G over that run by bitwise
//                      selection in registers, noise = original - denoised, edge/highlight mask.
//                      Writes noise column-major so the next kernel reads columns contiguously.
//   pn_column_kernel   one warp per (plane, column): lower median of the unmasked noise -> offset.
//   pn_offsets_kernel  one CTA per plane: lower median of the column offsets (colour-cast guard).
//   pn_apply_kernel    per pixel: clamp(clamp(v + offset[col], +-32767) - median_offset, 0, 32760).
// All integer, bit-exact.  ALU-bound (~6.5k compare-adds per quad), not HBM-bound.
#include "kernels.cuh"

namespace {

constexpr int PN_REACH = 25;        // strength 50 / 2 (patternnoise.c:106)
constexpr int PN_THR = 500;
constexpr int PN_MAXWIN = 50;
constexpr int PN_MASKED = INT_MIN;

// exact k-th smallest (k = lower-median index) of a[lo .. lo+n) by binary search on the value bits
__device__ __noinline__ int lower_median_window(const int16_t *a, int lo, int n)
{
    int v[PN_MAXWIN];
#pragma unroll
    for (int j = 0; j < PN_MAXWIN; j++) v[j] = (j < n) ? (int)a[lo + j] + 32768 : 0x7FFFFFFF;
    const int k = (n & 1) ? n / 2 : n / 2 - 1;
    int res = 0;
    for (int bit = 15; bit >= 0; bit--) {
        const int cand = res | (1 << bit);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < PN_MAXWIN; j++) cnt += (v[j] < cand) ? 1 : 0;
        if (cnt <= k) res = cand;
    }
    return res - 32768;
}

// raw: Bayer int16 [h][w]; planes are addressed as plane c = (dx = c & 1, dy = c >> 1), half-res pw x ph
__device__ __forceinline__ int16_t plane_at(const int16_t *raw, int w, int c, long long flat, int pw)
{
    const int y = (int)(flat / pw), x = (int)(flat - (long long)y * pw);
    return raw[(size_t)(2 * y + (c >> 1)) * w + 2 * x + (c & 1)];
}

__global__ void __launch_bounds__(256)
pn_blur_kernel(const int16_t *__restrict__ raw, int w, int h, int white, int *__restrict__ noiseT)
{
    extern __shared__ int16_t sm[];
    const int pw = w / 2, ph = h / 2, y = blockIdx.x;
    int16_t *pl[4] = {sm, sm + pw, sm + 2 * pw, sm + 3 * pw};      // r, g1, g2, b of this half-res row
    int16_t *avg = sm + 4 * pw, *drg = sm + 5 * pw, *dbg = sm + 6 * pw;
    for (int x = threadIdx.x; x < pw; x += blockDim.x) {
        const int16_t r = raw[(size_t)(2 * y) * w + 2 * x], g1 = raw[(size_t)(2 * y) * w + 2 * x + 1];
        const int16_t g2 = raw[(size_t)(2 * y + 1) * w + 2 * x], b = raw[(size_t)(2 * y + 1) * w + 2 * x + 1];
        const int16_t a = (int16_t)(((int)g1 + (int)g2) / 2);       // patternnoise.c:59-65
        pl[0][x] = r; pl[1][x] = g1; pl[2][x] = g2; pl[3][x] = b;
        avg[x] = a;
        drg[x] = (int16_t)(r - a);
        dbg[x] = (int16_t)(b - a);
    }
    __syncthreads();
    const long long nplane = (long long)pw * ph;
    for (int x = threadIdx.x; x < pw; x += blockDim.x) {
        const int p0 = avg[x];
        int xr = x + 1, xl = x - 1;
        const int rmax = min(x + PN_REACH, pw), lmin = max(x - PN_REACH, 0);
        while (xr < rmax && abs((int)avg[xr] - p0) <= PN_THR) xr++;
        while (xl >= lmin && abs((int)avg[xl] - p0) <= PN_THR) xl--;
        const int lo = xl + 1, n = xr - xl - 1;
        const int m1 = lower_median_window(pl[1], lo, n), m2 = lower_median_window(pl[2], lo, n);
        const int mg = (m1 + m2) / 2;
        int16_t den[4];
        den[0] = (int16_t)(lower_median_window(drg, lo, n) + mg);
        den[1] = (int16_t)m1;
        den[2] = (int16_t)m2;
        den[3] = (int16_t)(lower_median_window(dbg, lo, n) + mg);
        const long long i = (long long)y * pw + x;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int16_t o = pl[c][x];
            int hg = 0;                                             // patternnoise.c:76-85, flat index
            if (i >= 2 && i + 2 < nplane) {
                const int16_t a = (x >= 2) ? pl[c][x - 2] : plane_at(raw, w, c, i - 2, pw);
                const int16_t b = (x + 2 < pw) ? pl[c][x + 2] : plane_at(raw, w, c, i + 2, pw);
                hg = (int16_t)(a - b);
            }
            const bool masked = (abs(hg) > 500) || ((int)o >= white);
            noiseT[(size_t)c * nplane + (size_t)x * ph + y] = masked ? PN_MASKED : (int)(int16_t)(o - den[c]);
        }
    }
}

// lower median over the non-masked entries of v[0..n) (one warp); returns count through *count
__device__ int warp_lower_median(const int *v, int n, int lane, int *count)
{
    int cnt = 0;
    for (int i = lane; i < n; i += 32) cnt += (v[i] != PN_MASKED);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    *count = cnt;
    if (cnt == 0) return 0;
    const int k = (cnt & 1) ? cnt / 2 : cnt / 2 - 1;
    int res = 0;                                                    // values biased by 2^17 into [0, 2^18)
    for (int bit = 17; bit >= 0; bit--) {
        const int cand = res | (1 << bit);
        int c = 0;
        for (int i = lane; i < n; i += 32) {
            const int x = v[i];
            c += (x != PN_MASKED && x + (1 << 17) < cand);
        }
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if (c <= k) res = cand;
    }
    return res - (1 << 17);
}

__global__ void pn_column_kernel(const int *__restrict__ noiseT, int pw, int ph, int *__restrict__ offsets)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= 4 * pw) return;
    int cnt;
    const int med = warp_lower_median(noiseT + (size_t)warp * ph, ph, lane, &cnt);
    if (lane == 0) offsets[warp] = (cnt < 10) ? 0 : -med;           // patternnoise.c:252-254
}

__global__ void pn_offsets_kernel(const int *__restrict__ offsets, int pw, int *__restrict__ mc)
{
    // one warp per plane is plenty (pw <= a few thousand)
    const int c = blockIdx.x, lane = threadIdx.x;
    int cnt;
    const int med = warp_lower_median(offsets + (size_t)c * pw, pw, lane, &cnt);
    if (lane == 0) mc[c] = med;
}

__global__ void pn_apply_kernel(int16_t *__restrict__ raw, int w, int h, const int *__restrict__ offsets,
                                const int *__restrict__ mc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int pw = w / 2;
    if (x >= 2 * pw || y >= 2 * (h / 2)) return;
    const int c = (x & 1) | ((y & 1) << 1);
    int v = raw[(size_t)y * w + x];
    v = min(max(v + offsets[c * pw + (x >> 1)], -32767), 32767);   // patternnoise.c:258-264
    v = min(max((int)(int16_t)v - mc[c], 0), 32760);                 // patternnoise.c:266-274
    raw[(size_t)y * w + x] = (int16_t)v;
}

__global__ void pn_transpose_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, int w, int h)
{
    __shared__ int16_t tile[32][33];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int x = x0 + threadIdx.x, y = y0 + j;
        if (x < w && y < h) tile[j][threadIdx.x] = in[(size_t)y * w + x];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int y = y0 + threadIdx.x, x = x0 + j;               // out[x][y]
        if (x < w && y < h) out[(size_t)x * h + y] = tile[threadIdx.x][j];
    }
}

int one_direction(int16_t *raw, int w, int h, int white, int *d_noiseT, int *d_offsets, int *d_mc, cudaStream_t st)
{
    const int pw = w / 2, ph = h / 2;
    if (pw < 1 || ph < 1) return MLVB_OK;
    const size_t smem = (size_t)7 * pw * sizeof(int16_t);
    if (smem > 200 * 1024) return MLVB_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        MLVB_CUDA_OK(cudaFuncSetAttribute(pn_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pn_blur_kernel<<<ph, 256, smem, st>>>(raw, w, h, white, d_noiseT);
    pn_column_kernel<<<ceil_div((long long)4 * pw * 32, 256), 256, 0, st>>>(d_noiseT, pw, ph, d_offsets);
    pn_offsets_kernel<<<4, 32, 0, st>>>(d_offsets, pw, d_mc);
    pn_apply_kernel<<<dim3(ceil_div(w, 256), h), 256, 0, st>>>(raw, w, h, d_offsets, d_mc);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

}  // namespace

size_t pattern_noise_scratch_bytes(int w, int h)
{
    const size_t npix = (size_t)w * h;
    return npix * sizeof(int16_t) + npix * sizeof(int) + (size_t)(2 * (w > h ? w : h) + 64) * sizeof(int) + 256;
}

// in place on d_raw (one frame); d_scratch >= pattern_noise_scratch_bytes(w, h).  10 launches.
int launch_pattern_noise(int16_t *d_raw, int w, int h, int white, void *d_scratch, cudaStream_t st)
{
    const size_t npix = (size_t)w * h;
    int16_t *d_t = (int16_t *)d_scratch;
    int *d_noiseT = (int *)((uint8_t *)d_scratch + ((npix * sizeof(int16_t) + 255) & ~(size_t)255));
    int *d_offsets = d_noiseT + npix;
    int *d_mc = d_offsets + 2 * (w > h ? w : h);
    int rc = one_direction(d_raw, w, h, white, d_noiseT, d_offsets, d_mc, st);        // column noise
    if (rc) return rc;
    pn_transpose_kernel<<<dim3(ceil_div(w, 32), ceil_div(h, 32)), dim3(32, 8), 0, st>>>(d_raw, d_t, w, h);
    rc = one_direction(d_t, h, w, white, d_noiseT, d_offsets, d_mc, st);              // row noise (patternnoise.c:371-379)
    if (rc) return rc;
    pn_transpose_kernel<<<dim3(ceil_div(h, 32), ceil_div(w, 32)), dim3(32, 8), 0, st>>>(d_t, d_raw, h, w);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
