// hdrpreview.cu -- the fast dual-ISO preview and deflicker.
//
// Replaces reference hdr.c:40-227 (hdr_convert_data), histogram.c:33-84 (16-bit-counter histograms and
// their median) and main.c:895-906 (deflicker).
//
// Preview: (1) four green histograms over every 5th row -> row phase (host, O(bins));
// (2) focus pixels with the horizontal interpolator; (3) histogram matching of the dark against the
// bright exposure + weighted least squares (host, same operation order in fp64); (4) one thread per
// (column, row parity) walks down its column in place: the reference rewrites the frame row by row,
// so a patched pixel reads the ALREADY rewritten pixel two rows up and the NOT YET rewritten pixel
// two rows down -- a per-column sequential dependency that this mapping reproduces exactly;
// (5) << 2 to 16 bit.
#include <math.h>

#include <vector>

#include "context.cuh"

namespace {

__global__ void preview_hist_kernel(const uint16_t *__restrict__ img, int w, int h, int white, unsigned *__restrict__ hist)
{
    // rows y = 4, 9, 14, ... < h-4; samples x = x0 + 4k, x0 = (y+1)%2  (hdr.c:58-61, histogram.c:52-59)
    const int y = 4 + 5 * blockIdx.y;
    if (y >= h - 4) return;
    const int x0 = (y + 1) % 2, size = w - x0;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * k >= size) return;
    const int v = img[(size_t)y * w + x0 + 4 * k];
    atomicAdd(&hist[(y % 4) * (white + 1) + min(v, white)], 1u);
}

struct PreviewParams { int w, h, black, white, start, shadow; double a, b; };

__device__ __forceinline__ double pv_scale(int v, const PreviewParams &P)
{
    const double s = (double)(v - P.black) * P.a + (double)P.black + P.b;
    return ((double)P.white < s) ? (double)P.white : s;
}

__global__ void deflicker_hist_kernel(const uint16_t *__restrict__ img, size_t nsamples, int white, unsigned *__restrict__ hist)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // sample k = pixel 1 + 2k (main.c:901)
    if (k >= nsamples) return;
    atomicAdd(&hist[min((int)img[1 + 2 * k], white)], 1u);
}

// histogram.c:64-76 with its uint16 counters: counts wrap modulo 65536
int median_u16_bins(const unsigned *bins, int white, unsigned count)
{
    const unsigned middle = count / 2;
    unsigned cur = 0;
    for (int i = 0; i <= white; i++) { cur += (bins[i] & 0xFFFF); if (cur > middle) return i; }
    return 0;
}

}  // namespace

size_t hdr_preview_scratch_bytes(int white) { return (size_t)4 * (white + 2) * sizeof(unsigned) + 256; }
size_t deflicker_scratch_bytes(int bpp) { return (size_t)((1 << bpp) + 4) * sizeof(unsigned) + 256; }

namespace {

// One thread per (column, row parity) walks down in place.  The rewritten, still unshifted value two
// rows up stays in a register (the store below already carries the final << 2 of hdr.c:216-221); the
// pixel two rows down is read before this thread rewrites it, exactly like the reference's row loop.
__global__ void preview_columns_kernel(uint16_t *__restrict__ img, const PreviewParams P)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * P.w) return;
    const int x = t % P.w, parity = t / P.w;
    const int w = P.w, h = P.h;
    int up = 0;                                   // rewritten, unshifted value two rows up
    for (int y = parity; y < h; y += 2) {
        const size_t i = (size_t)y * w + x;
        const int v = img[i];
        const int dn = (y + 2 < h) ? (int)img[i + 2 * w] : 0;
        int out = v;
        if (((y - P.start + 4) % 4) >= 2) {
            if (v >= P.white) out = (uint16_t)(y > 2 ? (y < h - 2 ? (up + dn) / 2 : up) : dn);
            else out = (uint16_t)pv_scale(v, P);
        } else if (v < P.shadow) {
            double r;
            if (y > 2) r = (y < h - 2) ? ((double)up + pv_scale(dn, P)) / 2 : (double)up;
            else r = pv_scale(dn, P);
            out = (uint16_t)r;
        }
        up = out;
        img[i] = (uint16_t)((out << 2) & 0xFFFF);
    }
}

}  // namespace

// hdr_convert_data on a device frame; returns 1 converted (caller multiplies black/white by 4), 0 not dual ISO
int run_hdr_preview(mlvb_context *ctx, const struct frame_headers *hdr, const FrameGeom &g, uint16_t *d_img, void *d_aux,
                    cudaStream_t st)
{
    const int w = g.w, h = g.h;
    const int black = (uint16_t)g.black, white = (uint16_t)g.white;
    if (w < 8 || h < 12) return 0;
    unsigned *d_hist = (unsigned *)d_aux;
    const size_t nb = (size_t)4 * (white + 1);
    MLVB_CUDA_OK(cudaMemsetAsync(d_hist, 0, nb * sizeof(unsigned), st));
    const int nrows = (h - 4 - 4 + 4) / 5;                                    // y = 4 + 5j < h - 4
    if (nrows > 0) preview_hist_kernel<<<dim3(ceil_div((w + 3) / 4, 128), nrows), 128, 0, st>>>(d_img, w, h, white, d_hist);
    ctx->launches += 1;
    std::vector<unsigned> hist(nb);
    MLVB_CUDA_OK(cudaMemcpyAsync(hist.data(), d_hist, nb * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MLVB_CUDA_OK(stream_wait(ctx, st));
    unsigned count[4] = {0, 0, 0, 0};
    for (int y = 4; y < h - 4; y += 5) count[y % 4] += (unsigned)(w - (y + 1) % 2) / 4;   // histogram.c:58
    int m[4];
    for (int i = 0; i < 4; i++) m[i] = median_u16_bins(hist.data() + (size_t)i * (white + 1), white, count[i]) - black;
    int start, lo, hi;
    if (m[2] > m[0] * 2 && m[2] > m[1] * 2 && m[3] > m[0] * 2 && m[3] > m[1] * 2) { start = 0; lo = 0; hi = 2; }
    else if (m[0] > m[1] * 2 && m[0] > m[2] * 2 && m[3] > m[1] * 2 && m[3] > m[2] * 2) { start = 1; lo = 1; hi = 0; }
    else if (m[0] > m[2] * 2 && m[0] > m[3] * 2 && m[1] > m[2] * 2 && m[1] > m[3] * 2) { start = 2; lo = 2; hi = 0; }
    else if (m[1] > m[0] * 2 && m[1] > m[3] * 2 && m[2] > m[0] * 2 && m[2] > m[3] * 2) { start = 3; lo = 0; hi = 2; }   // sic, hdr.c:93-94
    else { fprintf(stderr, "Could not detect dual ISO interlaced lines\n"); return 0; }

    // fix_focus_pixels(.., 1) (hdr.c:117)
    std::shared_ptr<PixelList> focus;
    {
        std::lock_guard<std::mutex> lk(ctx->clip_mu);
        int rc = get_focus_pixel_map(ctx, hdr, g, &focus);
        if (rc) return rc;
    }
    if (focus && focus->nlevels && g.black <= MLVB_MAX_BLACK) {
        FrameGeom gw = g;
        gw.w = w; gw.h = h;
        int rc = apply_pixel_list(ctx, *focus, d_img, gw, g.npix, 1, 1, 1, st);
        if (rc) return rc;
    }

    // histogram matching + weighted least squares (hdr.c:119-182), same fp64 operation order
    const unsigned *hl = hist.data() + (size_t)lo * (white + 1), *hh = hist.data() + (size_t)hi * (white + 1);
    const int min_pix = 100, total = (int)count[0];
    std::vector<int> dx, dy;
    std::vector<double> dw;
    int acc_lo = 0, acc_hi = 0, raw_lo = 0, prev = 0;
    for (int raw_hi = 0; raw_hi < total && raw_hi <= white; raw_hi++) {
        acc_hi += (int)(hh[raw_hi] & 0xFFFF);
        while (acc_lo < acc_hi && raw_lo <= white) { acc_lo += (int)(hl[raw_lo] & 0xFFFF); raw_lo++; }
        if (raw_lo >= white) break;
        if (acc_hi - prev > min_pix && acc_hi > total * 1 / 100 && acc_hi < total * 99.99 / 100) {
            dx.push_back(raw_hi - black); dy.push_back(raw_lo - black);
            dw.push_back((raw_hi - black + 100) > 0 ? (raw_hi - black + 100) : 0);
            prev = acc_hi;
        }
    }
    double mx = 0, my = 0, mxy = 0, mx2 = 0, wsum = 0;
    for (size_t i = 0; i < dx.size(); i++) {
        mx += dx[i] * dw[i]; my += dy[i] * dw[i];
        mxy += (double)dx[i] * dy[i] * dw[i]; mx2 += (double)dx[i] * dx[i] * dw[i];
        wsum += dw[i];
    }
    mx /= wsum; my /= wsum; mxy /= wsum; mx2 /= wsum;
    PreviewParams P;
    P.w = w; P.h = h; P.black = black; P.white = white; P.start = start;
    P.a = (mxy - mx * my) / (mx2 - mx * mx);
    P.b = my - P.a * mx;
    P.shadow = (uint16_t)(black + 1 / (P.a * P.a) + P.b);
    preview_columns_kernel<<<ceil_div(2 * w, 128), 128, 0, st>>>(d_img, P);
    ctx->launches += 1;
    MLVB_CUDA_OK(cudaGetLastError());
    return 1;
}

// deflicker (main.c:895-906): median of every other sample -> exposure_bias
int run_deflicker(mlvb_context *ctx, const FrameGeom &g, const uint16_t *d_img, int target, void *d_aux, cudaStream_t st,
                  int32_t bias[2])
{
    const int white = (uint16_t)((1 << g.bpp) + 1);
    const size_t nsamples = ((g.npix * 2 - 1) / 2 + 1) / 2;                 // ceil(((size-1)/2) / 2) loop trips, i = 0,2,4,..
    const unsigned count = (unsigned)(((g.npix * 2 - 1) / 2) / 2);          // histogram.c:58: size / (skip+1)
    unsigned *d_hist = (unsigned *)d_aux;
    MLVB_CUDA_OK(cudaMemsetAsync(d_hist, 0, (size_t)(white + 1) * sizeof(unsigned), st));
    if (nsamples) deflicker_hist_kernel<<<ceil_div(nsamples, 256), 256, 0, st>>>(d_img, nsamples, white, d_hist);
    ctx->launches += 1;
    std::vector<unsigned> hist(white + 1);
    MLVB_CUDA_OK(cudaMemcpyAsync(hist.data(), d_hist, hist.size() * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MLVB_CUDA_OK(stream_wait(ctx, st));
    const int median = median_u16_bins(hist.data(), white, count);
    const int black = (uint16_t)g.black;
    const double correction = log2((double)(target - black) / (median - black));
    bias[0] = (int32_t)(correction * 10000);
    bias[1] = 10000;
    return MLVB_OK;
}
