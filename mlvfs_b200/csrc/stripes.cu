// stripes.cu -- vertical-stripe statistics (once per clip, from its first processed frame).
//
// Replaces reference stripes.c:108-140 (add_pixel) and stripes.c:143-248 (stripes_compute_correction).
// The per-frame application of the coefficients lives in chroma.cu (fused store epilogue).
//
// The reference walks every 8-pixel block in raster order and feeds up to 24 column pairs into
// eight 65536-bin histograms of log2((a+U1)/(b+U2)), where U1,U2 come from consecutive rand()
// calls (stripes.c:129-130).  To reproduce the same dither stream in parallel we
//   1. count, per block, how many of its 24 samples pass the brightness tests,
//   2. prefix-sum those counts in raster order -> index of each sample in the rand() stream,
//   3. histogram with the dither values looked up at that index in a host-generated table of
//      rand() % 1024 (the host owns the process-wide glibc-compatible generator state).
#include "kernels.cuh"
#include "scan.cuh"

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_BINS = 65536;

// sample order of stripes.c:180-208: for column group g = 2..7, four samples against the reference
// column of this block (p[g&1]) or of the next block (p[8 + (g&1)])
__device__ __forceinline__ int ref_index(int g, int r)
{
    const int near_reps = (g < 4) ? 3 : (g < 6 ? 2 : 1);
    return (r < near_reps) ? (g & 1) : 8 + (g & 1);
}

__device__ __forceinline__ bool pair_ok(int a, int b, double white_limit)
{
    return min(a, b) >= 32 && !((double)max(a, b) > white_limit);       // stripes.c:113-117
}

__device__ __forceinline__ void load_block(const uint16_t *im, size_t i, int black, int p[10])
{
#pragma unroll
    for (int k = 0; k < 10; k++) p[k] = (int)im[i + k] - black;
}

__global__ void __launch_bounds__(ST_THREADS)
stripes_count_kernel(const uint16_t *__restrict__ im, int w, int h, int nblk, int black, double white_limit,
                     unsigned long long *__restrict__ cta_counts)
{
    const size_t t = (size_t)blockIdx.x * ST_THREADS + threadIdx.x;
    unsigned c = 0;
    if (t < (size_t)nblk * h) {
        const int y = (int)(t / nblk), bx = (int)(t - (size_t)y * nblk);
        int p[10];
        load_block(im, (size_t)y * w + 8 * bx, black, p);
#pragma unroll
        for (int g = 2; g < 8; g++)
#pragma unroll
            for (int r = 0; r < 4; r++) c += pair_ok(p[ref_index(g, r)], p[g], white_limit) ? 1u : 0u;
    }
    unsigned total;
    block_exclusive_scan(c, total);
    if (threadIdx.x == 0) cta_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(ST_THREADS)
stripes_hist_kernel(const uint16_t *__restrict__ im, int w, int h, int nblk, int black, double white_limit,
                    const unsigned long long *__restrict__ cta_offsets, const uint16_t *__restrict__ dither,
                    unsigned *__restrict__ hist, unsigned *__restrict__ num)
{
    const size_t t = (size_t)blockIdx.x * ST_THREADS + threadIdx.x;
    const bool active = t < (size_t)nblk * h;
    int p[10];
    unsigned c = 0;
    if (active) {
        const int y = (int)(t / nblk), bx = (int)(t - (size_t)y * nblk);
        load_block(im, (size_t)y * w + 8 * bx, black, p);
#pragma unroll
        for (int g = 2; g < 8; g++)
#pragma unroll
            for (int r = 0; r < 4; r++) c += pair_ok(p[ref_index(g, r)], p[g], white_limit) ? 1u : 0u;
    }
    unsigned total;
    const unsigned ex = block_exclusive_scan(c, total);
    if (!active || c == 0) return;
    size_t s = 2 * (cta_offsets[blockIdx.x] + ex);      // two rand() calls per accepted sample
#pragma unroll
    for (int g = 2; g < 8; g++) {
        unsigned n_g = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int a = p[ref_index(g, r)], b = p[g];
            if (!pair_ok(a, b, white_limit)) continue;
            const double af = (double)a + (double)dither[s] / 1024.0 - 0.5;
            const double bf = (double)b + (double)dither[s + 1] / 1024.0 - 0.5;
            s += 2;
            const double ev = log2(af / bf);
            int bin = (int)(ST_BINS / 2 + ev * ST_BINS / 2);             // F2H, stripes.c:105
            bin = min(max(bin, 0), ST_BINS - 1);
            atomicAdd(&hist[g * ST_BINS + bin], 1u);
            n_g++;
        }
        if (n_g) atomicAdd(&num[g], n_g);
    }
}

// stripes.c:219-233: first bin where the running count reaches num/2 (one CTA per column group)
__global__ void stripes_median_kernel(const unsigned *__restrict__ hist, const unsigned *__restrict__ num,
                                      int *__restrict__ median_bin)
{
    const int g = blockIdx.x;
    const unsigned half = num[g] / 2;
    __shared__ unsigned carry;
    __shared__ int found;
    if (threadIdx.x == 0) { carry = 0; found = ST_BINS; }
    __syncthreads();
    for (int base = 0; base < ST_BINS; base += blockDim.x) {
        const unsigned v = hist[g * ST_BINS + base + threadIdx.x];
        unsigned tot;
        const unsigned ex = block_exclusive_scan(v, tot);
        const unsigned c = carry;
        if (c + ex + v >= half) atomicMin(&found, base + (int)threadIdx.x);
        __syncthreads();
        if (found < ST_BINS) break;
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) median_bin[g] = found;
}

}  // namespace

// number of 8-pixel blocks per row visited by stripes.c:158 (x < row_start + xRes - 10, step 8)
int stripes_blocks_per_row(int w) { return w > 10 ? (w - 10 + 7) / 8 : 0; }

int stripes_count_ctas(int w, int h) { return ceil_div((long long)stripes_blocks_per_row(w) * h, ST_THREADS); }

// phase 1: accepted-sample counts per CTA, scanned in raster order; total at d_counts[nctas]
int launch_stripes_count(const uint16_t *d_img, int w, int h, int black, int white, unsigned long long *d_counts,
                         cudaStream_t st)
{
    const int nblk = stripes_blocks_per_row(w), nctas = stripes_count_ctas(w, h);
    if (nctas == 0) return MLVB_OK;
    stripes_count_kernel<<<nctas, ST_THREADS, 0, st>>>(d_img, w, h, nblk, black, white / 1.5, d_counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(d_counts, (unsigned)nctas, d_counts + nctas);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

// phase 2: histograms (d_hist: 8*65536 uint32, d_num: 8 uint32, both zeroed here) and the eight median bins
int launch_stripes_hist(const uint16_t *d_img, int w, int h, int black, int white, const unsigned long long *d_offsets,
                        const uint16_t *d_dither, unsigned *d_hist, unsigned *d_num, int *d_median_bin, cudaStream_t st)
{
    const int nblk = stripes_blocks_per_row(w), nctas = stripes_count_ctas(w, h);
    MLVB_CUDA_OK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * 8 * ST_BINS, st));
    MLVB_CUDA_OK(cudaMemsetAsync(d_num, 0, sizeof(unsigned) * 8, st));
    if (nctas)
        stripes_hist_kernel<<<nctas, ST_THREADS, 0, st>>>(d_img, w, h, nblk, black, white / 1.5, d_offsets, d_dither,
                                                          d_hist, d_num);
    stripes_median_kernel<<<8, 1024, 0, st>>>(d_hist, d_num, d_median_bin);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}
