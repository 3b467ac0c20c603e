// scan.cuh -- ordered (raster-order preserving) prefix sums used by bad-pixel compaction and by the
// stripes statistics (position of each accepted sample in the reference's sequential rand() stream).
#pragma once
#include "common.cuh"

// Exclusive prefix sum of `v` over the CTA in thread order; `total` receives the CTA sum.
// blockDim.x must be a multiple of 32 and <= 1024.
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned &total)
{
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_total;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                      // protect s_warp / s_total reuse across calls
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < nw ? s_warp[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;          // exclusive warp offsets
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    total = s_total;
    return s_warp[wid] + inc - v;
}

// Single-CTA exclusive scan of counts[0..n) in place (64-bit running carry); total written to *total.
static __global__ void scan_counts_kernel(unsigned long long *counts, unsigned n, unsigned long long *total)
{
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned base = 0; base < n; base += blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < n ? (unsigned)counts[i] : 0u;
        unsigned tot;
        const unsigned ex = block_exclusive_scan(v, tot);
        const unsigned long long c = carry;
        if (i < n) counts[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
