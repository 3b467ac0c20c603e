// amaze.cu -- AMaZE + edge-directed interpolation of the dual-ISO half-resolution exposures.
//
// Replaces reference hdr.c:954-1229 (amaze_interpolate: squeeze the two exposures into half-height images,
// AMaZE demosaic, grayscale, edge-direction search, edge-directed interpolation in EV space) and
// amaze_demosaic_RT.c:113-1487 (tile program in amaze_tile.cuh).
//
// Kernels:
//   amz_squeeze_kernel     hdr.c:977-1026   20-bit mosaic -> float, dark rows on top / bright rows below, greens halved
//   amz_tiles_kernel       amaze_demosaic_RT.c:292-1470, persistent blocks pulling 160x160 tiles from a counter,
//                          work planes in a per-block global workspace (L2 resident), see amaze_tile.cuh
//   amz_gray_kernel        hdr.c:1045-1062  undo the green halving, clamp, gray = g/2 + r/4 + b/4, straight to raw2ev[gray]
//   amz_edge_dir_kernel    hdr.c:1094-1175  11 directions x 11 offsets of |EV| differences where high accuracy is needed
// The interpolation itself (hdr.c:1182-1210, edge_interp :940-952) is done inside dualiso.cu's per-pixel kernel
// through amz_edge_interp().
#include <stdlib.h>

#include <algorithm>

#include "amaze.cuh"

#include "amaze_tile.cuh"
#include "context.cuh"

int amz_threads()
{
    static const int v = [] { const char *e = getenv("MLVB_AMZ_THREADS"); int t = e ? atoi(e) : 256; return t >= 64 && t <= AMZ_THREADS_MAX && t % 32 == 0 ? t : 256; }();
    return v;
}

int amz_blocks()
{
    static const int v = [] {
        const char *e = getenv("MLVB_AMZ_BLOCKS_PER_SM");
        // 256 threads x 4 blocks per SM: measured best on B200 (tools/sweep_amz.sh); the row-sequential passes keep only
        // ~80-160 threads of a tile busy, so more, smaller tile programs per SM hide their latency better
        int b = e ? atoi(e) : 4, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        b = b >= 1 && b <= 6 ? b : 4;
        return std::min(b * sms, (int)AMZ_MAX_BLOCKS);
    }();
    return v;
}

#ifdef AMZ_PROFILE
__device__ unsigned long long g_amz_prof[32];      // cycles per section of the tile program, summed over tiles (thread 0 of each block)
#endif

namespace {

struct CudaCtx {
    int tid, nthr;
#ifdef AMZ_PROFILE
    long long last;
    __device__ __forceinline__ void mark(int k) { if (tid == 0) { const long long now = clock64(); atomicAdd(&g_amz_prof[k], (unsigned long long)(now - last)); last = now; } }
#else
    __device__ __forceinline__ void mark(int) {}
#endif
#if defined(AMZ_PF_L1)
    __device__ __forceinline__ void prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
    __device__ __forceinline__ void prefetch(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#if defined(AMZ_NO_STREAM)
    __device__ __forceinline__ float ld_stream(const float *p) { return *p; }
    __device__ __forceinline__ void st_stream(float *p, float v) { *p = v; }
#else
    __device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
    __device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
#endif
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void syncwarp() { __syncwarp(); }
    __device__ __forceinline__ void atomic_add(int *p, int v) { atomicAdd(p, v); }
};

// squeezed[y]: the half-height row later stages index with (0 for dropped rows, like the reference's zero-initialised
// array); sq_dst[y]: the row actually written (-1: none).  The reference walks the rows of one exposure in order and
// gives them consecutive rows yh starting at the exposure's first row (dark) or h / 4 * 2 + its first row (bright),
// stopping at yh >= h (hdr.c:977-1026): with the period-4 row pattern that is a closed form in y.
__global__ void amz_row_maps_kernel(int *__restrict__ squeezed, int *__restrict__ sq_dst, int h, int b0, int b1, int b2, int b3)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const int isb[4] = {b0, b1, b2, b3};
    const int p = isb[y & 3];
    int per = 0, before = 0, first = -1;
    for (int k = 0; k < 4; k++) {
        if (isb[k] != p) continue;
        if (first < 0) first = k;
        per++;
        if (k < (y & 3)) before++;
    }
    const int yh = (p ? h / 4 * 2 + first : first) + (y >> 2) * per + before;
    const bool ok = yh < h;
    squeezed[y] = ok ? yh : 0;
    sq_dst[y] = ok ? yh : -1;
}

__global__ void amz_squeeze_kernel(const uint32_t *__restrict__ raw32, float *__restrict__ rawf, const int *__restrict__ sq_dst,
                                   int w, int h, int ws, int black)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int yh = sq_dst[y];
    if (yh < 0) return;                                                  // bright rows dropped by the `yh >= h` break (hdr.c:1025)
    int p = (int)raw32[x + (size_t)y * w];
    if ((x & 1) != (y & 1)) p = (p - black) / 2 + black;                 // greens halved around black (hdr.c:991-992)
    rawf[(size_t)yh * ws + x] = (float)p;
}

__global__ void __launch_bounds__(AMZ_THREADS_MAX, AMZ_MIN_BLOCKS)
amz_tiles_kernel(const float *__restrict__ raw, float *__restrict__ red, float *__restrict__ green, float *__restrict__ blue,
                 int stride, int width, int height, int ntx, int nty, char *__restrict__ ws_base, unsigned *__restrict__ counter)
{
    __shared__ amaze::Shared S;
    __shared__ unsigned s_tile;
    const amaze::Ws W = amaze::carve(ws_base + (size_t)blockIdx.x * amaze::WS_BYTES);
    CudaCtx C{(int)threadIdx.x, (int)blockDim.x};
#ifdef AMZ_PROFILE
    C.last = clock64();
#endif
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned t = s_tile;
        if (t >= (unsigned)(ntx * nty)) break;
        const int ty = t / ntx, tx = t - ty * ntx;
        const amaze::Geom G = amaze::tile_geom(width, height, -16 + ty * (amaze::TS - 32), -16 + tx * (amaze::TS - 32));
#ifdef AMZ_PROFILE
        C.mark(31);                                                      // tile fetch
#endif
        amaze::tile_body(C, W, G, S, raw, red, green, blue, stride);
        __syncthreads();
    }
}

__global__ void amz_gray_kernel(const float *__restrict__ red, const float *__restrict__ green, const float *__restrict__ blue,
                                const int *__restrict__ squeezed, const int *__restrict__ raw2ev, int *__restrict__ grayev,
                                int w, int h, int ws, int black)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t i = (size_t)squeezed[y] * ws + x;
    const float g = amz_post_green(green[i], black), r = amz_post_rb(red[i]), b = amz_post_rb(blue[i]);
    const uint32_t gray = (uint32_t)(g / 2.0f + r / 4.0f + b / 4.0f);   // hdr.c:1062
    grayev[x + (size_t)y * w] = __ldg(raw2ev + (gray & 0xFFFFF));
}

__global__ void __launch_bounds__(128)
amz_edge_dir_kernel(const uint32_t *__restrict__ raw32, const int *__restrict__ grayev, uint8_t *__restrict__ edir,
                    int w, int h, int black, int white_darkened, int b0, int b1, int b2, int b3,
                    const double *__restrict__ fullres_curve, const int *__restrict__ fullres_lim)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int isb[4] = {b0, b1, b2, b3};
    uint8_t best = 5;
    if (x >= 5 && x < w - 5 && y >= 5 && y < h - 5) {
        const uint32_t p = raw32[x + (size_t)y * w];
        bool search;
        if (!isb[y & 3]) {                                                         // deep shadows of the dark exposure (hdr.c:1106-1120)
            // fullres_curve[p] > 0.8 from the per-black table of dualiso.cu (hdr.c:904-909);
            // where the table's crossing is a single index the test is a comparison
            const int lo = __ldg(fullres_lim + 2), hi = __ldg(fullres_lim + 3);
            search = lo == hi ? !((int)(p & 0xFFFFF) >= lo) : !(__ldg(fullres_curve + (p & 0xFFFFF)) > 0.8);
        }
        else search = !(p < (uint32_t)white_darkened);                             // bright exposure clipped (hdr.c:1122-1133)
        if (search) {
            const int s = (isb[y & 3] == isb[(y + 1) & 3]) ? -1 : 1;
            // Every direction reads the same four rows (y + 2s, y + s, y - 2s, y - 3s) at 11 consecutive columns around
            // its own column offsets: 11 x 11 x 4 = 484 loads of only 19 + 15 + 19 + 23 = 76 distinct cells.  The cells go
            // to registers once and the 11 direction costs come from them with compile-time indices.
            // flat indexing like the reference: x + dx + j may leave the row
            const int *ra = grayev + (long long)(y + 2 * s) * w + x, *rb = grayev + (long long)(y + s) * w + x;
            const int *rc = grayev + (long long)(y - 2 * s) * w + x, *rd = grayev + (long long)(y - 3 * s) * w + x;
            int A[19], B[15], Cc[19], Dd[23];
#pragma unroll
            for (int k = 0; k < 19; k++) { A[k] = __ldg(ra + k - 9); Cc[k] = __ldg(rc + k - 9); }
#pragma unroll
            for (int k = 0; k < 15; k++) B[k] = __ldg(rb + k - 7);
#pragma unroll
            for (int k = 0; k < 23; k++) Dd[k] = __ldg(rd + k - 11);
            constexpr int DX[11][4] = {{-4, -2, 4, 6}, {-3, -1, 3, 4}, {-2, -1, 2, 3}, {-1, -1, 1, 2}, {-1, 0, 1, 1}, {0, 0, 0, 0},
                                       {1, 0, -1, -1}, {1, 1, -1, -2}, {2, 1, -2, -3}, {3, 1, -3, -4}, {4, 2, -4, -6}};   // c_edge[d][0, 2, 4, 6]
            int e_best = 0x7FFFFFFF;
#pragma unroll
            for (int d = 0; d <= 10; d++) {
                int e = 0;
#pragma unroll
                for (int j = -5; j <= 5; j++) {
                    const int p1 = A[DX[d][0] + j + 9], p2 = B[DX[d][1] + j + 7], p3 = Cc[DX[d][2] + j + 9], p4 = Dd[DX[d][3] + j + 11];
                    e += abs(p1 - p2) + abs(p2 - p3) + abs(p3 - p4);
                }
                e += abs(d - 5) * (MLVB_EV_RES / 8);
                if (e < e_best) { e_best = e; best = (uint8_t)d; }
            }
        }
    }
    edir[x + (size_t)y * w] = best;
}

}  // namespace

size_t amaze_scratch_bytes(int w, int h, AmazeScratch *S, uint8_t *base)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { void *p = base ? base + o : nullptr; o += (bytes + 255) & ~(size_t)255; return p; };
    const size_t np = (size_t)w * h, nps = (size_t)(w + 16) * h;
    AmazeScratch s;
    s.rawf = (float *)take(nps * 4); s.red = (float *)take(nps * 4); s.green = (float *)take(nps * 4); s.blue = (float *)take(nps * 4);
    s.grayev = (int *)take(np * 4);
    s.edir = (uint8_t *)take(np);
    s.squeezed = (int *)take((size_t)h * 4); s.sq_dst = (int *)take((size_t)h * 4);
    s.counter = (unsigned *)take(256);
    const int ntiles = amaze::tiles_along(w) * amaze::tiles_along(h);
    s.nblocks = ntiles < amz_blocks() ? ntiles : amz_blocks();
    s.ws = (char *)take((size_t)s.nblocks * amaze::WS_BYTES);
    if (S) *S = s;
    return o;
}

// Everything of amaze_interpolate up to (and including) the direction map; the caller's per-pixel kernel
// then interpolates with amz_edge_interp().
int launch_amaze_stage(const uint32_t *d_raw32, int w, int h, int black, int white_darkened, const int is_bright[4],
                       const int *d_raw2ev, const double *d_fullres_curve, const int *d_fullres_lim, const AmazeScratch &A,
                       cudaStream_t st, int *launches)
{
    if (w & 3) {
        fprintf(stderr, "libmlvfs_b200: --amaze-edge needs a frame width that is a multiple of 4 (got %d)\n", w);
        return MLVB_ERR_UNSUPPORTED;
    }
    const int ws = w + 16;
    // row maps of the squeeze (hdr.c:977-1026), computed on the device (no host copy, no synchronisation per frame)
    amz_row_maps_kernel<<<ceil_div(h, 256), 256, 0, st>>>(A.squeezed, A.sq_dst, h, is_bright[0], is_bright[1], is_bright[2], is_bright[3]);
    MLVB_CUDA_OK(cudaMemsetAsync(A.rawf, 0, (size_t)ws * h * 4, st));
    MLVB_CUDA_OK(cudaMemsetAsync(A.counter, 0, sizeof(unsigned), st));
    const dim3 g2(ceil_div(w, 256), h);
    amz_squeeze_kernel<<<g2, 256, 0, st>>>(d_raw32, A.rawf, A.sq_dst, w, h, ws, black);
    const int ntx = amaze::tiles_along(w), nty = amaze::tiles_along(h);
    amz_tiles_kernel<<<A.nblocks, amz_threads(), 0, st>>>(A.rawf, A.red, A.green, A.blue, ws, w, h, ntx, nty, A.ws, A.counter);
    amz_gray_kernel<<<g2, 256, 0, st>>>(A.red, A.green, A.blue, A.squeezed, d_raw2ev, A.grayev, w, h, ws, black);
    amz_edge_dir_kernel<<<dim3(ceil_div(w, 128), h), 128, 0, st>>>(d_raw32, A.grayev, A.edir, w, h, black, white_darkened,
                                                                  is_bright[0], is_bright[1], is_bright[2], is_bright[3], d_fullres_curve,
                                                                  d_fullres_lim);
    if (launches) *launches += 4;
    MLVB_CUDA_OK(cudaGetLastError());
#ifdef AMZ_PROFILE
    {   // debug build only (make EXTRA=-DAMZ_PROFILE): cycles of thread 0 per section of the tile program, summed over tiles
        static int calls = 0;
        if (++calls == 8) {
            unsigned long long prof[32];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(prof, g_amz_prof, sizeof(prof));
            unsigned long long tot = 0;
            for (int k = 0; k < 32; k++) tot += prof[k];
            for (int k = 0; k < 32; k++)
                if (prof[k]) fprintf(stderr, "amz section %2d: %6.2f %%  (%.1f us per tile)\n", k, 100.0 * prof[k] / tot,
                                     prof[k] / 1.965e3 / (8.0 * ntx * nty));
        }
    }
#endif
    return MLVB_OK;
}

// AMaZE alone on a device-resident float mosaic (test / benchmark entry behind mlvb_amaze_demosaic)
int launch_amaze_planes(const float *d_raw, float *d_red, float *d_green, float *d_blue, int stride, int w, int h,
                        char *d_ws, int nblocks, unsigned *d_counter, cudaStream_t st)
{
    if (w & 3) return MLVB_ERR_UNSUPPORTED;
    MLVB_CUDA_OK(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), st));
    amz_tiles_kernel<<<nblocks, amz_threads(), 0, st>>>(d_raw, d_red, d_green, d_blue, stride, w, h, amaze::tiles_along(w),
                                                     amaze::tiles_along(h), d_ws, d_counter);
    MLVB_CUDA_OK(cudaGetLastError());
    return MLVB_OK;
}

size_t amaze_ws_bytes_per_block() { return amaze::WS_BYTES; }
int amaze_tile_count(int w, int h) { return amaze::tiles_along(w) * amaze::tiles_along(h); }

// Drop-in for the reference's exported AMaZE entry (amaze_demosaic_RT.c:113-120, called from hdr.c:1040):
// rawData/red/green/blue are arrays of `winh` row pointers, each row winw+16 floats (hdr.c:967-975).
// Only the whole-image window the reference itself uses (winx = winy = 0) is supported.
extern "C" void amaze_demosaic_RT(float **rawData, float **red, float **green, float **blue, int winx, int winy, int winw, int winh)
{
    mlvb_context *ctx = mlvb_default_context();
    if (!ctx) { fprintf(stderr, "libmlvfs_b200: amaze_demosaic_RT: no CUDA context (no CPU path)\n"); return; }
    if (winx || winy || winw < 32 || winh < 32 || (winw & 3)) {
        fprintf(stderr, "libmlvfs_b200: amaze_demosaic_RT: unsupported window %d,%d %dx%d\n", winx, winy, winw, winh);
        return;
    }
    const int ws = winw + 16;
    const size_t plane = (size_t)ws * winh;
    const int ntiles = amaze_tile_count(winw, winh), nblocks = ntiles < amz_blocks() ? ntiles : amz_blocks();
    const size_t need = 4 * plane * sizeof(float) + 256 + (size_t)nblocks * amaze::WS_BYTES;
    Slot *s = acquire_slot(ctx);
    cudaSetDevice(ctx->device);
    std::vector<float> host(plane);
    bool ok = reserve_device(&s->d_aux, &s->aux_cap, need) == MLVB_OK;
    if (ok) {
        float *d_raw = (float *)s->d_aux, *d_r = d_raw + plane, *d_g = d_r + plane, *d_b = d_g + plane;
        unsigned *d_counter = (unsigned *)(d_b + plane);
        char *d_ws = (char *)d_counter + 256;
        for (int y = 0; y < winh; y++) memcpy(host.data() + (size_t)y * ws, rawData[y], (size_t)ws * sizeof(float));
        ok = cudaMemcpyAsync(d_raw, host.data(), plane * sizeof(float), cudaMemcpyHostToDevice, s->stream) == cudaSuccess &&
             cudaStreamSynchronize(s->stream) == cudaSuccess &&
             launch_amaze_planes(d_raw, d_r, d_g, d_b, ws, winw, winh, d_ws, nblocks, d_counter, s->stream) == MLVB_OK;
        ctx->launches += 1;
        float **dst[3] = {red, green, blue};
        float *src[3] = {d_r, d_g, d_b};
        for (int k = 0; k < 3 && ok; k++) {
            ok = cudaMemcpyAsync(host.data(), src[k], plane * sizeof(float), cudaMemcpyDeviceToHost, s->stream) == cudaSuccess &&
                 cudaStreamSynchronize(s->stream) == cudaSuccess;
            if (ok) for (int y = 0; y < winh; y++) memcpy(dst[k][y], host.data() + (size_t)y * ws, (size_t)winw * sizeof(float));
        }
    }
    if (!ok) { cudaStreamSynchronize(s->stream); fprintf(stderr, "libmlvfs_b200: amaze_demosaic_RT failed\n"); }
    release_slot(ctx, s);
}
