// amaze.cuh -- interface of amaze.cu (AMaZE + edge-direction stage of the dual-ISO pipeline, hdr.c:954-1229)
#pragma once
#include <vector>

#include "common.cuh"

// launch bounds of the tile kernel: 256 threads, 5 blocks per SM = 48 registers (216 bytes of spill, L1-resident).
// Measured on C4: 64 registers x 4 blocks 162 frames/s; 48 x 4 blocks 171; 48 x 5 169; 40 x 6 170 -- the smaller
// register file share leaves room for the other batch lanes' kernels next to the tile programs.
#ifndef AMZ_THREADS_MAX
#define AMZ_THREADS_MAX 256
#endif
#ifndef AMZ_MIN_BLOCKS
#define AMZ_MIN_BLOCKS 5
#endif
#define AMZ_MAX_BLOCKS (148 * 6)      // upper bound of persistent tile blocks; each owns a 2.3 MB workspace
int amz_threads();                    // threads per tile block and tile blocks per SM (tunable: MLVB_AMZ_THREADS,
int amz_blocks();                     // MLVB_AMZ_BLOCKS_PER_SM), chosen for latency hiding on the global workspace

struct AmazeScratch {
    float *rawf, *red, *green, *blue;  // h rows of w+16 floats (hdr.c:967-975)
    int *grayev;                       // raw2ev[gray] of the de-squeezed grayscale image (hdr.c:1059-1062, :1148)
    uint8_t *edir;                     // best direction per pixel (hdr.c:1065-1173)
    int *squeezed, *sq_dst;            // row maps of the squeeze (hdr.c:959-1026)
    unsigned *counter;                 // tile queue head
    char *ws;                          // nblocks tile workspaces
    int nblocks;
};

// carves `base` (may be null: size query); returns the bytes needed
size_t amaze_scratch_bytes(int w, int h, AmazeScratch *S, uint8_t *base);
int launch_amaze_stage(const uint32_t *d_raw32, int w, int h, int black, int white_darkened, const int is_bright[4],
                       const int *d_raw2ev, const double *d_fullres_curve, const int *d_fullres_lim, const AmazeScratch &A,
                       cudaStream_t st, int *launches);
int launch_amaze_planes(const float *d_raw, float *d_red, float *d_green, float *d_blue, int stride, int w, int h,
                        char *d_ws, int nblocks, unsigned *d_counter, cudaStream_t st);
size_t amaze_ws_bytes_per_block();
int amaze_tile_count(int w, int h);

#ifdef __CUDACC__
// edge_directions[] of hdr.c:917-938 as {ack, a, b, bck} x {dx, dy}; y offsets are multiplied by s
static __constant__ int c_edge[11][8] = {
    {-4, 2, -2, 1,  4, -2,  6, -3}, {-3, 2, -1, 1,  3, -2,  4, -3}, {-2, 2, -1, 1,  2, -2,  3, -3}, {-1, 2, -1, 1,  1, -2,  2, -3},
    {-1, 2,  0, 1,  1, -2,  1, -3}, { 0, 2,  0, 1,  0, -2,  0, -3}, { 1, 2,  0, 1, -1, -2, -1, -3}, { 1, 2,  1, 1, -1, -2, -2, -3},
    { 2, 2,  1, 1, -2, -2, -3, -3}, { 3, 2,  1, 1, -3, -2, -4, -3}, { 4, 2,  2, 1, -4, -2, -6, -3}};

// post-scaling of the AMaZE planes (hdr.c:1045-1053), applied on the fly by the readers
__device__ __forceinline__ float amz_post_green(float g, int black)
{
    g = (g - (float)black) * 2.0f + (float)black;
    g = g < 1048575.0f ? g : 1048575.0f;
    return g > 0.0f ? g : 0.0f;
}
__device__ __forceinline__ float amz_post_rb(float v)
{
    v = v < 1048575.0f ? v : 1048575.0f;
    return v > 0.0f ? v : 0.0f;
}

struct AmazeView { const float *red, *green, *blue; const int *squeezed; const uint8_t *edir; int ws; };

// edge_interp (hdr.c:940-952): colour plane of pixel (x, y), two taps along direction `dir`, mixed 2:1 in EV
__device__ __forceinline__ int amz_edge_interp(const AmazeView &A, const int *__restrict__ raw2ev, int dir, int x, int y, int s, int black)
{
    const bool is_rg = (y & 1) == 0, xe = (x & 1) == 0;
    const bool is_green = is_rg != xe;
    const float *plane = is_green ? A.green : (is_rg ? A.red : A.blue);
    const float va = plane[(size_t)A.squeezed[y + c_edge[dir][3] * s] * A.ws + x + c_edge[dir][2]];
    const float vb = plane[(size_t)A.squeezed[y + c_edge[dir][5] * s] * A.ws + x + c_edge[dir][4]];
    const int pa = min(max((int)(is_green ? amz_post_green(va, black) : amz_post_rb(va)), 0), 0xFFFFF);
    const int pb = min(max((int)(is_green ? amz_post_green(vb, black) : amz_post_rb(vb)), 0), 0xFFFFF);
    return (__ldg(raw2ev + pa) * 2 + __ldg(raw2ev + pb)) / 3;
}
#endif
