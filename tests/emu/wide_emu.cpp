// tests/emu/wide_emu.cpp -- TEST INFRASTRUCTURE.  Runs the persistent wide fused kernel of
// mlvfs_b200/csrc/fused_wide.cuh (14-bit unpack + 3x3 median chroma smoothing + stripe gains, the work split into
// per-strip runs, the cp.async row staging, the shuffles between lanes) on the host: one std::thread per lane of a
// warp, a std::barrier for __syncwarp / __syncthreads, an exchange array for the shuffles.  The kernel source is
// compiled unchanged (-DMLVB_HOST_EMU only swaps the PTX of the async copies for memcpy and the extern shared array
// for a pointer), with one warp per block (-DFW_WARPS_CFG=1); the blocks of the grid run one after the other.
// Bad-pixel patches: the lists come from the library's own host helper (wide_build_items); the repaired values, which the
// library computes on the device (patch_values_kernel), are handed in by the test (the oracle's).
// Built by tests/test_wide_emu.py with  g++ -O1 -std=c++20 -shared -fPIC -pthread.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <barrier>
#include <thread>
#include <vector>

// ---- what the kernel source needs from CUDA -----------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define MLVB_EV_RES 32768
#define MLVB_EV_MAX (14 * MLVB_EV_RES - 1)

struct Idx3 { unsigned x, y, z; };
static thread_local Idx3 threadIdx, blockIdx, gridDim;
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

struct WarpX {
    std::barrier<> bar{32};
    uint32_t x[32];
};
static thread_local WarpX *g_wx;

template <class T> static inline T min(T a, T b) { return b < a ? b : a; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }
static inline int __vimin3_s32(int a, int b, int c) { return min(min(a, b), c); }
static inline int __vimax3_s32(int a, int b, int c) { return max(max(a, b), c); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __vminu2(uint32_t a, uint32_t b)
{
    return min(a & 0xFFFFu, b & 0xFFFFu) | (min(a >> 16, b >> 16) << 16);
}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
static inline void __syncwarp() { g_wx->bar.arrive_and_wait(); }
static inline void __syncthreads() { g_wx->bar.arrive_and_wait(); }          // one warp per block
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d)
{
    const int lane = threadIdx.x & 31;
    g_wx->x[lane] = (uint32_t)v;
    g_wx->bar.arrive_and_wait();
    const T r = lane >= d ? (T)g_wx->x[lane - d] : v;
    g_wx->bar.arrive_and_wait();
    return r;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d)
{
    const int lane = threadIdx.x & 31;
    g_wx->x[lane] = (uint32_t)v;
    g_wx->bar.arrive_and_wait();
    const T r = lane + d < 32 ? (T)g_wx->x[lane + d] : v;
    g_wx->bar.arrive_and_wait();
    return r;
}
// common.cuh: 32-bit wrap-around arithmetic
static inline int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
static inline int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

#include "../../mlvfs_b200/csrc/fused_wide.cuh"

static_assert(FW_WARPS == 1, "the emulation runs one warp per block");

// One launch of fused3_wide_kernel on `grid` blocks of one warp.  Tables as the library uploads them: raw2ev indexed by
// raw value for this black level (16384 entries), ev2raw13 = ev2raw[13 EV .. 14 EV) as uint16.  coef == NULL: no stripe
// correction.  bad_xy == NULL / n_bad == 0: no repaired pixels.  segments != 0: the equal-segment split with that many
// segments per strip instead of the per-strip runs.
// Returns the STRIPES variant that ran, or < 0 when the frame shape is not eligible (the checks of fused.cu).
extern "C" int wide_emu_run(const uint8_t *packed, size_t payload_stride, uint16_t *out, size_t out_stride_px, int w, int h,
                            int black, int white, int nframes, const int *raw2ev, const uint16_t *ev2raw13, const int *coef,
                            int grid, int segments, const int *bad_xy, int n_bad, const uint16_t *bad_vals)
{
    if ((w % 64) != 0 || ((uintptr_t)out % 16) != 0 || (out_stride_px % 8) != 0 || (payload_stride % 16) != 0 || (h & 1)) return -1;
    WideParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.packed = packed; Q.payload_stride = payload_stride; Q.out = out; Q.out_stride = out_stride_px;
    Q.w = w; Q.h = h; Q.black = black; Q.raw2ev = raw2ev; Q.ev2raw13 = ev2raw13;
    Q.black16 = black; Q.white16 = white;
    int variant = 0;
    if (coef) {
        for (int i = 0; i < 8; i++) {
            if (coef[i] <= 0 || coef[i] >= (1 << 18) || (16383LL - black) * coef[i] + ((long long)black << 16) >= (1LL << 32)) return -2;
            Q.gain[i].coef = (unsigned)coef[i]; Q.gain[i].k1 = 0u - (unsigned)black * (unsigned)coef[i];
            Q.coefh[i] = (unsigned)coef[i] << 14;
            Q.gainx[i].coef = (unsigned)coef[i];
            Q.gainx[i].kx = ((unsigned)black << 16) - (unsigned)black * (unsigned)coef[i];
        }
        variant = coef[0] == 65536 && coef[1] == 65536 && white > black + 64 ? 2 : 1;
    }
    Q.k4 = 0u - 4u * (unsigned)black;
    Q.whitex = ((unsigned)white << 16) | 0xFFFFu;
    Q.nstrips = (w + FW_STRIP_PX - 1) / FW_STRIP_PX; Q.nframes = nframes;
    Q.one = 1; Q.mone = -1;
    Q.shr[0] = 1u << 14; Q.shr[1] = 1u << 17;
    Q.shl[0] = 1u << 14; Q.shl[1] = 1u << 10; Q.shl[2] = 1u << 6; Q.shl[3] = 1u << 2;
    if (segments > 0) { Q.nseg = segments; Q.seg_rows = (h / 2 + segments - 1) / segments; }
    if (grid < Q.nstrips && segments <= 0) return -3;
    // repaired pixels: entry m at (bad_xy[2m], bad_xy[2m + 1]) takes bad_vals[frame * n_bad + m]
    std::vector<WideItem> items;
    std::vector<unsigned> row_start;
    if (n_bad > 0) {
        wide_build_items((size_t)n_bad, [&](size_t m, int &x, int &y) { x = bad_xy[2 * m]; y = bad_xy[2 * m + 1]; }, w, h, items, row_start);
        items.push_back(WideItem{0, 0xFFFF, 0});                       // never read (the kernel stops at row_start); keeps data() non-null
        Q.items = items.data(); Q.row_start = row_start.data(); Q.vals = bad_vals; Q.n_entries = (unsigned)n_bad;
    }

    std::vector<uint8_t> smem(FW_SMEM_BYTES + 64);
    uint8_t *sm = smem.data();
    sm += (16 - ((uintptr_t)sm & 15)) & 15;
    for (int b = 0; b < grid; b++) {
        WarpX wx;
        std::vector<std::thread> lanes;
        for (int t = 0; t < 32; t++)
            lanes.emplace_back([&, t] {
                threadIdx = Idx3{(unsigned)t, 0, 0}; blockIdx = Idx3{(unsigned)b, 0, 0}; gridDim = Idx3{(unsigned)grid, 1, 1};
                g_wx = &wx;
                fw_smem = sm;
                if (variant == 2) fused3_wide_kernel<2>(Q);
                else if (variant == 1) fused3_wide_kernel<1>(Q);
                else fused3_wide_kernel<0>(Q);
            });
        for (auto &th : lanes) th.join();
    }
    return variant;
}
