// common.cuh -- shared device/host helpers for libmlvfs_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MLVB_EV_RES 32768               // reference mlvfs.h:87
#define MLVB_MAX_BLACK 16384            // reference mlvfs.h:88
#define MLVB_EV_MAX (14 * MLVB_EV_RES - 1)

#define MLVB_CUDA_OK(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fprintf(stderr, "libmlvfs_b200: %s failed at %s:%d: %s\n", #expr, __FILE__, __LINE__, \
                    cudaGetErrorString(_e));                                                      \
            return MLVB_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

#include "../../include/mlvfs_b200.h"   // MLVB_OK / MLVB_ERR_* status codes, ABI structs

// Device-resident EV tables (reference main.c:128-196), built on the host with libm and uploaded
// once per context so that every entry is bit-identical to what the reference computes.
struct EvLuts {
    // log2 table indexed by d = v - black + MLVB_MAX_BLACK  (32768 ints: 0 below black, INT_MIN at
    // d == 0, (int)(log2(d) * 32768) above) -- same layout as the reference's raw2ev_base.
    const int *raw2ev_base;
    // ev2raw for e in [0, 14*EV): values < 16384, stored as uint16 (896 KiB, L2 resident).
    const uint16_t *ev2raw_pos;
    // full signed table, e in [-10*EV, 14*EV), pointer pre-offset like the reference's (dual ISO)
    const int *ev2raw_full;
};

__device__ __forceinline__ int clamp_ev(int e) { return min(max(e, 0), MLVB_EV_MAX); }

// 32-bit wrap-around arithmetic, matching what the compiled reference does with INT_MIN entries
__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
__device__ __forceinline__ int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
__device__ __forceinline__ int wabs(int a) { return a > 0 ? a : (int)(0u - (unsigned)a); }

// Exact integer <-> double conversions on the fp64 add pipe (I2F.F64 / F2I.F64 run on the quarter-rate conversion
// unit).  0x43300000_xxxxxxxx is 2^52 + x, so one exact subtraction gives x; a signed value goes through x + 2^31.
__device__ __forceinline__ double u2d(unsigned n) { return __hiloint2double(0x43300000, (int)n) - 4503599627370496.0; }
__device__ __forceinline__ double i2d(int n)
{
    return __hiloint2double(0x43300000, (int)((unsigned)n ^ 0x80000000u)) - 4503601774854144.0;     // 2^52 + 2^31
}
// floor(x) for |x| < 2^31: 2^52 + 2^31 + x lies in [2^52, 2^53) where doubles are the integers, round-toward-zero
// of a positive sum drops the fraction.  Equals (int)x for x >= 0, and wherever the result is clamped to >= 0.
__device__ __forceinline__ int d2i_floor(double x)
{
    return (int)((unsigned)__double2loint(__dadd_rz(x, 4503601774854144.0)) ^ 0x80000000u);
}
// a / D for an integer-valued a and a constant D, correctly rounded like the IEEE division it replaces: q = a * R,
// one remainder step with fused multiply-adds (R = 1 / D rounded).  MLVB_FAST_DIV_OK(D, amax) on the host checks every
// numerator the callers can pass (dualiso.cu); the plain division is used when that check fails.
__device__ __forceinline__ double div_const(double a, double D, double R)
{
    const double q = a * R;
    return __fma_rn(__fma_rn(-q, D, a), R, q);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
