// lpx_common.cuh — shared host/device helpers of liblpx (sm_100a only).
//
// Numerics contract: every tableau operation is a separate IEEE binary64
// multiply, subtract or divide (__dmul_rn/__dsub_rn/__ddiv_rn are never contracted into FMA),
// evaluated in the order of the reference's C# loops:
//   ChooseEntering  R/Models/PrimalSimplex.cs:205-220
//   ChooseLeaving   R/Models/PrimalSimplex.cs:222-243   (sequential epsilon-margin scan)
//   Pivot           R/Models/PrimalSimplex.cs:245-257
#pragma once
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/lpx.h"

#define LPX_EPS 1e-9            // PrimalSimplex.Eps / DualSimplex.Eps
#define LPX_MARGIN_PRIMAL 1e-9  // PrimalSimplex.cs:235
#define LPX_MARGIN_DUAL 1e-12   // DualSimplex.cs:85,220

namespace lpx {

// ---- host side error plumbing -------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define LPX_CUDA(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return lpx::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

int ensure_device();  // lazily initialises the library on the current device; LPX_OK or LPX_E_CUDA
int sm_count();
int max_smem_optin();

// ---- device helpers ---------------------------------------------------------------------------
struct ArgMin {
    double v;
    int i;
};

__device__ __forceinline__ ArgMin argmin_pick(ArgMin a, ArgMin b) {
    // smaller value wins; equal values -> lower index (what a left-to-right strict '<' scan keeps)
    return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

__device__ __forceinline__ ArgMin warp_argmin(ArgMin a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ArgMin o;
        o.v = __shfl_xor_sync(0xffffffffu, a.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, a.i, off);
        a = argmin_pick(a, o);
    }
    return a;
}

// Block-wide "most negative entry below thresh, lowest index on ties" over v[0..n).
// Returns -1 when no entry is < thresh.  NaN entries never win (comparison is false).
// red: shared scratch of at least 33 ArgMin.  Ends with a barrier: red may be reused at once.
template <int THREADS>
__device__ __forceinline__ int block_argmin_below(const double* v, int n, double thresh, ArgMin* red) {
    ArgMin a;
    a.v = thresh;
    a.i = INT_MAX;
    for (int j = threadIdx.x; j < n; j += THREADS) {
        double z = v[j];
        if (z < a.v) {
            a.v = z;
            a.i = j;
        }
    }
    a = warp_argmin(a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (warp == 0) {
        ArgMin b;
        b.v = thresh;
        b.i = INT_MAX;
        if (lane < THREADS / 32) b = red[lane];
        b = warp_argmin(b);
        if (lane == 0) red[32] = b;
    }
    __syncthreads();
    int idx = red[32].i;
    __syncthreads();
    return idx == INT_MAX ? -1 : idx;
}

// Same with a stride between consecutive entries (column scans of a row-major tableau).
template <int THREADS>
__device__ __forceinline__ int block_argmin_below_strided(const double* v, size_t stride, int n, double thresh,
                                                          ArgMin* red) {
    ArgMin a;
    a.v = thresh;
    a.i = INT_MAX;
    for (int j = threadIdx.x; j < n; j += THREADS) {
        double z = v[(size_t)j * stride];
        if (z < a.v) {
            a.v = z;
            a.i = j;
        }
    }
    a = warp_argmin(a);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (warp == 0) {
        ArgMin b;
        b.v = thresh;
        b.i = INT_MAX;
        if (lane < THREADS / 32) b = red[lane];
        b = warp_argmin(b);
        if (lane == 0) red[32] = b;
    }
    __syncthreads();
    int idx = red[32].i;
    __syncthreads();
    return idx == INT_MAX ? -1 : idx;
}

// The reference's ratio test is NOT an argmin:
//     if (ratio < bestRatio - margin) { bestRatio = ratio; bestRow = i; }     for i = 0, 1, 2, ...
// One warp reproduces the exact sequential outcome: 32 candidates per step; the first lane that
// beats the running best is accepted, lanes at or before it are retired (they were compared with
// the best of their own time), later lanes are re-tested against the new best.
// get(i, r) -> bool eligible, r = ratio.  Must be called by all 32 lanes of the warp.
template <class Get>
__device__ __forceinline__ int warp_margin_scan(int count, double margin, Get get) {
    const int lane = threadIdx.x & 31;
    double best = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    int row = -1;
    for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        double r = 0.0;
        bool live = false;
        if (i < count) live = get(i, r);
        while (true) {
            const double thr = __dsub_rn(best, margin);
            const unsigned hit = __ballot_sync(0xffffffffu, live && (r < thr));
            if (hit == 0u) break;
            const int first = __ffs(hit) - 1;
            best = __shfl_sync(0xffffffffu, r, first);
            row = base + first;
            if (lane <= first) live = false;
        }
    }
    return row;
}

__device__ __forceinline__ double neg_if(double v, bool flip) {
    // "v *= -1" of the reference (PrimalSimplex.cs:170-171, DualSimplex.cs:135,144-152)
    return flip ? __dmul_rn(v, -1.0) : v;
}

}  // namespace lpx
