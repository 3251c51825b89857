// lpx_bench.cu — measurement helpers exported for bench.py (not used by any solve path).
//
// lpx_measure_fp64_rate: the roofline denominator for the shared-memory / register resident
// batched kernels is the UNFUSED FP64 rate (one DMUL + one DADD per tableau element; FMA is
// forbidden by the bit-exactness contract).  MEASURED_PEAKS.json has no FP64 figure, so it is
// measured here: 8 independent mul/sub chains per thread, full occupancy.
#include "lpx_common.cuh"
#include "lpx_runtime.hpp"

namespace lpx {

__global__ void __launch_bounds__(256) fp64_rate_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double f = 1.0 + 1e-9 * seed, p = 1e-7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        a0 = __dsub_rn(a0, __dmul_rn(f, a1 * p));
        a1 = __dsub_rn(a1, __dmul_rn(f, a2));
        a2 = __dsub_rn(a2, __dmul_rn(f, a3));
        a3 = __dsub_rn(a3, __dmul_rn(f, a4));
        a4 = __dsub_rn(a4, __dmul_rn(f, a5));
        a5 = __dsub_rn(a5, __dmul_rn(f, a6));
        a6 = __dsub_rn(a6, __dmul_rn(f, a7));
        a7 = __dsub_rn(a7, __dmul_rn(f, a0));
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

// Dependent-chain latencies (cycles per operation, one warp alone on an SM): what bounds one pivot of
// a per-tableau kernel.  out[0..7] = DADD, DMUL, DDIV, REDUX.MIN u32, SHFL, LDS round trip (store +
// barrier-free load), BAR.SYNC (13 warps), DSETP+SEL.
__global__ void __launch_bounds__(416) latency_kernel(double* out, int iters, double seed) {
    __shared__ double s_v[64];
    const int lane = threadIdx.x & 31;
    long long t0, t1;
    double a = seed + 1.0, b = 1.0 + 1e-9 * seed;
    unsigned u = (unsigned)(seed * 1000.0) + lane;
    double res[8];
    // all warps take part in the barrier test, warp 0 lane 0 reports
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = __dadd_rn(a, b);
    t1 = clock64();
    res[0] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = __dmul_rn(a, b);
    t1 = clock64();
    res[1] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = __ddiv_rn(a, b);
    t1 = clock64();
    res[2] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) u = __reduce_min_sync(0xffffffffu, u + lane) + 1u;
    t1 = clock64();
    res[3] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) u = __shfl_xor_sync(0xffffffffu, u, 1) + 1u;
    t1 = clock64();
    res[4] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) {
        s_v[lane] = a;
        __syncwarp();
        a = s_v[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64();
    res[5] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    t1 = clock64();
    res[6] = (double)(t1 - t0) / iters;
    t0 = clock64();
    for (int i = 0; i < iters; i++) a = (a < b) ? b : __dadd_rn(a, -0.5);
    t1 = clock64();
    res[7] = (double)(t1 - t0) / iters;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 8; k++) out[k] = res[k];
        out[8] = a + (double)u;
    }
}

}  // namespace lpx

using namespace lpx;

extern "C" int lpx_measure_latencies(double* cycles8) {
    if (!cycles8) return LPX_E_BAD_ARGS;
    int rc = ensure_device();
    if (rc != LPX_OK) return rc;
    Runtime& r = rt();
    double* d = ws_dev_as<double>(WS_MISC0, 16);
    if (!d) return LPX_E_CUDA;
    for (int rep = 0; rep < 2; rep++) latency_kernel<<<1, 416, 0, r.stream>>>(d, 2000, 1.0 + rep);
    LPX_CUDA(cudaMemcpyAsync(cycles8, d, 8 * sizeof(double), cudaMemcpyDeviceToHost, r.stream));
    LPX_CUDA(cudaStreamSynchronize(r.stream));
    return LPX_OK;
}

extern "C" int lpx_measure_fp64_rate(double* tflops) {
    if (!tflops) return LPX_E_BAD_ARGS;
    int rc = ensure_device();
    if (rc != LPX_OK) return rc;
    Runtime& r = rt();
    double* d = ws_dev_as<double>(WS_MISC0, 16);
    if (!d) return LPX_E_CUDA;
    const int blocks = r.sms * 8, iters = 1 << 14;
    cudaEvent_t e0, e1;
    LPX_CUDA(cudaEventCreate(&e0));
    LPX_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        LPX_CUDA(cudaEventRecord(e0, r.stream));
        fp64_rate_kernel<<<blocks, 256, 0, r.stream>>>(d, iters, 1.0 + rep);
        LPX_CUDA(cudaEventRecord(e1, r.stream));
        LPX_CUDA(cudaStreamSynchronize(r.stream));
        float ms = 0;
        LPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        // 17 DP instructions per iteration (16 in the chains + the a1*p product)
        const double flops = (double)blocks * 256 * iters * 17.0;
        const double t = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return LPX_OK;
}
