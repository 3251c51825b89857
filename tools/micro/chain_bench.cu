// Microbenchmark: cycles per step of the knapsack add chain variants (one warp alone on an SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int blocks, double limit) {
    __shared__ double2 st[32];
    const int lane = threadIdx.x;
    st[lane] = make_double2(1.0 + lane * 1e-3, 2.0 + lane * 1e-3);
    __syncwarp();
    double W = 0, P = 0;
    unsigned fail = 0;
    const long long lim = __double_as_longlong(limit);
    long long t0 = clock64();
    for (int b = 0; b < blocks; b++) {
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = st[j0 + j];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                W = __dadd_rn(W, v[j].x);
                P = __dadd_rn(P, v[j].y);
                if (__double_as_longlong(W) > lim) fail |= 1u << (j0 + j);
            }
        }
    }
    long long t1 = clock64();
    double W2 = 0, P2 = 0;
    for (int b = 0; b < blocks; b++) {
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = st[j0 + j];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                W2 = __dadd_rn(W2, v[j].x);
                P2 = __dadd_rn(P2, v[j].y);
            }
        }
    }
    long long t2 = clock64();
    double W3 = 0;
    for (int b = 0; b < blocks; b++) {
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = st[j0 + j];
#pragma unroll
            for (int j = 0; j < 8; j++) W3 = __dadd_rn(W3, v[j].x);
        }
    }
    long long t3 = clock64();
    // packed: both chains in ONE instruction?  (no such FP64 op) -- instead float2-style trick is impossible; test DFMA-free add with constant
    double W4 = 0;
    for (int b = 0; b < blocks * 32; b++) W4 = __dadd_rn(W4, limit);
    long long t4 = clock64();
    if (lane == 0) {
        out[0] = W + P + fail + W2 + P2 + W3 + W4;
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3;
    }
}
int main() {
    double* o; long long* c;
    cudaMalloc(&o, 8); cudaMalloc(&c, 64);
    const int blocks = 200;
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(o, c, blocks, 1e300);
    long long h[4];
    cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    const double steps = blocks * 32.0;
    printf("cycles/step: two chains + int compare %.2f | two chains %.2f | one chain %.2f | one chain, register operand %.2f\n",
           h[0] / steps, h[1] / steps, h[2] / steps, h[3] / steps);
    return 0;
}
