// Microbenchmark 2: the chain with its real per-block surroundings (double-buffered staging, mask test, fetch, syncwarp).
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void k(const double* gw_s, const double* gp_s, const unsigned* decg, double* out, long long* cyc, int W, int n, double limit) {
    __shared__ double2 bufA[32], bufB[32];
    __shared__ unsigned dec[128];
    const int lane = threadIdx.x;
    __shared__ double sw[2048], sp[2048];
    for (int i = lane; i < W; i += 32) dec[i] = decg[i];
    for (int i = lane; i < 2048; i += 32) { sw[i] = gw_s[i]; sp[i] = gp_s[i]; }
    __syncwarp();
    const double* w_s = (V & 1) ? sw : gw_s;
    const double* p_s = (V & 1) ? sp : gp_s;
    const long long lim_bits = __double_as_longlong(limit);
    double weight = 0, profit = 0;
    auto fetch = [&](int blk, double& wv, double& pv) {
        const int r = (blk << 5) + lane;
        wv = 0.0; pv = 0.0;
        if (blk < W && r >= 0 && r < n) {
            const unsigned m = dec[blk];
            if (!((m >> lane) & 1u)) { wv = w_s[r]; pv = p_s[r]; }
        }
    };
    long long t0 = clock64();
    int b = 0;
    double wn, pn;
    fetch(b, wn, pn);
    bufA[lane] = make_double2(wn, pn);
    fetch(b + 1, wn, pn);
    __syncwarp();
    unsigned anyfail = 0;
    for (; b < W; b++) {
        const double2* cur = (b & 1) ? bufB : bufA;
        ((b & 1) ? bufA : bufB)[lane] = make_double2(wn, pn);
        fetch(b + 2, wn, pn);
        double Wc = weight, Pc = profit;
        unsigned fail = 0u;
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = cur[j0 + j];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                Wc = __dadd_rn(Wc, v[j].x);
                Pc = __dadd_rn(Pc, v[j].y);
                if (__double_as_longlong(Wc) > lim_bits) fail |= 1u << (j0 + j);
            }
        }
        if (fail) { anyfail = fail; break; }
        weight = Wc; profit = Pc;
        if (!(V & 2)) __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = weight + profit + anyfail; cyc[0] = t1 - t0; cyc[1] = b; }
}
int main() {
    const int n = 2000, W = 63;
    double hw[2048], hp[2048]; unsigned hd[64] = {0};
    for (int i = 0; i < 2048; i++) { hw[i] = 1.0 + i * 1e-3; hp[i] = 2.0 + i * 1e-3; }
    double *w, *p, *o; unsigned* d; long long* c;
    cudaMalloc(&w, sizeof hw); cudaMalloc(&p, sizeof hp); cudaMalloc(&d, sizeof hd); cudaMalloc(&o, 8); cudaMalloc(&c, 64);
    cudaMemcpy(w, hw, sizeof hw, cudaMemcpyHostToDevice); cudaMemcpy(p, hp, sizeof hp, cudaMemcpyHostToDevice);
    cudaMemcpy(d, hd, sizeof hd, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 8; rep++) {
        if (rep / 2 == 0) k<0><<<1, 32>>>(w, p, d, o, c, W, n, 1e300);
        if (rep / 2 == 1) k<1><<<1, 32>>>(w, p, d, o, c, W, n, 1e300);
        if (rep / 2 == 2) k<2><<<1, 32>>>(w, p, d, o, c, W, n, 1e300);
        if (rep / 2 == 3) k<3><<<1, 32>>>(w, p, d, o, c, W, n, 1e300);
        long long h[2];
        cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
        printf("variant %d (1: tables in smem, 2: no syncwarp) rep %d: %lld cycles for %lld blocks = %.1f per block, %.2f per step\n", rep / 2, rep, h[0], h[1], (double)h[0] / h[1], (double)h[0] / h[1] / 32);
    }
    return 0;
}
