// lpx_reg.cu — register-resident batched Primal Simplex (BASELINE config 2: 64 x 128).
//
// One CTA per tableau, the tableau lives in REGISTERS for the whole solve: warp w owns rows
// w*R .. w*R+R-1, lane l owns columns l, l+32, .., l+32(C-1) of those rows (R*C doubles per
// thread); the RHS column is spread over lanes 0..R-1 of each warp.  Per pivot only three small
// vectors cross shared memory — the entering column (factors), the RHS and the normalised pivot
// row — so shared-memory traffic drops from 2 x 100 KB per pivot (lpx_cta.cuh) to ~3 KB and the
// rank-1 update issues at the FP64 rate: R*C unfused DMUL + DSUB per thread, 97 % of them useful
// at 65 x 193 with NW = 13, R = 5, C = 6.
//
// Same rules, same order, same rounding as the reference (R/Models/PrimalSimplex.cs:205-257):
// entering = warp-shuffle argmin with lowest index on ties, leaving = exact sequential margin
// scan (every warp repeats it on the staged vectors instead of waiting for a broadcast), pivot =
// true division then separate multiply and subtract, zero-factor rows included.
#include "lpx_cta.cuh"
#include "lpx_stream.hpp"

namespace lpx {

struct RegBatch {
    const double* A;
    const double* b;
    const double* c;
    int m, n, sense, max_iter;
    int* status;
    int* n_pivots;
    int* basis;
    double* x;
    double* z;
    double* tableau;
    unsigned long long* total_pivots;
};

template <int NW, int R, int C, int OCC>
__global__ void __launch_bounds__(NW * 32, OCC) reg_simplex_kernel(const RegBatch B) {
    constexpr int ROWS = NW * R;   // padded rows (>= m + 1)
    constexpr int COLS = 32 * C;   // padded non-RHS columns (>= n + m)
    __shared__ double s_f[ROWS];   // entering column = update factors
    __shared__ double s_rhs[ROWS];
    __shared__ double s_p[COLS + 1];    // normalised pivot row, [COLS] = its RHS
    __shared__ double s_raw[COLS + 1];  // the leaving row before normalisation
    __shared__ double s_ratio[ROWS];
    __shared__ int s_basis[ROWS];
    __shared__ int s_ctl[4];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int p = blockIdx.x;
    const int m = B.m, n = B.n, width = n + m + 1;
    const double* Ap = B.A + (size_t)p * m * n;
    const double* bp = B.b + (size_t)p * m;
    const double* cp = B.c + (size_t)p * n;
    const int wz = m / R, rz = m % R;  // owner of the z-row

    // ---- BuildTableau straight into registers (PrimalSimplex.cs:179-203) ---------------------
    double t[R][C];
    double rhsv = 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = w * R + r;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int j = lane + 32 * c;
            double v = 0.0;
            if (i < m) {
                if (j < n) v = Ap[(size_t)i * n + j];
                else if (j == n + i) v = 1.0;
            } else if (i == m && j < n) {
                double cj = cp[j];
                if (B.sense == 1) cj = dneg(cj);
                v = dneg(cj);
            }
            t[r][c] = v;
        }
    }
    if (lane < R && w * R + lane < m) rhsv = bp[w * R + lane];
    // the reference's up-front check (PrimalSimplex.cs:73-76); '>=' rows cannot occur here
    if (tid == 0) s_ctl[1] = LPX_RUNNING;
    for (int i = tid; i < m; i += NW * 32) s_basis[i] = n + i;
    __syncthreads();
    if (lane < R && w * R + lane < m && rhsv < -1e-9) s_ctl[1] = LPX_S_NEG_RHS;
    __syncthreads();
    int status = s_ctl[1];
    int n_piv = 0;

    if (status == LPX_RUNNING) {
        // entering column of the first pivot (ChooseEntering on the initial z-row)
        auto choose_entering = [&]() {
            ArgMin a;
            a.v = -LPX_EPS;
            a.i = INT_MAX;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (r == rz) {
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const double zv = t[r][c];
                        if (zv < a.v) {
                            a.v = zv;
                            a.i = lane + 32 * c;
                        }
                    }
                }
            }
            a = warp_argmin(a);
            if (lane == 0) s_ctl[0] = a.i == INT_MAX ? -1 : a.i;
        };
        if (w == wz) choose_entering();

        // Serial latency, not throughput, bounds a pivot: a double division is a ~40-instruction
        // dependent chain, so every division below is done by a DIFFERENT thread (ratios: one per
        // row, lanes 0..R-1 of each warp; pivot row: one per column), never several by one warp.
        int iter = 1;
        while (true) {
            __syncthreads();  // (A) s_ctl[0] holds the entering column; all updates are done
            if (iter > B.max_iter) {
                status = LPX_S_ITER_LIMIT;
                break;
            }
            const int e = s_ctl[0];
            if (e < 0) {
                status = LPX_OPTIMAL;
                break;
            }
            // ---- entering column -> factors and ratios, inside each warp by shuffle -------------
            const int ce = e >> 5, le = e & 31;
            double my_a = 0.0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < C; c++)
                    if (c == ce) v = t[r][c];
                v = __shfl_sync(0xffffffffu, v, le);
                if (lane == r) my_a = v;
            }
            if (lane < R) {
                const int i = w * R + lane;
                double ratio = __longlong_as_double(0x7ff8000000000000LL);  // NaN = not eligible
                if (i < m && my_a > LPX_EPS) ratio = __ddiv_rn(rhsv, my_a);
                s_f[i] = my_a;
                s_ratio[i] = ratio;
            }
            __syncthreads();  // (B)
            // ---- ChooseLeaving: every warp repeats the exact sequential scan on the ratios -------
            const int lr = warp_margin_scan(m, LPX_MARGIN_PRIMAL, [&](int i, double& ratio) {
                ratio = s_ratio[i];
                return ratio == ratio;
            });
            if (lr < 0) {
                status = LPX_UNBOUNDED;
                break;
            }
            const double piv = s_f[lr];
            // ---- the owner of the leaving row publishes it; one division per thread -------------
            const int wl = lr / R, rl = lr - wl * R;
            if (w == wl) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (r == rl) {
#pragma unroll
                        for (int c = 0; c < C; c++) s_raw[lane + 32 * c] = t[r][c];
                    }
                }
                if (lane == rl) s_raw[COLS] = rhsv;
            }
            __syncthreads();  // (D)
            for (int j = tid; j <= COLS; j += NW * 32) s_p[j] = __ddiv_rn(s_raw[j], piv);
            __syncthreads();  // (E)
            // ---- rank-1 update in registers.  The warp that owns the z-row updates that row
            // first and picks the NEXT entering column while the other warps are still updating.
            double f[R];
#pragma unroll
            for (int r = 0; r < R; r++) f[r] = s_f[w * R + r];
            if (w == wz) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const double pc = s_p[lane + 32 * c];
#pragma unroll
                    for (int r = 0; r < R; r++)
                        if (r == rz) t[r][c] = __dsub_rn(t[r][c], __dmul_rn(f[r], pc));
                }
                choose_entering();
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const double pc = s_p[lane + 32 * c];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (r != rz) {
                            const double upd = __dsub_rn(t[r][c], __dmul_rn(f[r], pc));
                            t[r][c] = (w * R + r == lr) ? pc : upd;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const double pc = s_p[lane + 32 * c];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const double upd = __dsub_rn(t[r][c], __dmul_rn(f[r], pc));
                        t[r][c] = (w * R + r == lr) ? pc : upd;
                    }
                }
            }
            if (lane < R) {
                const double pr = s_p[COLS];
                const double upd = __dsub_rn(rhsv, __dmul_rn(s_f[w * R + lane], pr));
                rhsv = (w * R + lane == lr) ? pr : upd;
            }
            if (tid == 0) s_basis[lr] = e;
            n_piv++;
            iter++;
        }
        __syncthreads();

        // ---- results ----------------------------------------------------------------------------
        if (lane < R) s_rhs[w * R + lane] = rhsv;
        if (B.tableau) {
            double* To = B.tableau + (size_t)p * (m + 1) * width;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = w * R + r;
                if (i <= m) {
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const int j = lane + 32 * c;
                        if (j < width - 1) To[(size_t)i * width + j] = t[r][c];
                    }
                }
            }
            if (lane < R && w * R + lane <= m) To[(size_t)(w * R + lane) * width + width - 1] = rhsv;
        }
        if (B.x) {
            double* xo = B.x + (size_t)p * n;
            for (int j = tid; j < n; j += NW * 32) xo[j] = 0.0;
        }
        __syncthreads();
        if (B.basis)
            for (int i = tid; i < m; i += NW * 32) B.basis[(size_t)p * m + i] = s_basis[i];
        if (tid == 0) {
            if (B.x) {
                double* xo = B.x + (size_t)p * n;
                for (int i = 0; i < m; i++)
                    if (s_basis[i] < n) xo[s_basis[i]] = s_rhs[i];
            }
            if (B.z) B.z[p] = s_rhs[m];
        }
    }
    if (tid == 0) {
        B.status[p] = status;
        if (B.n_pivots) B.n_pivots[p] = n_piv;
        if (B.total_pivots && n_piv) atomicAdd(B.total_pivots, (unsigned long long)n_piv);
    }
}

// Shapes served by the <13, 5, 6> instance: m + 1 <= 65 rows, n + m <= 192 columns, all '<='.
bool reg_kernel_supports(int m, int n, int m_expanded, bool has_rel) {
    if (has_rel || m_expanded != m) return false;
    return m + 1 <= 65 && n + m <= 192 && m >= 1 && n >= 1;
}

int reg_launch_batched(int count, int m, int n, int sense, const double* A, const double* b, const double* c,
                       const lpx_options& opt, int* status, int* n_pivots, int* basis, double* x, double* z,
                       double* tableau, unsigned long long* total_pivots, cudaStream_t stream) {
    if (!reg_kernel_supports(m, n, m, false)) {
        set_error("LPX_KERNEL_CTA_REG: no register-resident kernel is built for this shape (need m <= 64, n + m <= 192, "
                  "all '<=' rows)");
        return LPX_E_CAPACITY;
    }
    if (count <= 0) return LPX_OK;
    RegBatch B;
    B.A = A;
    B.b = b;
    B.c = c;
    B.m = m;
    B.n = n;
    B.sense = sense;
    B.max_iter = opt.max_iterations;
    B.status = status;
    B.n_pivots = n_pivots;
    B.basis = basis;
    B.x = x;
    B.z = z;
    B.tableau = tableau;
    B.total_pivots = total_pivots;
    // Two builds: one CTA per SM with everything in registers (122 registers), or two CTAs per SM
    // with ~11 doubles per thread spilled to L1-resident local memory (72 registers).  The pivot is
    // bound by its serial latency chain, not by issue rate, so the second CTA per SM wins (measured
    // 2.2 ms vs 2.8 ms per 4096-LP batch); reg_variant = 1 forces the spill-free build.
    if (opt.reg_variant == 1) reg_simplex_kernel<13, 5, 6, 1><<<count, 13 * 32, 0, stream>>>(B);
    else reg_simplex_kernel<13, 5, 6, 2><<<count, 13 * 32, 0, stream>>>(B);
    LPX_CUDA(cudaGetLastError());
    count_launch();
    return LPX_OK;
}

}  // namespace lpx
