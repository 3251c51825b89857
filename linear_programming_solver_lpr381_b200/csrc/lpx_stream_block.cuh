// lpx_stream_block.cuh — blocked look-ahead pivoting for the single large tableau.
//
// Observation: everything pivot k+1 needs to be CHOSEN — the z-row, one column, the RHS, one row —
// can be brought up to date from the tableau as it stood before pivot k plus the rank-1 terms of
// the pivots decided since (f = factor column, p = normalised pivot row), using exactly the
// operations the reference would have applied to those entries, in the same order:
//       v <- (row == l_s) ? p_s[col] : v - f_s[row] * p_s[col]          for s = 1, 2, ...
// So K pivots are decided first by one small look-ahead kernel (it touches K columns and K rows,
// not the tableau), and then ONE pass over HBM applies all K updates to every element, again as
// K separate multiply/subtract pairs in pivot order.  Results are bit-identical to K single
// passes (R/Models/PrimalSimplex.cs:245-257 applied K times); HBM traffic per pivot drops by K.
// The pass stays HBM-bound until 2K flops per element outrun the FP64 pipe (K ~ 20 on B200).
//
// Included by lpx_stream.cu (uses its StreamParams / StreamCtl / block_min_int).
#pragma once

namespace lpx {

#define LPX_BLOCK_KMAX 16

// One CTA, 1024 threads.  Dynamic shared memory: z[ld] | rhs[cs] | ratio[cs] | col[cs].
// budget = 0: probe only (resolve OPTIMAL / UNBOUNDED / ITER_LIMIT for the tableau as it stands).
__global__ void __launch_bounds__(1024) stream_lookahead_kernel(StreamParams P, int budget) {
    extern __shared__ double sm_la[];
    __shared__ ArgMin red[34];
    __shared__ int ired[34];
    __shared__ int s_L[LPX_BLOCK_KMAX];
    __shared__ double s_pe[LPX_BLOCK_KMAX];  // p_s[e]   for the column being brought up to date
    __shared__ double s_fl[LPX_BLOCK_KMAX];  // f_s[l]   for the row being brought up to date
    constexpr int TH = 1024;
    const int tid = threadIdx.x;
    const int ld = P.ld, cs = P.colstride, m = P.m, rows = P.rows, width = P.width;
    double* z = sm_la;
    double* rhs = z + ld;
    double* ratio = rhs + cs;
    double* col = ratio + cs;
    StreamCtl* ctl = P.ctl;
    const int status0 = ctl->status;
    const int done = ctl->pivots;
    if (tid == 0) ctl->block_cnt = 0;
    if (status0 != LPX_RUNNING) return;

    const double* Tz = P.T + (size_t)m * ld;
    for (int j = tid; j < ld; j += TH) z[j] = Tz[j];
    for (int i = tid; i < rows; i += TH) rhs[i] = P.rhsbuf[i];
    __syncthreads();

    const bool probe = budget <= 0;
    const int steps = probe ? 1 : min(budget, P.kblock);
    int cnt = 0, st = LPX_RUNNING;
    for (int k = 0; k < steps; k++) {
        if (done + cnt >= P.max_iter) {  // "if (iter > MaxIterations) throw" precedes the optimality test
            st = LPX_S_ITER_LIMIT;
            break;
        }
        // ---- ChooseEntering on the up-to-date z-row ---------------------------------------------
        const int e = block_argmin_below<TH>(z, width - 1, -LPX_EPS, red);
        if (e < 0) {
            st = LPX_OPTIMAL;
            break;
        }
        // ---- column e: gather from HBM, apply the pivots decided so far, form the ratios ---------
        if (tid < cnt) s_pe[tid] = P.Pbuf[(size_t)tid * ld + e];
        __syncthreads();
        for (int base = 0; base < rows; base += 4 * TH) {
            double c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * TH + tid;
                c[u] = i < rows ? P.T[(size_t)i * ld + e] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = base + u * TH + tid;
                if (i < rows) {
                    double v = c[u];
                    for (int s = 0; s < cnt; s++) {
                        const double ps = s_pe[s];
                        v = (i == s_L[s]) ? ps : __dsub_rn(v, __dmul_rn(P.Fbuf[(size_t)s * cs + i], ps));
                    }
                    col[i] = v;
                    if (i < m) {
                        double r = __longlong_as_double(0x7ff8000000000000LL);
                        if (v > LPX_EPS) r = __ddiv_rn(rhs[i], v);
                        ratio[i] = r;
                    }
                }
            }
        }
        __syncthreads();
        // ---- ChooseLeaving: the exact sequential margin rule (first-hit rounds) -----------------
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int row = -1, start = 0;
        while (true) {
            const double thr = __dsub_rn(best, LPX_MARGIN_PRIMAL);
            int cand = INT_MAX;
            int i = tid;
            if (start > tid) i = tid + ((start - tid + TH - 1) / TH) * TH;
            for (; i < m; i += TH)
                if (ratio[i] < thr) {
                    cand = i;
                    break;
                }
            cand = block_min_int<TH>(cand, ired);
            if (cand == INT_MAX) break;
            best = ratio[cand];
            row = cand;
            start = cand + 1;
        }
        if (row < 0) {
            st = LPX_UNBOUNDED;
            break;
        }
        if (probe) break;
        const int l = row;
        const double piv = col[l], fz = col[m];
        // ---- row l: gather, apply the pivots decided so far, normalise, advance the z-row -------
        if (tid < cnt) s_fl[tid] = P.Fbuf[(size_t)tid * cs + l];
        __syncthreads();
        double* pout = P.Pbuf + (size_t)cnt * ld;
        const double* Tl = P.T + (size_t)l * ld;
        for (int base = 0; base < ld; base += 4 * TH) {
            double r4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = base + u * TH + tid;
                r4[u] = j < ld ? Tl[j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = base + u * TH + tid;
                if (j < ld) {
                    double v = r4[u];
                    for (int s = 0; s < cnt; s++) {
                        const double ps = P.Pbuf[(size_t)s * ld + j];
                        v = (l == s_L[s]) ? ps : __dsub_rn(v, __dmul_rn(s_fl[s], ps));
                    }
                    const double pj = __ddiv_rn(v, piv);
                    pout[j] = pj;
                    if (j < width - 1) z[j] = __dsub_rn(z[j], __dmul_rn(fz, pj));
                }
            }
        }
        // ---- RHS column and bookkeeping ----------------------------------------------------------
        const double prhs = __ddiv_rn(rhs[l], piv);
        __syncthreads();  // every thread has read rhs[l]; z is complete for the next argmin
        double* fout = P.Fbuf + (size_t)cnt * cs;
        for (int i = tid; i < rows; i += TH) {
            const double f = col[i];
            fout[i] = f;
            rhs[i] = (i == l) ? prhs : __dsub_rn(rhs[i], __dmul_rn(f, prhs));
        }
        if (tid == 0) {
            s_L[cnt] = l;
            P.Lbuf[cnt] = l;
            P.basis[l] = e;
            if (done + cnt < P.pivlog_cap) {
                P.pivlog[2 * (done + cnt)] = e;
                P.pivlog[2 * (done + cnt) + 1] = l;
            }
        }
        cnt++;
        __syncthreads();
    }
    for (int i = tid; i < rows; i += TH) P.rhsbuf[i] = rhs[i];
    if (tid == 0) {
        ctl->block_cnt = cnt;
        ctl->pivots = done + cnt;
        if (st != LPX_RUNNING) ctl->status = st;
    }
}

// The HBM pass of a block: every element takes the block's updates in pivot order.
// grid = (column strips of 512, row chunks); p_s[j] for the thread's two columns stay in
// registers, f_s[i] are CTA-uniform loads.  Rows that were a pivot row inside the block take the
// select path; all others (4097 - K of them) the straight multiply/subtract chain.
template <int KMAX, int UNROLL>
__global__ void __launch_bounds__(256, 2) stream_update_block_kernel(StreamParams P) {
    const int cnt = P.ctl->block_cnt;
    if (cnt <= 0) return;
    __shared__ int sL[KMAX];
    if (threadIdx.x < KMAX) sL[threadIdx.x] = threadIdx.x < cnt ? P.Lbuf[threadIdx.x] : -1;
    __syncthreads();
    const int j0 = (blockIdx.x * 256 + threadIdx.x) * 2;
    if (j0 >= P.ld) return;
    const int rpc = (P.rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rpc;
    const int r1 = min(P.rows, r0 + rpc);
    const int n_rows = r1 - r0;
    if (n_rows <= 0) return;
    double2 p[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        p[s].x = 0.0;
        p[s].y = 0.0;
        if (s < cnt) p[s] = *reinterpret_cast<const double2*>(P.Pbuf + (size_t)s * P.ld + j0);
    }
    const size_t ld = (size_t)P.ld, cs = (size_t)P.colstride;
    double* __restrict__ Tc = P.T + j0;
    const double* __restrict__ F = P.Fbuf;
    const bool rev = (P.ctl->pivots & 1) != 0;  // alternate the sweep direction between passes (L2 reuse)

    for (int q = 0; q < n_rows; q += UNROLL) {
        double2 t[UNROLL];
        int ri[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int qq = q + u;
            ri[u] = qq < n_rows ? (rev ? r1 - 1 - qq : r0 + qq) : -1;
            if (ri[u] >= 0) t[u] = *reinterpret_cast<const double2*>(Tc + (size_t)ri[u] * ld);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int i = ri[u];
            if (i < 0) continue;
            bool pivot_row = false;
#pragma unroll
            for (int s = 0; s < KMAX; s++) pivot_row = pivot_row || (i == sL[s]);
            double2 v = t[u];
            if (!pivot_row) {
#pragma unroll
                for (int s = 0; s < KMAX; s++) {
                    if (s < cnt) {
                        const double f = __ldg(F + (size_t)s * cs + i);
                        v.x = __dsub_rn(v.x, __dmul_rn(f, p[s].x));
                        v.y = __dsub_rn(v.y, __dmul_rn(f, p[s].y));
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < KMAX; s++) {
                    if (s < cnt) {
                        const double f = __ldg(F + (size_t)s * cs + i);
                        const bool is_l = i == sL[s];
                        v.x = is_l ? p[s].x : __dsub_rn(v.x, __dmul_rn(f, p[s].x));
                        v.y = is_l ? p[s].y : __dsub_rn(v.y, __dmul_rn(f, p[s].y));
                    }
                }
            }
            *reinterpret_cast<double2*>(Tc + (size_t)i * ld) = v;
        }
    }
}

}  // namespace lpx
