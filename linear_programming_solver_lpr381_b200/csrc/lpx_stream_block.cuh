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
#include <cooperative_groups.h>

namespace lpx {

#define LPX_BLOCK_KMAX 16

__device__ __forceinline__ unsigned long long lpx_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase timestamps of the first look-ahead step of a launch (development aid; P.dbg may be null)
#define LPX_STAMP(k_) do { if (P.dbg && tid == 0 && cnt == 0) P.dbg[k_] = lpx_gtime(); } while (0)

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
    LPX_STAMP(0);
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
        LPX_STAMP(1);
        // ---- column e: gather from HBM, apply the pivots decided so far, form the ratios ---------
        if (tid < cnt) s_pe[tid] = P.Pbuf[(size_t)tid * ld + e];
        __syncthreads();
        // s is the OUTER loop so that each round issues all of a thread's loads together: with s
        // inside, the f_s[i] loads of one element are serialised by the dependent chain and this
        // phase cost ~1 us per decided pivot and element column
        for (int base = 0; base < rows; base += 8 * TH) {
            double c[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = base + u * TH + tid;
                c[u] = i < rows ? P.T[(size_t)i * ld + e] : 0.0;
            }
            for (int s = 0; s < cnt; s++) {
                const double ps = s_pe[s];
                const int ls = s_L[s];
                double f[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = base + u * TH + tid;
                    f[u] = i < rows ? P.Fbuf[(size_t)s * cs + i] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = base + u * TH + tid;
                    c[u] = (i == ls) ? ps : __dsub_rn(c[u], __dmul_rn(f[u], ps));
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = base + u * TH + tid;
                if (i < rows) {
                    const double v = c[u];
                    col[i] = v;
                    if (i < m) {
                        double r = __longlong_as_double(0x7ff8000000000000LL);
                        if (v > LPX_EPS) r = ddiv_by_pos(rhs[i], v);
                        ratio[i] = r;
                    }
                }
            }
        }
        __syncthreads();
        LPX_STAMP(2);
        // ---- ChooseLeaving: the exact sequential margin rule (first-hit rounds) -----------------
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int row = -1, start = 0;
        int rounds = 0;
        while (true) {
            rounds++;
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
        LPX_STAMP(3);
        if (P.dbg && tid == 0 && cnt == 0) P.dbg[7] = rounds;
        const int l = row;
        const double piv = col[l], fz = col[m];
        // ---- row l: gather, apply the pivots decided so far, normalise, advance the z-row -------
        if (tid < cnt) s_fl[tid] = P.Fbuf[(size_t)tid * cs + l];
        __syncthreads();
        double* pout = P.Pbuf + (size_t)cnt * ld;
        const double* Tl = P.T + (size_t)l * ld;
        for (int base = 0; base < ld; base += 8 * TH) {
            double r8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = base + u * TH + tid;
                r8[u] = j < ld ? Tl[j] : 0.0;
            }
            for (int s = 0; s < cnt; s++) {
                const double fs = s_fl[s];
                const bool same = l == s_L[s];
                double ps[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int j = base + u * TH + tid;
                    ps[u] = j < ld ? P.Pbuf[(size_t)s * ld + j] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) r8[u] = same ? ps[u] : __dsub_rn(r8[u], __dmul_rn(fs, ps[u]));
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = base + u * TH + tid;
                if (j < ld) {
                    const double pj = ddiv_by_pos(r8[u], piv);
                    pout[j] = pj;
                    if (j < width - 1) z[j] = __dsub_rn(z[j], __dmul_rn(fz, pj));
                }
            }
        }
        // ---- RHS column and bookkeeping ----------------------------------------------------------
        const double prhs = ddiv_by_pos(rhs[l], piv);
        __syncthreads();  // every thread has read rhs[l]; z is complete for the next argmin
        LPX_STAMP(4);
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
        LPX_STAMP(5);
        cnt++;
        __syncthreads();
    }
    if (P.dbg && tid == 0) P.dbg[6] = lpx_gtime();
    for (int i = tid; i < rows; i += TH) P.rhsbuf[i] = rhs[i];
    if (tid == 0) {
        ctl->block_cnt = cnt;
        ctl->pivots = done + cnt;
        if (st != LPX_RUNNING) ctl->status = st;
    }
}

// ---- cluster look-ahead ----------------------------------------------------------------------
// One SM cannot keep enough loads in flight: gathering one tableau column (4097 scattered 32-byte
// sectors) or one row (98 KB) from HBM takes 6-7 us from a single CTA, and a look-ahead step
// needs both.  The same step spread over a thread-block cluster of 8 CTAs: CTA r owns rows
// [r*RS, (r+1)*RS) for the column phase and columns [r*CW, (r+1)*CW) for the row phase and the
// z-row; the up-to-date column is all-gathered through distributed shared memory, after which
// every CTA repeats the (cheap, exact) ratio scan on its own copy, so no decision ever has to be
// broadcast.  Three cluster barriers per decided pivot.
//
// The ratio scan itself uses the record property of the reference's rule: a row can only be
// accepted by "ratio < best - 1e-9" if its ratio is a strict prefix minimum, so a
// parallel prefix-min marks the few records (about ln m) and one thread replays the sequential
// rule over them.
#ifndef LPX_LA_CLUSTER
#define LPX_LA_CLUSTER 8  // CTAs of the look-ahead cluster (16, a non-portable size, was tried: the row phase drops from
                         // 5.4 to 3.2 us per pivot but the 16-way all-gather + cluster barrier of the column phase
                         // grows from 5.9 to 8.1 us: 154 vs 146 us per block)
#endif
#define LPX_LA_THREADS 512
#define LPX_LA_RECCAP 1024

struct LaRecord {
    double v;
    int i;
    int pad;
};

// Exact ChooseLeaving over ratio[0..m) (NaN = not eligible) by one CTA of LPX_LA_THREADS threads.
// ratio is stored permuted: element i lives at (i % Q) * TH + i / Q with Q = ceil(m / TH), so that
// thread t reads its Q consecutive rows t*Q .. t*Q+Q-1 without bank conflicts.
__device__ __forceinline__ int la_leaving_scan(const double* ratio, int m, int Q, double* s_wmin, int* s_wcnt,
                                               LaRecord* s_rec, int* s_out, int* s_ired) {
    constexpr int TH = LPX_LA_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    // ---- certified shortcut (see warp_margin_scan64 in lpx_common.cuh): the minimum ratio at its
    // lowest row is the sequential answer when it is the only eligible ratio r with !(min < r - margin).
    // Two block reductions, each published by one barrier and finished by every warp on its own.
    {
        unsigned long long kl = ~0ULL;
        for (int u = 0; u < Q; u++) {
            const int i = tid * Q + u;
            if (i < m) {
                const double r = ratio[u * TH + tid];
                if (r == r) {
                    const unsigned long long k = dkey(r);
                    kl = k < kl ? k : kl;
                }
            }
        }
        const unsigned long long Kw = warp_min_u64(kl);
        if (lane == 0) s_wmin[warp] = __longlong_as_double((long long)Kw);
        __syncthreads();
        const unsigned long long k2 = lane < TH / 32 ? (unsigned long long)__double_as_longlong(s_wmin[lane]) : ~0ULL;
        const unsigned long long K = warp_min_u64(k2);
        if (K == ~0ULL) {
            __syncthreads();  // s_wmin is free again
            return -1;        // no eligible row
        }
        const double vmin = dkey_inv(K);
        int il = INT_MAX, close = 0;
        for (int u = 0; u < Q; u++) {
            const int i = tid * Q + u;
            if (i < m) {
                const double r = ratio[u * TH + tid];
                if (r == r) {
                    if (dkey(r) == K && i < il) il = i;
                    if (!(vmin < __dsub_rn(r, LPX_MARGIN_PRIMAL))) close++;
                }
            }
        }
        const int iw = __reduce_min_sync(0xffffffffu, il);
        const int cw = __reduce_add_sync(0xffffffffu, close);
        if (lane == 0) {
            s_ired[warp] = iw;
            s_wcnt[warp] = cw;
        }
        __syncthreads();
        const int imin = __reduce_min_sync(0xffffffffu, lane < TH / 32 ? s_ired[lane] : INT_MAX);
        const int nclose = __reduce_add_sync(0xffffffffu, lane < TH / 32 ? s_wcnt[lane] : 0);
        __syncthreads();  // the scratch arrays are free again
        if (nclose == 1 && vmin < INF) return imin;
    }
    double lmin = INF;
    for (int u = 0; u < Q; u++) {
        const int i = tid * Q + u;
        double r = i < m ? ratio[u * TH + tid] : INF;
        if (!(r == r)) r = INF;
        lmin = fmin(lmin, r);
    }
    // exclusive prefix-min over threads
    double incl = lmin;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl = fmin(incl, o);
    }
    if (lane == 31) s_wmin[warp] = incl;
    __syncthreads();
    double pm = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) pm = INF;
    for (int w = 0; w < warp; w++) pm = fmin(pm, s_wmin[w]);
    // records of this thread, in row order
    int cnt = 0;
    double run = pm;
    for (int u = 0; u < Q; u++) {
        const int i = tid * Q + u;
        const double r = i < m ? ratio[u * TH + tid] : INF;
        if (r < run) {
            run = r;
            cnt++;
        }
    }
    int incl_c = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl_c, off);
        if (lane >= off) incl_c += o;
    }
    if (lane == 31) s_wcnt[warp] = incl_c;
    __syncthreads();
    int offset = incl_c - cnt;
    int total = 0;
    for (int w = 0; w < TH / 32; w++) {
        const int c = s_wcnt[w];
        if (w < warp) offset += c;
        total += c;
    }
    if (total <= LPX_LA_RECCAP) {
        run = pm;
        for (int u = 0; u < Q; u++) {
            const int i = tid * Q + u;
            const double r = i < m ? ratio[u * TH + tid] : INF;
            if (r < run) {
                run = r;
                s_rec[offset].v = r;
                s_rec[offset].i = i;
                offset++;
            }
        }
        __syncthreads();
        if (tid == 0) {
            double best = INF;
            int row = -1;
            for (int k = 0; k < total; k++) {
                const double r = s_rec[k].v;
                if (r < __dsub_rn(best, LPX_MARGIN_PRIMAL)) {
                    best = r;
                    row = s_rec[k].i;
                }
            }
            *s_out = row;
        }
        __syncthreads();
        return *s_out;
    }
    // more records than the list holds (monotone ratios): first-hit rounds, still exact
    double best = INF;
    int row = -1, start = 0;
    while (true) {
        const double thr = __dsub_rn(best, LPX_MARGIN_PRIMAL);
        int cand = INT_MAX;
        for (int u = 0; u < Q; u++) {
            const int i = tid * Q + u;
            if (i >= start && i < m && ratio[u * TH + tid] < thr) {
                cand = i;
                break;
            }
        }
        cand = block_min_int<TH>(cand, s_ired);
        if (cand == INT_MAX) break;
        best = ratio[(cand % Q) * TH + cand / Q];
        row = cand;
        start = cand + 1;
    }
    return row;
}

__global__ void __cluster_dims__(LPX_LA_CLUSTER, 1, 1) __launch_bounds__(LPX_LA_THREADS)
    stream_lookahead_cluster_kernel(StreamParams P, int budget) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double sm_lc[];
    __shared__ ArgMin red[34];
    __shared__ ArgMin s_part[LPX_LA_CLUSTER];  // every CTA's z-slice argmin, written by its owner
    __shared__ double s_wmin[LPX_LA_THREADS / 32];
    __shared__ int s_wcnt[LPX_LA_THREADS / 32 + 2];
    __shared__ LaRecord s_rec[LPX_LA_RECCAP];
    __shared__ int s_out;
    __shared__ int s_ired[34];
    __shared__ int s_L[LPX_BLOCK_KMAX];
    __shared__ double s_pe[LPX_BLOCK_KMAX];
    __shared__ double s_fl[LPX_BLOCK_KMAX];
    constexpr int TH = LPX_LA_THREADS, CL = LPX_LA_CLUSTER;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int ld = P.ld, cs = P.colstride, m = P.m, rows = P.rows, width = P.width;
    const int RS = (rows + CL - 1) / CL;            // rows per CTA (column phase)
    const int CW = (((ld + CL - 1) / CL) + 1) & ~1;  // columns per CTA (row phase, z slice)
    const int i_lo = rank * RS, i_hi = min(rows, i_lo + RS);
    const int j_lo = rank * CW, j_hi = min(ld, j_lo + CW);
    const int Q = (m + TH - 1) / TH;
    double* zloc = sm_lc;          // CW
    double* rhs = zloc + CW;       // cs, full copy in every CTA
    double* col = rhs + cs;        // cs, full copy (all-gathered)
    double* ratio = col + cs;      // Q * TH, permuted
    StreamCtl* ctl = P.ctl;
    const int status0 = ctl->status;
    const int done = ctl->pivots;
    cluster.sync();  // every CTA has read the control block before rank 0 rewrites it
    if (rank == 0 && tid == 0) ctl->block_cnt = 0;
    if (status0 != LPX_RUNNING) return;

    const double* Tz = P.T + (size_t)m * ld;
    for (int j = j_lo + tid; j < j_hi; j += TH) zloc[j - j_lo] = Tz[j];
    for (int i = tid; i < rows; i += TH) rhs[i] = P.rhsbuf[i];
    __syncthreads();

    const bool probe = budget <= 0;
    const int steps = probe ? 1 : min(budget, P.kblock);
    int cnt = 0, st = LPX_RUNNING;
    for (int k = 0; k < steps; k++) {
        if (done + cnt >= P.max_iter) {
            st = LPX_S_ITER_LIMIT;
            break;
        }
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[0] = lpx_gtime();
        // ---- ChooseEntering: slice argmin, exchanged through distributed shared memory -----------
        {
            ArgMin a;
            a.v = -LPX_EPS;
            a.i = INT_MAX;
            for (int j = j_lo + tid; j < j_hi && j < width - 1; j += TH) {
                const double zv = zloc[j - j_lo];
                if (zv < a.v) {
                    a.v = zv;
                    a.i = j;
                }
            }
            a = warp_argmin(a);
            if (lane == 0) red[warp] = a;
            __syncthreads();
            if (warp == 0) {
                ArgMin b2;
                b2.v = -LPX_EPS;
                b2.i = INT_MAX;
                if (lane < TH / 32) b2 = red[lane];
                b2 = warp_argmin(b2);
                if (lane < CL) {
                    ArgMin* dst = cluster.map_shared_rank(s_part, lane);
                    dst[rank] = b2;
                }
            }
        }
        cluster.sync();  // (1)
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[1] = lpx_gtime();
        ArgMin g = s_part[0];
#pragma unroll
        for (int r = 1; r < CL; r++) g = argmin_pick(g, s_part[r]);
        const int e = g.i == INT_MAX ? -1 : g.i;
        if (e < 0) {
            st = LPX_OPTIMAL;
            break;
        }
        // ---- column e, rows of this CTA: gather, bring up to date, all-gather -------------------
        if (tid < cnt) s_pe[tid] = P.Pbuf[(size_t)tid * ld + e];
        __syncthreads();
        // All factor entries a thread needs are loaded BEFORE the dependent multiply/subtract chain:
        // with the load inside the chain each decided pivot cost one L2 round trip (~0.7 us).
        for (int base = i_lo; base < i_hi; base += 2 * TH) {
            double c[2];
            double f[2][LPX_BLOCK_KMAX];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = base + u * TH + tid;
                c[u] = i < i_hi ? P.T[(size_t)i * ld + e] : 0.0;
            }
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int i = base + u * TH + tid;
                    f[u][s] = (s < cnt && i < i_hi) ? P.Fbuf[(size_t)s * cs + i] : 0.0;
                }
            }
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
                if (s < cnt) {
                    const double ps = s_pe[s];
                    const int ls = s_L[s];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int i = base + u * TH + tid;
                        c[u] = (i == ls) ? ps : __dsub_rn(c[u], __dmul_rn(f[u][s], ps));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = base + u * TH + tid;
                if (i < i_hi) {
#pragma unroll
                    for (int r = 0; r < CL; r++) cluster.map_shared_rank(col, r)[i] = c[u];
                }
            }
        }
        cluster.sync();  // (2) every CTA holds the whole up-to-date column
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[2] = lpx_gtime();
        for (int i = tid; i < Q * TH; i += TH) {
            // permuted slot (u, t) <-> row t*Q + u
            const int t = i % TH, u = i / TH;
            const int row_i = t * Q + u;
            double r = __longlong_as_double(0x7ff8000000000000LL);
            if (row_i < m) {
                const double a = col[row_i];
                if (a > LPX_EPS) r = ddiv_by_pos(rhs[row_i], a);
            }
            ratio[i] = r;
        }
        __syncthreads();
        const int l = la_leaving_scan(ratio, m, Q, s_wmin, s_wcnt, s_rec, &s_out, s_ired);
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[3] = lpx_gtime();
        if (l < 0) {
            st = LPX_UNBOUNDED;
            break;
        }
        if (probe) break;
        const double piv = col[l], fz = col[m];
        // ---- row l, columns of this CTA: gather, bring up to date, normalise, advance z ----------
        if (tid < cnt) s_fl[tid] = P.Fbuf[(size_t)tid * cs + l];
        __syncthreads();
        double* pout = P.Pbuf + (size_t)cnt * ld;
        const double* Tl = P.T + (size_t)l * ld;
        for (int base = j_lo; base < j_hi; base += 2 * TH) {
            double r2[2];
            double pp[2][LPX_BLOCK_KMAX];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int j = base + u * TH + tid;
                r2[u] = j < j_hi ? Tl[j] : 0.0;
            }
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int j = base + u * TH + tid;
                    pp[u][s] = (s < cnt && j < j_hi) ? P.Pbuf[(size_t)s * ld + j] : 0.0;
                }
            }
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
                if (s < cnt) {
                    const double fs = s_fl[s];
                    const bool same = l == s_L[s];
#pragma unroll
                    for (int u = 0; u < 2; u++) r2[u] = same ? pp[u][s] : __dsub_rn(r2[u], __dmul_rn(fs, pp[u][s]));
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int j = base + u * TH + tid;
                if (j < j_hi) {
                    const double pj = ddiv_by_pos(r2[u], piv);
                    pout[j] = pj;
                    if (j < width - 1) zloc[j - j_lo] = __dsub_rn(zloc[j - j_lo], __dmul_rn(fz, pj));
                }
            }
        }
        // ---- RHS (every CTA keeps the full vector), factor column slice, bookkeeping ------------
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[4] = lpx_gtime();
        const double prhs = ddiv_by_pos(rhs[l], piv);
        __syncthreads();
        double* fout = P.Fbuf + (size_t)cnt * cs;
        for (int i = tid; i < rows; i += TH) {
            const double f = col[i];
            if (i >= i_lo && i < i_hi) fout[i] = f;
            rhs[i] = (i == l) ? prhs : __dsub_rn(rhs[i], __dmul_rn(f, prhs));
        }
        if (tid == 0) {
            s_L[cnt] = l;
            if (rank == 0) {
                P.Lbuf[cnt] = l;
                P.basis[l] = e;
                if (done + cnt < P.pivlog_cap) {
                    P.pivlog[2 * (done + cnt)] = e;
                    P.pivlog[2 * (done + cnt) + 1] = l;
                }
            }
        }
        cnt++;
        __threadfence();
        cluster.sync();  // (3) Pbuf / Fbuf slices of this pivot are visible to the whole cluster
        if (P.dbg && rank == 0 && tid == 0 && cnt == 3) P.dbg[5] = lpx_gtime();
    }
    if (rank == 0) {
        for (int i = tid; i < rows; i += TH) P.rhsbuf[i] = rhs[i];
        if (tid == 0) {
            ctl->block_cnt = cnt;
            ctl->pivots = done + cnt;
            if (st != LPX_RUNNING) ctl->status = st;
        }
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

// The HBM pass of a block: every element takes the block's updates in pivot order.
// grid = (column strips of 256 * VEC, row chunks).  p_s[j] for the thread's VEC columns stay in
// registers; the factor entries f_s[i] of the CTA's row chunk are staged once in shared memory
// (K x rows-per-CTA doubles), so the inner loop is one broadcast LDS plus an unfused multiply and
// subtract per pivot and element, and the only global traffic is the tableau itself.
// Rows that were a pivot row inside the block take the select path; all others the straight chain.
template <int KMAX, int VEC, int UNROLL>
__global__ void __launch_bounds__(256) stream_update_block_kernel(StreamParams P, int rpc) {
    extern __shared__ double sF[];  // [s][ii], ii < rpc
    __shared__ int sL[KMAX];
    const int cnt = P.ctl->block_cnt;
    if (cnt <= 0) return;
    const int r0 = blockIdx.y * rpc;
    const int r1 = min(P.rows, r0 + rpc);
    const int n_rows = r1 - r0;
    if (n_rows <= 0) return;
    unsigned char* sPiv = reinterpret_cast<unsigned char*>(sF + (size_t)KMAX * rpc);
    if (threadIdx.x < KMAX) sL[threadIdx.x] = threadIdx.x < cnt ? P.Lbuf[threadIdx.x] : -1;
    __syncthreads();
    const size_t cs = (size_t)P.colstride;
    for (int s = 0; s < cnt; s++)
        for (int ii = threadIdx.x; ii < n_rows; ii += 256) sF[s * rpc + ii] = P.Fbuf[(size_t)s * cs + r0 + ii];
    for (int ii = threadIdx.x; ii < n_rows; ii += 256) {
        bool pv = false;
#pragma unroll
        for (int s = 0; s < KMAX; s++) pv = pv || (r0 + ii == sL[s]);
        sPiv[ii] = pv ? 1 : 0;
    }
    __syncthreads();
    const int j0 = (blockIdx.x * 256 + threadIdx.x) * VEC;
    if (j0 >= P.ld) return;
    double p[KMAX][VEC];
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
#pragma unroll
        for (int v = 0; v < VEC; v++) p[s][v] = s < cnt ? P.Pbuf[(size_t)s * P.ld + j0 + v] : 0.0;
    }
    const size_t ld = (size_t)P.ld;
    double* __restrict__ Tc = P.T + j0;
    const bool rev = ((P.ctl->pivots / max(P.kblock, 1)) & 1) != 0;  // alternate the sweep direction (L2 reuse)

    for (int q = 0; q < n_rows; q += UNROLL) {
        double t[UNROLL][VEC];
        int ri[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int qq = q + u;
            ri[u] = qq < n_rows ? (rev ? n_rows - 1 - qq : qq) : -1;
            if (ri[u] >= 0) {
                const double* src = Tc + (size_t)(r0 + ri[u]) * ld;
                if (VEC == 2) {
                    const double2 d = *reinterpret_cast<const double2*>(src);
                    t[u][0] = d.x;
                    t[u][VEC - 1] = d.y;
                } else {
                    t[u][0] = *src;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int ii = ri[u];
            if (ii < 0) continue;
            if (!sPiv[ii]) {
#pragma unroll
                for (int s = 0; s < KMAX; s++) {
                    if (s < cnt) {
                        const double f = sF[s * rpc + ii];
#pragma unroll
                        for (int v = 0; v < VEC; v++) t[u][v] = __dsub_rn(t[u][v], __dmul_rn(f, p[s][v]));
                    }
                }
            } else {
                const int i = r0 + ii;
#pragma unroll
                for (int s = 0; s < KMAX; s++) {
                    if (s < cnt) {
                        const double f = sF[s * rpc + ii];
                        const bool is_l = i == sL[s];
#pragma unroll
                        for (int v = 0; v < VEC; v++)
                            t[u][v] = is_l ? p[s][v] : __dsub_rn(t[u][v], __dmul_rn(f, p[s][v]));
                    }
                }
            }
            double* dst = Tc + (size_t)(r0 + ii) * ld;
            if (VEC == 2) {
                double2 d;
                d.x = t[u][0];
                d.y = t[u][VEC - 1];
                *reinterpret_cast<double2*>(dst) = d;
            } else {
                *dst = t[u][0];
            }
        }
    }
}

// ---- TMA-staged variant of the blocked pass -----------------------------------------------------
// Same arithmetic, different data movement: row segments are brought into shared memory by the
// bulk-copy engine (cp.async.bulk, completion on an mbarrier), updated in place by the compute
// threads, and written back with cp.async.bulk shared->global.  Loads in flight no longer depend on
// registers or resident warps (4 stages x 16 KB per CTA), the factor entries of a tile arrive with it
// (K copies of 64 B), and the compute threads only issue LDS / DMUL / DSUB / STS.
//   tile  = 8 rows x 256 columns (2 KB per row segment) + f[K][8]
//   CTA   = 256 threads: thread t owns the column pair t % 128 of rows 4*(t / 128) .. +3 of a tile
//   grid  = (column strips of 256, row chunks) sized to one resident wave
#define LPX_TMA_ROWS 8
#define LPX_TMA_COLS 256
#define LPX_TMA_STAGES 4

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}

template <int KMAX>
struct TmaTile {
    double t[LPX_TMA_ROWS][LPX_TMA_COLS];  // 16 KB
    double f[KMAX][LPX_TMA_ROWS];          // factor entries of the tile's rows
};

template <int KMAX>
__global__ void __launch_bounds__(256) stream_update_block_tma_kernel(StreamParams P, int rpc) {
    extern __shared__ __align__(128) unsigned char smem_tma[];
    __shared__ __align__(8) unsigned long long full[LPX_TMA_STAGES];
    __shared__ int sL[KMAX];
    constexpr int TR = LPX_TMA_ROWS, TC = LPX_TMA_COLS, ST = LPX_TMA_STAGES;
    typedef TmaTile<KMAX> Tile;
    Tile* tiles = reinterpret_cast<Tile*>(smem_tma);
    const int cnt = P.ctl->block_cnt;
    if (cnt <= 0) return;
    const int tid = threadIdx.x;
    const int r0 = blockIdx.y * rpc;
    const int r1 = min(P.rows, r0 + rpc);
    const int jbase = blockIdx.x * TC;
    if (r1 <= r0 || jbase >= P.ld) return;
    const int seg = min(TC, P.ld - jbase);            // columns of this strip (multiple of 16)
    const unsigned seg_bytes = (unsigned)seg * 8;
    const int n_tiles = (r1 - r0 + TR - 1) / TR;
    const size_t ld = (size_t)P.ld, cs = (size_t)P.colstride;

    if (tid < KMAX) sL[tid] = tid < cnt ? P.Lbuf[tid] : -1;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer side (thread 0): one tile = its row segments + K slices of the factor columns
    auto issue_load = [&](int k) {
        Tile& tl = tiles[k % ST];
        const int rr = r0 + k * TR;
        const int nr = min(TR, r1 - rr);
        // factor slices: 8 doubles per pivot; the vectors are padded so reading past `rows` is safe
        mbar_expect_tx(&full[k % ST], (unsigned)nr * seg_bytes + (unsigned)cnt * TR * 8);
        for (int i = 0; i < nr; i++) bulk_g2s(&tl.t[i][0], P.T + (size_t)(rr + i) * ld + jbase, seg_bytes, &full[k % ST]);
        for (int s = 0; s < cnt; s++) bulk_g2s(&tl.f[s][0], P.Fbuf + (size_t)s * cs + rr, TR * 8, &full[k % ST]);
    };
    if (tid == 0)
        for (int k = 0; k < ST - 1 && k < n_tiles; k++) issue_load(k);

    // this thread's columns and pivot-row entries
    const int cp2 = (tid & 127) * 2;       // column pair inside the strip
    const int rsub = (tid >> 7) * 4;       // first of its 4 rows inside a tile
    const bool col_ok = cp2 < seg;
    double p[KMAX][2];
#pragma unroll
    for (int s = 0; s < KMAX; s++) {
        p[s][0] = (s < cnt && col_ok) ? P.Pbuf[(size_t)s * ld + jbase + cp2] : 0.0;
        p[s][1] = (s < cnt && col_ok) ? P.Pbuf[(size_t)s * ld + jbase + cp2 + 1] : 0.0;
    }

    for (int k = 0; k < n_tiles; k++) {
        Tile& tl = tiles[k % ST];
        mbar_wait(&full[k % ST], (unsigned)((k / ST) & 1));
        const int rr = r0 + k * TR;
        const int nr = min(TR, r1 - rr);
        bool has_pivot_row = false;
#pragma unroll
        for (int s = 0; s < KMAX; s++) has_pivot_row = has_pivot_row || (sL[s] >= rr && sL[s] < rr + nr);
        if (col_ok) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int i = rsub + q;
                if (i < nr) {
                    double2 v = *reinterpret_cast<double2*>(&tl.t[i][cp2]);
                    if (!has_pivot_row) {
#pragma unroll
                        for (int s = 0; s < KMAX; s++) {
                            if (s < cnt) {
                                const double f = tl.f[s][i];
                                v.x = __dsub_rn(v.x, __dmul_rn(f, p[s][0]));
                                v.y = __dsub_rn(v.y, __dmul_rn(f, p[s][1]));
                            }
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < KMAX; s++) {
                            if (s < cnt) {
                                const double f = tl.f[s][i];
                                const bool is_l = (rr + i) == sL[s];
                                v.x = is_l ? p[s][0] : __dsub_rn(v.x, __dmul_rn(f, p[s][0]));
                                v.y = is_l ? p[s][1] : __dsub_rn(v.y, __dmul_rn(f, p[s][1]));
                            }
                        }
                    }
                    *reinterpret_cast<double2*>(&tl.t[i][cp2]) = v;
                }
            }
        }
        // make the generic-proxy writes visible to the bulk-copy engine, then write the tile back
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < nr; i++) bulk_s2g(P.T + (size_t)(rr + i) * ld + jbase, &tl.t[i][0], seg_bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the stage of tile k-1 may be refilled once its write-back has finished reading shared memory
            const int nxt = k + ST - 1;
            if (nxt < n_tiles) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue_load(nxt);
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace lpx
