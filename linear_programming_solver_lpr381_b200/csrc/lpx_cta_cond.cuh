// lpx_cta_cond.cuh — one CTA per Branch & Bound node LP on a CONDENSED tableau.
//
// Of the n + m + 1 columns of the reference's tableau (R/Models/PrimalSimplex.cs:179-203), m are always
// the basic variables' columns: exact unit vectors that every pivot leaves unchanged (their pivot-row
// entry is 0, so T[i,q] - f * 0 = T[i,q]).  Nothing a decision reads lives there — ChooseEntering can
// never pick a basic column (its objective entry is 0), the ratio test and the Dual Simplex rules read
// only the entering column, the RHS, the objective row and the leaving row's non-basic entries — so
// this kernel stores and updates only the n non-basic columns and the RHS: (m + 1) x (n + 1) doubles
// instead of (m + 1) x (n + m + 1).  At depth 90 of a 60 x 120 problem that is 151 x 121 instead of
// 151 x 271: the node fits one SM's shared memory (no cluster), and a pivot moves 2.2x fewer bytes
// through it — which is what bounds a shared-memory tableau.
//
// A pivot on (row l, slot e): the leaving variable q = basis[l] takes over slot e.  Its column in the
// full tableau is the unit vector of row l, so the slot is first overwritten with that unit vector
// (after the old column has been saved as the update's factors) and then goes through the SAME update
// as every other column: T[l,e] = 1 / piv, T[i,e] = 0 - f_i * (1 / piv) — exactly the operations the
// reference applies to column q.  Every number a decision, x or z depends on is therefore bit-identical
// to the full tableau's for Primal Simplex nodes; for Dual Simplex nodes (whose results B&B discards,
// SURVEY F5) a basic column of the full tableau can carry -0.0 where the unit vector has +0.0, which
// changes no magnitude and no comparison, hence no pivot.
//
// Used for B&B node batches that want no tableau / history output (the throughput path); callbacks and
// histories keep the full-tableau kernels (lpx_cta.cuh, lpx_cta_cluster.cuh).  Decision rules, margins,
// iteration limits and the control-warp pipeline are those of cta_simplex_kernel.
#pragma once
#include "lpx_cta.cuh"

namespace lpx {

struct CondCarve {
    size_t prow, fcol, zc, rbuf, red, rsrc, rsgn, basis, nbvar, slotof, ctl, T, total;
};
__host__ __device__ inline CondCarve cond_carve(int max_rows, int n) {
    const int cols = n + 1, vars = n + max_rows;  // max_rows = m + 1 >= m
    CondCarve c;
    size_t off = 0;
    c.prow = off;
    off += (size_t)((cols + 2) & ~1) * 8;
    c.fcol = off;
    off += (size_t)max_rows * 8;
    c.zc = off;
    off += (size_t)cols * 8;
    c.rbuf = off;  // ratios of the margin scans: by row (primal) or by VARIABLE index (dual)
    off += (size_t)vars * 8;
    c.red = off;
    off += 34 * 16;
    c.rsrc = off;
    off += (size_t)max_rows * 4;
    c.rsgn = off;
    off += (size_t)max_rows * 4;
    c.basis = off;
    off += (size_t)max_rows * 4;
    c.nbvar = off;  // variable held by slot j
    off += (size_t)n * 4;
    c.slotof = off;  // slot of variable v, -1 while it is basic
    off += (size_t)vars * 4;
    c.ctl = off;
    off += 16 * 4;
    off = (off + 15) & ~(size_t)15;
    c.T = off;
    off += (size_t)max_rows * ((cols + 1) & ~1) * 8;
    c.total = off;
    return c;
}

template <int THREADS, bool PROF>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) cta_condensed_kernel(const CtaBatch B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const int p = blockIdx.x;

    const int inst = B.node_inst ? B.node_inst[p] : p;
    const int nex = B.node_extra_cnt ? B.node_extra_cnt[p] : 0;
    const int exo = B.node_extra_off ? B.node_extra_off[p] : 0;
    const int mode = B.node_mode ? B.node_mode[p] : B.mode;
    const int n = B.n;
    const int m = B.m_base + nex;
    const int rows = m + 1, cols = n + 1, ld = (cols + 1) & ~1, rhs = n;

    const CondCarve cv = cond_carve(B.max_rows, n);
    double* prow = reinterpret_cast<double*>(smem_raw + cv.prow);
    double* fcol = reinterpret_cast<double*>(smem_raw + cv.fcol);
    double* zc = reinterpret_cast<double*>(smem_raw + cv.zc);
    double* rbuf = reinterpret_cast<double*>(smem_raw + cv.rbuf);
    ArgMin* red = reinterpret_cast<ArgMin*>(smem_raw + cv.red);
    int* rsrc = reinterpret_cast<int*>(smem_raw + cv.rsrc);
    int* rsgn = reinterpret_cast<int*>(smem_raw + cv.rsgn);
    int* sbasis = reinterpret_cast<int*>(smem_raw + cv.basis);
    int* nbvar = reinterpret_cast<int*>(smem_raw + cv.nbvar);
    int* slotof = reinterpret_cast<int*>(smem_raw + cv.slotof);
    int* ctl = reinterpret_cast<int*>(smem_raw + cv.ctl);
    double* T = reinterpret_cast<double*>(smem_raw + cv.T);

    const double* Ai = B.A + (size_t)inst * B.strideA;
    const double* bi = B.b + (size_t)inst * B.strideB;
    const double* ci = B.c + (size_t)inst * B.strideC;

    // ---- row map + the reference's up-front checks (PrimalSimplex.cs:66-77) ------------------------
    cta_row_map<THREADS>(B, bi, nex, exo, mode, rsrc, rsgn, ctl);
    __syncthreads();
    int status = ctl[0];
    int n_piv = 0, n_silent = 0;
    // LPX_BNB_TRACE: thread 0's clock64() per phase, summed over the launch's CTAs into B.dbg (14 slots)
    const bool prof = PROF && B.dbg != nullptr && tid == 0;
    __shared__ long long pc[PROF ? 10 : 1];
    if (prof)
        for (int k = 0; k < 10; k++) pc[k] = 0;
    long long c0 = prof ? clock64() : 0;
    const long long c_begin = c0;
    int tick_base = 0, n_dual = 0;
#define COND_TICK(slot)                 \
    if (prof) {                         \
        const long long c1 = clock64(); \
        pc[slot] += c1 - c0;            \
        c0 = c1;                        \
    }

    if (status == LPX_RUNNING) {
        // ---- BuildTableau, non-basic columns only: the slack basis starts basic -----------------------
        for (int i = warp; i < rows; i += NW) {
            double* Ti = T + (size_t)i * ld;
            if (i < m) {
                const int src = rsrc[i];
                const bool flip = rsgn[i] != 0;
                const int xv = src < 0 ? B.ex_var[exo + (-1 - src)] : -1;
                const double* Ar = src >= 0 ? Ai + (size_t)src * n : nullptr;
                for (int j = lane; j < ld; j += 32) {
                    double v = 0.0;
                    if (j < n) v = neg_if(Ar ? Ar[j] : (j == xv ? 1.0 : 0.0), flip);
                    else if (j == rhs) v = neg_if(src >= 0 ? bi[src] : B.ex_rhs[exo + (-1 - src)], flip);
                    Ti[j] = v;
                }
            } else {
                for (int j = lane; j < ld; j += 32) {
                    double v = 0.0;
                    if (j < n) {
                        double cj = ci[j];
                        if (B.sense == 1) cj = dneg(cj);  // Min -> Max (PrimalSimplex.cs:62-63)
                        v = dneg(cj);                     // T[m,j] = -C[j]
                    }
                    Ti[j] = v;
                }
            }
        }
        for (int i = tid; i < m; i += THREADS) {
            sbasis[i] = n + i;
            slotof[n + i] = -1;
        }
        for (int j = tid; j < n; j += THREADS) {
            nbvar[j] = j;
            slotof[j] = j;
        }
        __syncthreads();
        COND_TICK(8)

        const double* zrow = T + (size_t)m * ld;

        // most negative objective entry below -1e-9 over the non-basic slots, LOWEST VARIABLE INDEX on ties
        // (the reference scans the columns in variable order; basic columns hold 0 and never win)
        auto entering_from = [&](const double* z) -> int {
            unsigned long long kl = ~0ULL;
            int vl = INT_MAX;
            for (int j = lane; j < n; j += 32) {
                const double zv = z[j];
                if (zv < -LPX_EPS) {
                    const unsigned long long k = dkey(zv);
                    const int v = nbvar[j];
                    if (k < kl || (k == kl && v < vl)) {
                        kl = k;
                        vl = v;
                    }
                }
            }
            const unsigned long long K = warp_min_u64(kl);
            const int v = __reduce_min_sync(0xffffffffu, kl == K ? vl : INT_MAX);
            return K == ~0ULL ? -1 : slotof[v];
        };

        // Gauss-Jordan pivot on (row l, slot e), PrimalSimplex.cs:245-257.  Warp 0 is the control warp: while
        // the others update, it advances its private copy of the objective row, does the basis bookkeeping
        // and (primal steps) picks the next entering column.  Ends with a block barrier.
        auto pivot = [&](int l, int e, bool with_zc) {
            const double piv = T[(size_t)l * ld + e];
            // factors = the entering column as it stands; the pivot row normalised, with the slot's entry taken
            // as 1 (the leaving variable's unit column, which the slot holds from here on)
            for (int i = tid; i < rows; i += THREADS) fcol[i] = T[(size_t)i * ld + e];
            for (int j = tid; j < cols; j += THREADS)
                prow[j] = ddiv_by_pivot(j == e ? 1.0 : T[(size_t)l * ld + j], piv);
            if (tid == 0 && (cols & 1)) prow[cols] = 0.0;
            __syncthreads();
            COND_TICK(tick_base + 2)
            if (warp == 0) {
                if (with_zc) {
                    const double fz = fcol[m];
                    for (int j = lane; j < cols; j += 32)
                        zc[j] = __dsub_rn(j == e ? 0.0 : zc[j], __dmul_rn(fz, prow[j]));
                }
                if (lane == 0) {
                    const int q = sbasis[l], ve = nbvar[e];
                    sbasis[l] = ve;
                    nbvar[e] = q;
                    slotof[ve] = -1;
                    slotof[q] = e;
                    if (B.pivots && n_piv < B.pivots_cap) {
                        int* plog = B.pivots + (size_t)p * B.pivots_cap * 2;
                        plog[2 * n_piv] = ve;  // the reference logs the entering COLUMN = variable index
                        plog[2 * n_piv + 1] = l;
                    }
                }
                __syncwarp();
                if (with_zc) {
                    const int en = entering_from(zc);
                    if (lane == 0) ctl[3] = en;
                }
            } else {
                // two adjacent columns per thread; slot e enters the update as the unit vector of row l
                const int utid = tid - 32, UT = THREADS - 32;
                const int pairs = (cols + 1) >> 1;
                const int cw2 = (pairs + 31) & ~31;
                const int G = cw2 >= UT ? 1 : UT / cw2;
                for (int q0 = 0; q0 < pairs; q0 += UT) {
                    const int g = cw2 >= UT ? 0 : utid / cw2;
                    const int q = cw2 >= UT ? q0 + utid : utid - g * cw2;
                    if (g < G && q < pairs) {
                        const double2 pj = *reinterpret_cast<const double2*>(prow + 2 * q);
                        const int which = (e >> 1) == q ? (e & 1) : -1;  // this thread's pair holds slot e?
                        cta_update_pair_column<true>(T + (size_t)g * ld + 2 * q, (size_t)G * ld, g, G, rows, l, pj, fcol, which);
                    }
                    if (cw2 < UT) break;
                }
            }
            __syncthreads();
            COND_TICK(tick_base + 3)
        };

        // primal pivots (PrimalSimplex.cs:92-124; ForceDualFeasibility, DualSimplex.cs:195-228):
        // 0 = `limit` pivots done, 1 = no entering column, 2 = no leaving row
        auto primal_steps = [&](double margin, int limit, bool silent) -> int {
            if (warp == 0) {
                for (int j = lane; j < cols; j += 32) zc[j] = zrow[j];
                __syncwarp();
                const int e0 = entering_from(zc);
                if (lane == 0) ctl[3] = e0;
            }
            __syncthreads();
            int steps = 0;
            while (true) {
                if (steps >= limit) return 0;
                const int e = ctl[3];
                if (e < 0) return 1;
                cta_stage_ratios<THREADS>(T, ld, m, e, rhs, rbuf);
                COND_TICK(0)
                if (warp == 0) {
                    const int lv = warp_margin_scan_staged(m, margin, rbuf);
                    if (lane == 0) ctl[2] = lv;
                }
                __syncthreads();
                COND_TICK(1)
                const int l = ctl[2];
                if (l < 0) return 2;
                pivot(l, e, true);
                n_piv++;
                if (silent) n_silent++;
                steps++;
            }
        };

        if (mode == 1) primal_steps(LPX_MARGIN_DUAL, 100, true);  // <= 100 silent pivots, ratio margin 1e-12
        if (mode == 0) {
            const int why = primal_steps(LPX_MARGIN_PRIMAL, B.max_iter, false);
            status = why == 0 ? LPX_S_ITER_LIMIT : (why == 1 ? LPX_OPTIMAL : LPX_UNBOUNDED);
        }
        int iter = 1;
        tick_base = 4;
        COND_TICK(9)
        while (mode == 1) {
            if (iter > LPX_DUAL_MAX_ITER) {
                status = LPX_S_ITER_LIMIT;
                break;
            }
            // dual: leaving row = most negative RHS below -1e-9 (DualSimplex.cs:45-55)
            const int l = block_argmin_below_strided<THREADS>(T + rhs, (size_t)ld, m, -LPX_EPS, red);
            COND_TICK(4)
            if (l < 0) {
                status = LPX_OPTIMAL;
                break;
            }
            // entering column: min z_j / (-a) over a < -1e-9 in VARIABLE order, margin 1e-12 (DualSimplex.cs:76-91);
            // the ratios are laid out by variable index (basic variables: not eligible, their a is 0 or 1)
            for (int v = tid; v < n + m; v += THREADS) rbuf[v] = __longlong_as_double(0x7ff8000000000000LL);
            __syncthreads();
            {
                const double* lrow = T + (size_t)l * ld;
                for (int j = tid; j < n; j += THREADS) {
                    const double a = lrow[j];
                    if (a < -LPX_EPS) rbuf[nbvar[j]] = ddiv_by_pivot(zrow[j], dneg(a));
                }
            }
            __syncthreads();
            if (warp == 0) {
                const int ev = warp_margin_scan_cert(n + m, LPX_MARGIN_DUAL, [&](int j, double& r) {
                    r = rbuf[j];
                    return r == r;
                });
                if (lane == 0) ctl[2] = ev < 0 ? -1 : slotof[ev];
            }
            __syncthreads();
            const int e = ctl[2];
            COND_TICK(5)
            if (e < 0) {
                status = LPX_INFEASIBLE;
                break;
            }
            pivot(l, e, false);
            n_piv++;
            n_dual++;
            iter++;
        }
        __syncthreads();

        // ---- FinalizeReport's numeric part (PrimalSimplex.cs:132-138) -------------------------------
        if (B.x) {
            double* xo = B.x + (size_t)p * n;
            for (int j = tid; j < n; j += THREADS) xo[j] = 0.0;
            __syncthreads();
            for (int i = tid; i < m; i += THREADS)
                if (sbasis[i] < n) xo[sbasis[i]] = T[(size_t)i * ld + rhs];
        }
        if (B.z && tid == 0) B.z[p] = T[(size_t)m * ld + rhs];
        if (B.node_flags && B.x && mode == 0) {
            __syncthreads();
            cta_node_epilogue<THREADS>(B, p, inst, nex, exo, B.x + (size_t)p * n, prow, red);
        }
    } else if (B.x) {
        for (int j = tid; j < n; j += THREADS) B.x[(size_t)p * n + j] = 0.0;
        if (B.z && tid == 0) B.z[p] = 0.0;
    }

    if (prof) {
        COND_TICK(9)
        for (int k = 0; k < 10; k++) atomicAdd((unsigned long long*)B.dbg + k, (unsigned long long)pc[k]);
        atomicAdd((unsigned long long*)B.dbg + 10, (unsigned long long)(n_piv - n_dual));
        atomicAdd((unsigned long long*)B.dbg + 11, (unsigned long long)n_dual);
        atomicAdd((unsigned long long*)B.dbg + 12, (unsigned long long)(clock64() - c_begin));
        atomicAdd((unsigned long long*)B.dbg + 13, 1ULL);
    }
#undef COND_TICK
    if (tid == 0) {
        B.status[p] = status;
        if (B.n_pivots) B.n_pivots[p] = n_piv;
        if (B.silent) B.silent[p] = n_silent;
        if (B.n_history) B.n_history[p] = 0;
        if (B.total_pivots && n_piv) atomicAdd(B.total_pivots, (unsigned long long)n_piv);
    }
}

}  // namespace lpx
