// lpx_cta_cluster.cuh — one thread-block CLUSTER per tableau: the same whole-solve kernel as
// lpx_cta.cuh for tableaux that do not fit the shared memory of one SM but do fit that of 2 or 4.
//
// Branch & Bound nodes grow by one row and one column per level (R/Models/Branch&Bound.cs:233-248):
// at depth 60 a 60 x 120 base problem is a 121 x 241 tableau, 233 KB, and the per-CTA kernel falls
// back to a tableau in global memory — 6x slower per pivot, HBM bound once a few hundred nodes are in
// flight.  Here the ROWS are split over the CTAs of a cluster (CTA r owns rows r*H .. r*H+H-1 in its
// own shared memory), the small vectors every decision needs are replicated, and the three exchanges
// of a pivot go through distributed shared memory:
//   primal: entering column e (owner of the objective row -> all) | ratios of own rows -> all, every
//           CTA repeats the exact margin scan | normalised pivot row (owner of row l -> all)
//   dual:   leaving-row candidates (all -> all) | raw row l (its owner -> all) | entering ratios
//           (owner of the objective row -> all); the pivot row is normalised locally.
// Decisions are taken from replicated data, so control flow is uniform across the cluster.
// Arithmetic, order and rounding are those of lpx_cta.cuh (PrimalSimplex.cs:205-257,
// DualSimplex.cs:45-113, 195-246).
#pragma once
#include <cooperative_groups.h>

#include "lpx_cta.cuh"

namespace lpx {

struct CtaClusterCarve {
    size_t prow, ratio, fcol, zc, red, part, rsrc, rsgn, basis, ctl, T, total;
    int H;  // rows per CTA
};
__host__ __device__ inline CtaClusterCarve cta_cluster_carve(int max_rows, int max_width, int cl) {
    CtaClusterCarve c;
    c.H = (max_rows + cl - 1) / cl;
    const int vec = max_rows > max_width ? max_rows : max_width;
    size_t off = 0;
    c.prow = off;
    off += (size_t)((max_width + 2) & ~1) * 8;  // even length: the update reads it two entries at a time
    c.ratio = off;
    off += (size_t)vec * 8;
    c.fcol = off;
    off += (size_t)c.H * 8;
    c.zc = off;  // every CTA's control warp keeps its own copy of the whole objective row
    off += (size_t)max_width * 8;
    c.red = off;
    off += 34 * 16;
    c.part = off;
    off += 8 * 16;
    c.rsrc = off;
    off += (size_t)max_rows * 4;
    c.rsgn = off;
    off += (size_t)max_rows * 4;
    c.basis = off;
    off += (size_t)max_rows * 4;
    c.ctl = off;
    off += 16 * 4;
    off = (off + 15) & ~(size_t)15;
    c.T = off;
    off += (size_t)c.H * ((max_width + 1) & ~1) * 8;  // rows padded to an even length (16-byte aligned)
    c.total = off;
    return c;
}

template <int THREADS, int CL>
__global__ void __launch_bounds__(THREADS) cta_cluster_simplex_kernel(const CtaBatch B) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const int rank = (int)cluster.block_rank();
    const int p = blockIdx.x / CL;

    const int inst = B.node_inst ? B.node_inst[p] : p;
    const int nex = B.node_extra_cnt ? B.node_extra_cnt[p] : 0;
    const int exo = B.node_extra_off ? B.node_extra_off[p] : 0;
    const int mode = B.node_mode ? B.node_mode[p] : B.mode;
    const int n = B.n;
    const int m = B.m_base + nex;
    const int rows = m + 1, width = n + m + 1, ld = (width + 1) & ~1;
    const int rhs = width - 1;
    const int H = (rows + CL - 1) / CL;                  // rows per CTA for THIS tableau
    const int r_lo = min(rows, rank * H), r_hi = min(rows, r_lo + H);
    const int zr = m / H;                                // owner of the objective row

    const CtaClusterCarve cv = cta_cluster_carve(B.max_rows, B.max_width, CL);
    double* prow = reinterpret_cast<double*>(smem_raw + cv.prow);
    double* ratio = reinterpret_cast<double*>(smem_raw + cv.ratio);
    double* fcol = reinterpret_cast<double*>(smem_raw + cv.fcol);
    double* zc = reinterpret_cast<double*>(smem_raw + cv.zc);
    ArgMin* red = reinterpret_cast<ArgMin*>(smem_raw + cv.red);
    ArgMin* part = reinterpret_cast<ArgMin*>(smem_raw + cv.part);
    int* rsrc = reinterpret_cast<int*>(smem_raw + cv.rsrc);
    int* rsgn = reinterpret_cast<int*>(smem_raw + cv.rsgn);
    int* sbasis = reinterpret_cast<int*>(smem_raw + cv.basis);
    int* ctl = reinterpret_cast<int*>(smem_raw + cv.ctl);
    double* T = reinterpret_cast<double*>(smem_raw + cv.T);  // local rows r_lo .. r_hi-1
    auto loc = [&](int i) -> double* { return T + (size_t)(i - r_lo) * ld; };
    auto owner = [&](int i) -> int { return i / H; };

    const double* Ai = B.A + (size_t)inst * B.strideA;
    const double* bi = B.b + (size_t)inst * B.strideB;
    const double* ci = B.c + (size_t)inst * B.strideC;

    // ---- row map + the reference's up-front checks, replicated in every CTA (PrimalSimplex.cs:66-77)
    cta_row_map<THREADS>(B, bi, nex, exo, mode, rsrc, rsgn, ctl);
    __syncthreads();
    int status = ctl[0];

    int n_piv = 0, n_silent = 0, n_hist = 0;
    int* plog = (B.pivots && rank == 0) ? B.pivots + (size_t)p * B.pivots_cap * 2 : nullptr;
    double* hist = B.history ? B.history + (size_t)p * B.history_stride : nullptr;
    const size_t tsize = (size_t)rows * width;
    cluster.sync();  // every CTA of the cluster is running before anybody writes into a peer

    if (status == LPX_RUNNING) {
        // ---- BuildTableau, own rows ---------------------------------------------------------------
        for (int i = r_lo + warp; i < r_hi; i += NW) {
            double* Ti = loc(i);
            if (i < m) {
                const int src = rsrc[i];
                const bool flip = rsgn[i] != 0;
                const int xv = src < 0 ? B.ex_var[exo + (-1 - src)] : -1;
                const double* Ar = src >= 0 ? Ai + (size_t)src * n : nullptr;
                for (int j = lane; j < width; j += 32) {
                    double v;
                    if (j < n) {
                        v = Ar ? Ar[j] : (j == xv ? 1.0 : 0.0);
                        v = neg_if(v, flip);
                    } else if (j == rhs) {
                        v = src >= 0 ? bi[src] : B.ex_rhs[exo + (-1 - src)];
                        v = neg_if(v, flip);
                    } else {
                        v = (j == n + i) ? 1.0 : 0.0;
                    }
                    Ti[j] = v;
                }
            } else {
                for (int j = lane; j < width; j += 32) {
                    double v = 0.0;
                    if (j < n) {
                        double cj = ci[j];
                        if (B.sense == 1) cj = dneg(cj);
                        v = dneg(cj);
                    }
                    Ti[j] = v;
                }
            }
        }
        if (ld > width) {  // pad column and pad entry of the pivot row (never read back, kept finite)
            for (int i = r_lo + tid; i < r_hi; i += THREADS) loc(i)[width] = 0.0;
            if (tid == 0) prow[width] = 0.0;
        }
        for (int i = tid; i < m; i += THREADS) sbasis[i] = n + i;
        __syncthreads();

        // Gauss-Jordan pivot of a DUAL step on (l, e).  raw_row_everywhere: prow of every CTA already holds the raw
        // row l (dual step); otherwise its owner normalises it and sends the quotients.
        auto pivot = [&](int l, int e, bool raw_row_everywhere) {
            if (!raw_row_everywhere) {
                if (rank == owner(l)) {
                    const double* Tl = loc(l);
                    const double piv = Tl[e];
                    for (int j = tid; j < width; j += THREADS) {
                        const double pj = ddiv_by_pivot(Tl[j], piv);
#pragma unroll
                        for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(prow, rk)[j] = pj;
                    }
                }
                cluster.sync();
            } else {
                const double piv = prow[e];
                __syncthreads();
                for (int j = tid; j < width; j += THREADS) prow[j] = ddiv_by_pivot(prow[j], piv);
            }
            for (int i = r_lo + tid; i < r_hi; i += THREADS) fcol[i - r_lo] = loc(i)[e];
            __syncthreads();
            // two adjacent columns per thread, 16-byte shared accesses (rows start 16-byte aligned)
            const int pairs = (width + 1) >> 1;
            const int cw2 = (pairs + 31) & ~31;
            const int G = cw2 >= THREADS ? 1 : THREADS / cw2;
            for (int q0 = 0; q0 < pairs; q0 += THREADS) {
                const int g = cw2 >= THREADS ? 0 : tid / cw2;
                const int q = cw2 >= THREADS ? q0 + tid : tid - g * cw2;
                if (g < G && q < pairs) {
                    const double2 pj = *reinterpret_cast<const double2*>(prow + 2 * q);
                    double* t = T + (size_t)g * ld + 2 * q;
                    const size_t step = (size_t)G * ld;
#pragma unroll 4
                    for (int i = r_lo + g; i < r_hi; i += G, t += step) {
                        double2 cur = *reinterpret_cast<double2*>(t);
                        const double f = fcol[i - r_lo];
                        if (i == l) {
                            cur = pj;
                        } else {
                            cur.x = __dsub_rn(cur.x, __dmul_rn(f, pj.x));
                            cur.y = __dsub_rn(cur.y, __dmul_rn(f, pj.y));
                        }
                        *reinterpret_cast<double2*>(t) = cur;
                    }
                }
                if (cw2 < THREADS) break;
            }
            __syncthreads();
        };
        auto log_pivot = [&](int e, int l) {
            if (tid == 0) {
                sbasis[l] = e;
                if (plog && n_piv < B.pivots_cap) {
                    plog[2 * n_piv] = e;
                    plog[2 * n_piv + 1] = l;
                }
            }
        };
        auto snapshot = [&]() {  // own rows of the iteration tableau (AppendTableau's numbers)
            if (hist && n_hist < B.history_cap) {
                double* dst = hist + (size_t)n_hist * tsize;
                for (int i = r_lo; i < r_hi; i++)
                    for (int j = tid; j < width; j += THREADS) dst[(size_t)i * width + j] = loc(i)[j];
                n_hist++;
            }
        };

        // ---- primal pivots: two cluster barriers and two block barriers each -------------------------
        // Warp 0 of EVERY CTA is a control warp with its own copy of the whole objective row (zc): all
        // copies see the same operations in the same order as the tableau's row, so they stay bit-identical
        // to it and to one another, and each CTA picks the next entering column by itself while its other
        // warps update their rows — no exchange, no barrier for ChooseEntering.  What still crosses the
        // cluster: the ratios of every CTA's rows (all -> all; every control warp repeats the exact scan)
        // and the normalised pivot row (owner -> all).
        // Returns why it stopped: 0 = `limit` pivots done, 1 = no entering column, 2 = no leaving row.
        auto primal_steps = [&](double margin, int limit, bool silent) -> int {
            // the objective row as it stands: its owner sends it to every CTA's copy
            if (rank == zr) {
                const double* zrow = loc(m);
                for (int j = tid; j < width; j += THREADS) {
                    const double v = zrow[j];
#pragma unroll
                    for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(zc, rk)[j] = v;
                }
            }
            cluster.sync();
            if (warp == 0) {
                const int e0 = warp_argmin_below(zc, width - 1, -LPX_EPS);
                if (lane == 0) ctl[3] = e0;
            }
            __syncthreads();
            int steps = 0;
            while (true) {
                if (steps >= limit) return 0;
                const int e = ctl[3];
                if (e < 0) return 1;
                // ratios and factors of the own rows (the tableau is current: the update ended with a barrier)
                for (int i = r_lo + tid; i < r_hi; i += THREADS) {
                    const double a = loc(i)[e];
                    fcol[i - r_lo] = a;
                    if (i < m) {
                        double r = __longlong_as_double(0x7ff8000000000000LL);
                        if (a > LPX_EPS) r = ddiv_by_pos(loc(i)[rhs], a);
#pragma unroll
                        for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(ratio, rk)[i] = r;
                    }
                }
                cluster.sync();
                if (warp == 0) {
                    const int lv = warp_margin_scan_staged(m, margin, ratio);
                    if (lane == 0) ctl[2] = lv;
                }
                __syncthreads();
                const int l = ctl[2];
                if (l < 0) return 2;
                if (rank == owner(l)) {  // the owner normalises the pivot row and sends the quotients
                    const double* Tl = loc(l);
                    const double piv = Tl[e];
                    for (int j = tid; j < width; j += THREADS) {
                        const double pj = ddiv_by_pivot(Tl[j], piv);
#pragma unroll
                        for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(prow, rk)[j] = pj;
                    }
                }
                cluster.sync();
                if (warp == 0) {
                    const double fz = zc[e];  // the objective row's factor: its entry in the entering column
                    __syncwarp();
                    for (int j = lane; j < width; j += 32) zc[j] = __dsub_rn(zc[j], __dmul_rn(fz, prow[j]));
                    __syncwarp();
                    const int en = warp_argmin_below(zc, width - 1, -LPX_EPS);
                    if (lane == 0) {
                        ctl[3] = en;
                        sbasis[l] = e;
                        if (plog && n_piv < B.pivots_cap) {
                            plog[2 * n_piv] = e;
                            plog[2 * n_piv + 1] = l;
                        }
                    }
                } else {
                    // own rows, two adjacent columns per thread, 16-byte shared accesses
                    const int utid = tid - 32, UT = THREADS - 32;
                    const int pairs = (width + 1) >> 1;
                    const int cw2 = (pairs + 31) & ~31;
                    const int G = cw2 >= UT ? 1 : UT / cw2;
                    for (int q0 = 0; q0 < pairs; q0 += UT) {
                        const int g = cw2 >= UT ? 0 : utid / cw2;
                        const int q = cw2 >= UT ? q0 + utid : utid - g * cw2;
                        if (g < G && q < pairs) {
                            const double2 pj = *reinterpret_cast<const double2*>(prow + 2 * q);
                            double* t = T + (size_t)g * ld + 2 * q;
                            const size_t step = (size_t)G * ld;
#pragma unroll 4
                            for (int i = r_lo + g; i < r_hi; i += G, t += step) {
                                double2 cur = *reinterpret_cast<double2*>(t);
                                const double f = fcol[i - r_lo];
                                if (i == l) {
                                    cur = pj;
                                } else {
                                    cur.x = __dsub_rn(cur.x, __dmul_rn(f, pj.x));
                                    cur.y = __dsub_rn(cur.y, __dmul_rn(f, pj.y));
                                }
                                *reinterpret_cast<double2*>(t) = cur;
                            }
                        }
                        if (cw2 < UT) break;
                    }
                }
                __syncthreads();
                n_piv++;
                if (silent) n_silent++;
                else snapshot();
                steps++;
            }
        };

        if (mode == 1) primal_steps(LPX_MARGIN_DUAL, 100, true);  // ForceDualFeasibility: <= 100 silent pivots
        snapshot();

        if (mode == 0) {
            const int why = primal_steps(LPX_MARGIN_PRIMAL, B.max_iter, false);
            status = why == 0 ? LPX_S_ITER_LIMIT : (why == 1 ? LPX_OPTIMAL : LPX_UNBOUNDED);
        }
        int iter = 1;
        while (mode == 1) {
            if (iter > LPX_DUAL_MAX_ITER) {
                status = LPX_S_ITER_LIMIT;
                break;
            }
            int e, l;
            {
                // dual: leaving row = most negative RHS below -1e-9, lowest row on ties (DualSimplex.cs:45-55)
                {
                    const int cnt_own = max(0, min(r_hi, m) - r_lo);
                    unsigned long long kl = ~0ULL;
                    int il = INT_MAX;
                    for (int k = tid; k < cnt_own; k += THREADS) {
                        const double v = loc(r_lo + k)[rhs];
                        if (v < -LPX_EPS) {
                            const unsigned long long kk = dkey(v);
                            if (kk < kl) {
                                kl = kk;
                                il = r_lo + k;
                            }
                        }
                    }
                    const unsigned long long K = warp_min_u64(kl);
                    const int iw = __reduce_min_sync(0xffffffffu, kl == K ? il : INT_MAX);
                    if (lane == 0) {
                        red[warp].v = __longlong_as_double((long long)K);
                        red[warp].i = iw;
                    }
                    __syncthreads();
                    if (warp == 0) {
                        unsigned long long k2 = ~0ULL;
                        int i2 = INT_MAX;
                        if (lane < NW) {
                            k2 = (unsigned long long)__double_as_longlong(red[lane].v);
                            i2 = red[lane].i;
                        }
                        const unsigned long long K2 = warp_min_u64(k2);
                        const int idx = __reduce_min_sync(0xffffffffu, k2 == K2 ? i2 : INT_MAX);
                        if (lane < CL) {
                            ArgMin a;
                            a.v = __longlong_as_double((long long)K2);
                            a.i = idx;
                            cluster.map_shared_rank(part, lane)[rank] = a;
                        }
                    }
                    cluster.sync();
                    unsigned long long best = ~0ULL;
                    l = -1;
#pragma unroll
                    for (int rk = 0; rk < CL; rk++) {  // ranks hold ascending row ranges: first minimum wins
                        const unsigned long long kk = (unsigned long long)__double_as_longlong(part[rk].v);
                        if (kk < best) {
                            best = kk;
                            l = part[rk].i;
                        }
                    }
                    if (best == ~0ULL) l = -1;
                }
                if (l < 0) {
                    status = LPX_OPTIMAL;
                    break;
                }
                // raw row l -> every CTA
                if (rank == owner(l)) {
                    const double* Tl = loc(l);
                    for (int j = tid; j < width; j += THREADS) {
                        const double v = Tl[j];
#pragma unroll
                        for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(prow, rk)[j] = v;
                    }
                }
                cluster.sync();
                // entering column: min z_j / (-a) over a < -1e-9, margin 1e-12 (DualSimplex.cs:76-91)
                if (rank == zr) {
                    const double* zrow = loc(m);
                    for (int j = tid; j < width - 1; j += THREADS) {
                        const double a = prow[j];
                        double r = __longlong_as_double(0x7ff8000000000000LL);
                        if (a < -LPX_EPS) r = ddiv_by_pivot(zrow[j], dneg(a));
#pragma unroll
                        for (int rk = 0; rk < CL; rk++) cluster.map_shared_rank(ratio, rk)[j] = r;
                    }
                }
                cluster.sync();
                if (warp == 0) {
                    const int ev = warp_margin_scan_cert(width - 1, LPX_MARGIN_DUAL, [&](int j, double& r) {
                        r = ratio[j];
                        return r == r;
                    });
                    if (lane == 0) ctl[2] = ev;
                }
                __syncthreads();
                e = ctl[2];
                if (e < 0) {
                    status = LPX_INFEASIBLE;
                    break;
                }
                pivot(l, e, true);
            }
            log_pivot(e, l);
            n_piv++;
            snapshot();
            iter++;
        }
        __syncthreads();

        // ---- FinalizeReport's numeric part (PrimalSimplex.cs:132-138) -----------------------------
        if (B.basis && rank == 0)
            for (int i = tid; i < m; i += THREADS) B.basis[(size_t)p * B.basis_stride + i] = sbasis[i];
        if (B.x) {
            double* xo = B.x + (size_t)p * n;
            if (rank == 0)
                for (int j = tid; j < n; j += THREADS) xo[j] = 0.0;
            __threadfence();
            cluster.sync();
            for (int i = r_lo + tid; i < r_hi && i < m; i += THREADS)
                if (sbasis[i] < n) xo[sbasis[i]] = loc(i)[rhs];
        }
        if (B.z && rank == zr && tid == 0) B.z[p] = loc(m)[rhs];
        if (B.node_flags && B.x && mode == 0) {
            cluster.sync();  // every CTA's part of x is written and visible
            if (rank == 0) cta_node_epilogue<THREADS>(B, p, inst, nex, exo, B.x + (size_t)p * n, prow, red);
        }
        if (B.tableau) {
            double* dst = B.tableau + (size_t)p * B.tableau_stride;
            for (int i = r_lo; i < r_hi; i++)
                for (int j = tid; j < width; j += THREADS) dst[(size_t)i * width + j] = loc(i)[j];
        }
    } else {
        cta_zero_outputs<THREADS>(B, p, m, n, rows, width, rank, CL);
    }

    if (rank == 0 && tid == 0) {
        B.status[p] = status;
        if (B.n_pivots) B.n_pivots[p] = n_piv;
        if (B.silent) B.silent[p] = n_silent;
        if (B.n_history) B.n_history[p] = n_hist;
        if (B.total_pivots && n_piv) atomicAdd(B.total_pivots, (unsigned long long)n_piv);
    }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

bool cta_cluster_fits(int max_rows, int max_width, int cl);

}  // namespace lpx
