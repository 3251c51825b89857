// lpx_stream_pipe.cuh — software-pipelined blocked pivoting: the look-ahead of block B+1 runs
// CONCURRENTLY with the HBM pass of block B.
//
// The pass is out of place (two tableau buffers, ping-pong): while pass(B) streams T[in] -> T[out],
// the look-ahead of block B+1 reads the columns and rows it needs from T[in], which nobody writes,
// and brings them up to date by applying block B's rank-1 terms first and its own afterwards — the
// same operations, in the same pivot order, that the passes apply to the whole tableau.  Blocks
// hold at most LPX_PIPE_K = 8 pivots, so the combined list fits the 16-entry machinery of
// lpx_stream_block.cuh.  Factor columns, pivot rows and leaving rows live in two halves of
// Fbuf / Pbuf / Lbuf (rows par*8 ..), the z-row and the RHS travel between look-aheads in compact
// vectors.  Per block the GPU time is max(pass, look-ahead) instead of their sum.
//
// Included by lpx_stream.cu after lpx_stream_block.cuh.
#pragma once

namespace lpx {

#define LPX_PIPE_K 8
// row of Fbuf / Pbuf / Lbuf holding entry s of the combined list (pc pending of the previous block, then own)
#define PIPE_ROW(s_) (((s_) < pc) ? (q * LPX_PIPE_K + (s_)) : (par * LPX_PIPE_K + (s_) - pc))

__global__ void __cluster_dims__(LPX_LA_CLUSTER, 1, 1) __launch_bounds__(LPX_LA_THREADS)
    stream_lookahead_pipe_kernel(StreamParams P, int budget, int par, int overlap) {
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
    // previous block (parity q): its pass may still be running, so its updates are applied here
    const int q = par ^ 1;
    const int prev_in = ctl->blk_in[q], prev_cnt = ctl->blk_cnt[q];
    const int in_buf = overlap ? prev_in : (prev_in ^ (prev_cnt > 0 ? 1 : 0));  // buffer this kernel reads
    const int pc = overlap ? prev_cnt : 0;                                      // pending updates on top of it
    const double* Tin = in_buf ? P.T1 : P.T;
    cluster.sync();  // every CTA has read the control block before rank 0 rewrites it
    const bool probe = budget <= 0;
    if (status0 != LPX_RUNNING) {
        if (!probe && rank == 0 && tid == 0) {
            ctl->blk_cnt[par] = 0;
            ctl->blk_in[par] = prev_in ^ (prev_cnt > 0 ? 1 : 0);
        }
        return;
    }
    for (int j = j_lo + tid; j < j_hi; j += TH) zloc[j - j_lo] = P.zbuf[j];
    for (int i = tid; i < rows; i += TH) rhs[i] = P.rhsbuf[i];
    if (tid < pc) s_L[tid] = P.Lbuf[q * LPX_PIPE_K + tid];
    __syncthreads();

    // ChooseEntering, first half: this CTA's z-slice argmin (REDUX minima on the value key, then on the
    // index), pushed into every CTA's s_part through distributed shared memory.  The cluster barrier
    // that follows — (3) of the previous pivot, or the one before the loop — publishes it, so a pivot
    // costs two cluster barriers plus this one, not three plus one.
    auto push_entering_candidates = [&]() {
        unsigned long long kl = ~0ULL;
        int il = INT_MAX;
        for (int j = j_lo + tid; j < j_hi && j < width - 1; j += TH) {
            const double zv = zloc[j - j_lo];
            if (zv < -LPX_EPS) {
                const unsigned long long kk = dkey(zv);
                if (kk < kl) {
                    kl = kk;
                    il = j;
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
            if (lane < TH / 32) {
                k2 = (unsigned long long)__double_as_longlong(red[lane].v);
                i2 = red[lane].i;
            }
            const unsigned long long K2 = warp_min_u64(k2);
            const int idx = __reduce_min_sync(0xffffffffu, k2 == K2 ? i2 : INT_MAX);
            if (lane < CL) {
                ArgMin b2;
                b2.v = K2 == ~0ULL ? -LPX_EPS : dkey_inv(K2);
                b2.i = idx;
                ArgMin* dst = cluster.map_shared_rank(s_part, lane);
                dst[rank] = b2;
            }
        }
    };

    const int steps = probe ? 1 : min(min(budget, P.kblock), LPX_PIPE_K);
    int cnt = 0, st = LPX_RUNNING;
    push_entering_candidates();
    cluster.sync();
    for (int k = 0; k < steps; k++) {
        if (done + cnt >= P.max_iter) {
            st = LPX_S_ITER_LIMIT;
            break;
        }
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[0] = lpx_gtime();
        // ---- ChooseEntering, second half: combine the CL slice candidates -------------------------
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
        // One global round trip for the whole phase: the p_s[e] scalars, the gathered column entries
        // and the factor entries are all requested before anything waits (the loads inside a
        // dependent chain cost ~1 us each under the memory load of the concurrent pass).
        if (tid < pc + cnt) s_pe[tid] = P.Pbuf[(size_t)PIPE_ROW(tid) * ld + e];
        for (int base = i_lo; base < i_hi; base += 2 * TH) {
            double c[2];
            double f[2][LPX_BLOCK_KMAX];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = base + u * TH + tid;
                c[u] = i < i_hi ? Tin[(size_t)i * ld + e] : 0.0;
            }
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int i = base + u * TH + tid;
                    f[u][s] = (s < pc + cnt && i < i_hi) ? P.Fbuf[(size_t)PIPE_ROW(s) * cs + i] : 0.0;
                }
            }
            __syncthreads();  // s_pe is in shared memory (first trip only matters; later trips are rare)
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
                if (s < pc + cnt) {
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
                    // the ratio of the row travels with the column entry (one division per thread here
                    // instead of eight per thread after the all-gather)
                    double r = __longlong_as_double(0x7ff8000000000000LL);
                    if (i < m && c[u] > LPX_EPS) r = ddiv_by_pos(rhs[i], c[u]);
                    const int slot = (i % Q) * TH + i / Q;  // permuted position of row i
#pragma unroll
                    for (int rk = 0; rk < CL; rk++) {
                        cluster.map_shared_rank(col, rk)[i] = c[u];
                        if (i < m) cluster.map_shared_rank(ratio, rk)[slot] = r;
                    }
                }
            }
        }
        cluster.sync();  // (2) every CTA holds the whole up-to-date column
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[2] = lpx_gtime();
        const int l = la_leaving_scan(ratio, m, Q, s_wmin, s_wcnt, s_rec, &s_out, s_ired);
        if (P.dbg && rank == 0 && tid == 0 && cnt == 2) P.dbg[3] = lpx_gtime();
        if (l < 0) {
            st = LPX_UNBOUNDED;
            break;
        }
        if (probe) break;
        const double piv = col[l], fz = col[m];
        // ---- row l, columns of this CTA: gather, bring up to date, normalise, advance z ----------
        if (tid < pc + cnt) s_fl[tid] = P.Fbuf[(size_t)PIPE_ROW(tid) * cs + l];
        double* pout = P.Pbuf + (size_t)(par * LPX_PIPE_K + cnt) * ld;
        const double* Tl = Tin + (size_t)l * ld;
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
                    pp[u][s] = (s < pc + cnt && j < j_hi) ? P.Pbuf[(size_t)PIPE_ROW(s) * ld + j] : 0.0;
                }
            }
            __syncthreads();  // s_fl is in shared memory
#pragma unroll
            for (int s = 0; s < LPX_BLOCK_KMAX; s++) {
                if (s < pc + cnt) {
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
        double* fout = P.Fbuf + (size_t)(par * LPX_PIPE_K + cnt) * cs;
        for (int i = tid; i < rows; i += TH) {
            const double f = col[i];
            if (i >= i_lo && i < i_hi) fout[i] = f;
            rhs[i] = (i == l) ? prhs : __dsub_rn(rhs[i], __dmul_rn(f, prhs));
        }
        if (tid == 0) {
            s_L[pc + cnt] = l;
            if (rank == 0) {
                P.Lbuf[par * LPX_PIPE_K + cnt] = l;
                P.basis[l] = e;
                if (done + cnt < P.pivlog_cap) {
                    P.pivlog[2 * (done + cnt)] = e;
                    P.pivlog[2 * (done + cnt) + 1] = l;
                }
            }
        }
        cnt++;
        __syncthreads();              // zloc is complete
        push_entering_candidates();   // for the NEXT pivot, published by the barrier below
        __threadfence();
        cluster.sync();  // (3) Pbuf / Fbuf slices of this pivot are visible to the whole cluster
        if (P.dbg && rank == 0 && tid == 0 && cnt == 3) P.dbg[5] = lpx_gtime();
    }
    if (!probe) {
        for (int j = j_lo + tid; j < j_hi; j += TH) P.zbuf[j] = zloc[j - j_lo];
        if (rank == 0) {
            for (int i = tid; i < rows; i += TH) P.rhsbuf[i] = rhs[i];
            if (tid == 0) {
                ctl->blk_cnt[par] = cnt;
                ctl->blk_in[par] = prev_in ^ (prev_cnt > 0 ? 1 : 0);  // what the pass of this block reads
                ctl->pivots = done + cnt;
            }
        }
    }
    if (rank == 0 && tid == 0 && st != LPX_RUNNING) ctl->status = st;
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}


template <int KMAX>
__global__ void __launch_bounds__(256) stream_update_pipe_tma_kernel(StreamParams P, int rpc, int par) {
    extern __shared__ __align__(128) unsigned char smem_tma[];
    __shared__ __align__(8) unsigned long long full[LPX_TMA_STAGES];
    __shared__ int sL[KMAX];
    constexpr int TR = LPX_TMA_ROWS, TC = LPX_TMA_COLS, ST = LPX_TMA_STAGES;
    typedef TmaTile<KMAX> Tile;
    Tile* tiles = reinterpret_cast<Tile*>(smem_tma);
    const int cnt = P.ctl->blk_cnt[par];
    if (cnt <= 0) return;
    const int in_buf = P.ctl->blk_in[par];
    const double* Tin = in_buf ? P.T1 : P.T;  // out of place: the look-ahead of the NEXT block reads Tin meanwhile
    double* Tout = in_buf ? P.T : P.T1;
    const int row0 = par * LPX_PIPE_K;  // this block's rows of Fbuf / Pbuf / Lbuf
    const int tid = threadIdx.x;
    const int r0 = blockIdx.y * rpc;
    const int r1 = min(P.rows, r0 + rpc);
    const int jbase = blockIdx.x * TC;
    if (r1 <= r0 || jbase >= P.ld) return;
    const int seg = min(TC, P.ld - jbase);            // columns of this strip (multiple of 16)
    const unsigned seg_bytes = (unsigned)seg * 8;
    const int n_tiles = (r1 - r0 + TR - 1) / TR;
    const size_t ld = (size_t)P.ld, cs = (size_t)P.colstride;

    if (tid < KMAX) sL[tid] = tid < cnt ? P.Lbuf[row0 + tid] : -1;
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
        for (int i = 0; i < nr; i++) bulk_g2s(&tl.t[i][0], Tin + (size_t)(rr + i) * ld + jbase, seg_bytes, &full[k % ST]);
        for (int s = 0; s < cnt; s++) bulk_g2s(&tl.f[s][0], P.Fbuf + (size_t)(row0 + s) * cs + rr, TR * 8, &full[k % ST]);
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
        p[s][0] = (s < cnt && col_ok) ? P.Pbuf[(size_t)(row0 + s) * ld + jbase + cp2] : 0.0;
        p[s][1] = (s < cnt && col_ok) ? P.Pbuf[(size_t)(row0 + s) * ld + jbase + cp2 + 1] : 0.0;
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
            for (int i = 0; i < nr; i++) bulk_s2g(Tout + (size_t)(rr + i) * ld + jbase, &tl.t[i][0], seg_bytes);
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
