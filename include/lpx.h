/* lpx — C ABI of the B200-native LP/IP solve engine (liblpx.so).
 *
 * Drop-in boundary for the solver seam of Jellyman750/Linear_Programming_Solver_LPR381
 * (R = Linear_Programming_Solver/ in that repository):
 *
 *     interface ILPAlgorithm { SimplexResult Solve(LPProblem, Action<string,bool[,]> updatePivot); }
 *                                                                    R/Models/IPLAlgorithm.cs:5-8
 *
 * The reference has no FFI of its own; a C# shim class per algorithm (see INTEGRATION.md and
 * linear_programming_solver_lpr381_b200/csharp/) flattens LPProblem into the arrays below and
 * P/Invokes these entry points.  Everything here is blittable: plain pointers, ints, doubles.
 * No torch types, no C++ types, no callbacks from device threads.
 *
 * Conventions
 *   sense : 0 = Max, 1 = Min                      (enum Sense, R/Models/PrimalSimplex.cs:8)
 *   rel   : 0 = LE, 1 = GE, 2 = EQ                (enum Rel,   R/Models/PrimalSimplex.cs:9)
 *   A     : row-major m x n (Constraints[i].A[j]), b[m], c[n]
 *   tableau: row-major rows x cols, rows = m' + 1, cols = n + m' + 1 where m' = m + #EQ rows;
 *            the z-row is the LAST row, the RHS the last column (R/Models/PrimalSimplex.cs:179-203)
 *   pivots: (entering column, leaving row) pairs, one per Pivot call
 *   All arithmetic is IEEE binary64 with separate multiply / subtract and true division, in the
 *   reference's evaluation order, so results are bit-identical to the C# loops.
 *
 * Return value of every int function: LPX_OK (0) or a negative LPX_E_* for failures of the call
 * itself (bad arguments, CUDA errors).  Solver outcomes, including the reference's exception
 * cases, are reported per problem in *status.
 *
 * There is NO CPU fallback: without a usable sm_100 device every solve entry point fails with
 * LPX_E_CUDA and lpx_last_error() says why.
 */
#ifndef LPX_H_
#define LPX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (per problem) ------------------------------------------------------------- */
#define LPX_OPTIMAL            0   /* "OPTIMAL"    R/Models/PrimalSimplex.cs:126 */
#define LPX_UNBOUNDED          1   /* "UNBOUNDED"  R/Models/PrimalSimplex.cs:102-106 */
#define LPX_INFEASIBLE         2   /* "INFEASIBLE" R/Models/DualSimplex.cs:92-96 (dual simplex only) */
#define LPX_RUNNING            3   /* session not finished yet */
#define LPX_S_GE_ROW          -1   /* exception, R/Models/PrimalSimplex.cs:68-71 */
#define LPX_S_NEG_RHS         -2   /* exception, R/Models/PrimalSimplex.cs:73-76 */
#define LPX_S_ITER_LIMIT      -3   /* exception, R/Models/PrimalSimplex.cs:95-96, DualSimplex.cs:39, RevisedPrimalSimplex.cs:144 */
#define LPX_S_REV_UNSUPPORTED -10  /* exception, R/Models/RevisedPrimalSimplex.cs:19-21 (needs all <= rows, b >= -1e-9) */
#define LPX_S_SINGULAR        -11  /* exception, R/Models/RevisedPrimalSimplex.cs:433 "Singular basis encountered." */

/* ---- call-level error codes ----------------------------------------------------------------- */
#define LPX_OK                 0
#define LPX_E_BAD_ARGS        -4
#define LPX_E_CUDA            -5
#define LPX_E_CAPACITY        -8   /* problem too large for the selected kernel / buffer */
#define LPX_E_NCCL            -9

/* ---- options; zero-initialise then lpx_default_options() ------------------------------------ */
typedef struct lpx_options {
    int max_iterations;   /* PrimalSimplex.MaxIterations = 10000 (R/Models/PrimalSimplex.cs:54) */
    int kernel;           /* LPX_KERNEL_*: force a kernel family (tests/benchmarks); 0 = auto */
    int threads;          /* CTA size override for the per-tableau kernels; 0 = auto */
    int knap_spec_nodes;  /* knapsack B&B: heap nodes speculated per round (0 = default 16) */
    int knap_spec_depth;  /* knapsack B&B: look-ahead depth under the front runner (0 = default 2) */
    int stream_protocol;  /* streaming kernels: 0 auto (pipelined blocked look-ahead: the look-ahead of block
                             B+1 overlaps the HBM pass of block B), 1 single-CTA select per pivot, 2 multi-CTA
                             prep per pivot, 3 blocked with a one-CTA look-ahead, 4 blocked, not pipelined */
    int reg_variant;      /* register-resident kernel: 0 = default build (4 when n <= 128, else 2), 1 one CTA per SM
                             (all registers), 2 two CTAs per SM (two column slots in shared memory), 3 three CTAs per
                             SM (four slots), 4 condensed tableau (non-basic columns only), three CTAs per SM */
    int stream_block;     /* streaming kernels: pivots applied per HBM pass (0 = default, max 16) */
    int stream_pass_variant; /* blocked pass: 0 auto, 1 two doubles per thread, 2 one double per thread,
                                3 TMA-staged (cp.async.bulk + mbarrier pipeline) */
    int knap_ordered_sums; /* knapsack: 1 = always sum in the reference's order, even for exactly summable
                              integer data (which otherwise take the warp-parallel exact path) */
    int knap_shard_tree;   /* knapsack: 1 = ONE search tree evaluated by all ranks of lpx_comm_init: every rank
                              plans and commits identically, evaluates the speculative subtrees it owns
                              (round robin), and the relaxations are merged by an NCCL all-reduce per round */
    int knap_warps;        /* knapsack, warps per instance: 0 auto (2 = a main + helper pair when there are at most four
                              instances per SM, else 1), 1 or 2 to force */
    int reserved[4];
} lpx_options;

#define LPX_KERNEL_AUTO        0
#define LPX_KERNEL_CTA_SMEM    1   /* one CTA per tableau, tableau resident in shared memory */
#define LPX_KERNEL_CTA_GLOBAL  2   /* one CTA per tableau, tableau in global memory (L2/HBM) */
#define LPX_KERNEL_STREAM      3   /* one tableau over the whole GPU, HBM-streamed rank-1 pivots */
#define LPX_KERNEL_CTA_REG     4   /* one CTA per tableau, tableau resident in registers */
#define LPX_KERNEL_CTA_CLUSTER 5   /* one 2- or 4-CTA cluster per tableau, rows split over their shared memories */

void lpx_default_options(lpx_options* opt);

/* ---- library / device ------------------------------------------------------------------------ */
const char* lpx_version(void);
const char* lpx_last_error(void);              /* thread-local, never NULL */
const char* lpx_status_message(int status);    /* the reference's exception text for status < 0 */
int lpx_device_count(void);
int lpx_init(int device);                      /* device < 0: keep the current CUDA device */
void lpx_shutdown(void);
void* lpx_host_alloc(size_t bytes);            /* page-locked host memory for fast H2D/D2H */
void lpx_host_free(void* p);
/* rows/cols of the tableau after EQ expansion (R/Models/PrimalSimplex.cs:161-177). */
int lpx_tableau_dims(int m, int n, const int* rel, int* rows, int* cols);

/* ---- Primal Simplex: PrimalSimplex.Solve, R/Models/PrimalSimplex.cs:57-127 ------------------ */
/* One LP, host buffers.  Any output pointer may be NULL.  history (nullable) receives the
 * tableau after iteration 0..k, at most history_cap tableaux: what AppendTableau prints each
 * iteration (R/Models/PrimalSimplex.cs:88-90,113-114). */
int lpx_primal_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                     const lpx_options* opt, int* status, int* n_pivots, int* pivots, int pivots_cap, int* basis,
                     double* x, double* z, double* tableau, double* history, int history_cap);

/* `count` LPs of one shape (shared m, n, sense, rel pattern), host buffers, strided per problem:
 * A[count][m][n], b[count][m], c[count][n]; outputs status[count], n_pivots[count],
 * basis[count][m'], x[count][n], z[count], tableau[count][rows][cols] (nullable).
 * Copies are staged and overlapped with the solve; pass lpx_host_alloc memory for full speed. */
int lpx_primal_solve_batched(int count, int m, int n, int sense, const double* A, const int* rel, const double* b,
                             const double* c, const lpx_options* opt, int* status, int* n_pivots, int* basis,
                             double* x, double* z, double* tableau, long long* total_pivots);

/* Same, all pointers are DEVICE pointers (rel too), asynchronous on `stream` (a cudaStream_t;
 * NULL = default stream).  total_pivots is a device pointer to one 64-bit counter (nullable),
 * accumulated with atomicAdd — zero it first. */
int lpx_primal_solve_batched_dev(int count, int m, int n, int sense, const double* A, const int* rel,
                                 const double* b, const double* c, const lpx_options* opt, int* status,
                                 int* n_pivots, int* basis, double* x, double* z, double* tableau,
                                 unsigned long long* total_pivots, void* stream);

/* ---- Dual Simplex: DualSimplex.Solve, R/Models/DualSimplex.cs:15-114 ------------------------ */
/* pivots lists the *silent ForceDualFeasibility pivots first (R/Models/DualSimplex.cs:195-228);
 * history[0] is the tableau after them ("Iteration 0"). */
int lpx_dual_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                   const lpx_options* opt, int* status, int* n_pivots, int* silent_pivots, int* pivots,
                   int pivots_cap, int* basis, double* x, double* z, double* tableau, double* history,
                   int history_cap);

/* ---- Large single LP, stepwise (HBM-streamed pivots) ---------------------------------------- */
typedef struct lpx_session lpx_session;
/* Host buffers are copied to the device once; the tableau then lives in HBM for the session. */
lpx_session* lpx_session_open(int m, int n, int sense, const double* A, const int* rel, const double* b,
                              const double* c, const lpx_options* opt);
/* A, b, c are DEVICE pointers (rel stays a host pointer: it only shapes the tableau). */
lpx_session* lpx_session_open_dev(int m, int n, int sense, const double* A, const int* rel, const double* b,
                                  const double* c, const lpx_options* opt);
/* Run at most max_pivots more pivots (stops early at OPTIMAL / UNBOUNDED / ITER_LIMIT).
 * *status = LPX_RUNNING while unfinished.  *pivots_total = pivots performed since open. */
int lpx_session_step(lpx_session* s, int max_pivots, int* status, int* pivots_total);
/* Enqueue max_pivots pivots without waiting (for timing with caller-side CUDA events). */
int lpx_session_step_async(lpx_session* s, int max_pivots);
int lpx_session_sync(lpx_session* s, int* status, int* pivots_total);
void* lpx_session_stream(lpx_session* s);            /* the cudaStream_t the session launches on */
int lpx_session_dims(const lpx_session* s, int* rows, int* cols);
int lpx_session_read_tableau(lpx_session* s, double* tableau);          /* rows x cols, compact */
int lpx_session_read_solution(lpx_session* s, int* basis, double* x, double* z);
int lpx_session_read_pivots(lpx_session* s, int* pivots, int pivots_cap);
void lpx_session_close(lpx_session* s);

/* ---- Branch & Bound (simplex): BranchAndBound.Solve, R/Models/Branch&Bound.cs:30-123 -------- */
/* One record per LP relaxation the reference solves (root LP, then every SolveNode call), in the
 * reference's order, delivered from the host thread after the node is committed. */
typedef struct lpx_bnb_node {
    int index;            /* 0 = root LP (R/Models/Branch&Bound.cs:57), 1.. = SolveNode calls */
    int depth;
    int parent;           /* index of the parent record, -1 for the root */
    int is_ceil_child;    /* 1: "x_k >= ceil" child, 0: "x_k <= floor" child / root */
    int id_path_len;      /* hierarchical id, e.g. {2,4} = "Subproblem 2.4" */
    const int* id_path;
    int bound_var;        /* variable of the branching row that created this node, -1 for root */
    int bound_val;        /* its floor / ceil value */
    int algo;             /* 0 Primal Simplex, 1 Dual Simplex (ChooseAlgorithm, :262-266) */
    int lp_status;        /* LPX_OPTIMAL.. / LPX_S_* */
    int outcome;          /* LPX_BNB_* */
    int n_pivots, silent_pivots;
    const int* pivots;    /* n_pivots pairs */
    int rows, cols;       /* node tableau shape */
    double z;
    const double* x;      /* n values (primal nodes) */
    int branch_var, floor_val, ceil_val;   /* valid when outcome == LPX_BNB_BRANCHED */
    int n_history;        /* tableaux available in history (0 unless requested) */
    const double* history;
} lpx_bnb_node;
typedef void (*lpx_bnb_node_fn)(const lpx_bnb_node* node, void* user);

#define LPX_BNB_ERROR       0   /* LP threw (R/Models/Branch&Bound.cs:150-154) */
#define LPX_BNB_INVALID     1   /* "Invalid Simplex result": every Dual Simplex node (:157-161) */
#define LPX_BNB_INFEASIBLE  2   /* IsFeasible failed (:175-179) */
#define LPX_BNB_PRUNED      3   /* z <= best + 1e-6 (:182-186) */
#define LPX_BNB_INCUMBENT   4   /* integral: new incumbent (:189-195) */
#define LPX_BNB_BRANCHED    5
#define LPX_BNB_NOFRAC      6   /* (:215-219) */
#define LPX_BNB_DEPTH       7   /* depth > 200 (:132-136) */

#define LPX_BNB_WANT_HISTORY 1  /* flags: deliver per-iteration tableaux to the node callback */

/* `count` independent IPs of one shape; instance k's records are delivered with user data
 * `user` and node->index local to the instance (on_node receives instance via lpx_bnb_instance()).
 * Outputs: found[count], best_z[count], best_x[count][n], n_nodes[count], n_lp_pivots[count]. */
int lpx_bnb_simplex_batched(int count, int m, int n, int sense, const double* A, const int* rel, const double* b,
                            const double* c, const lpx_options* opt, int flags, int* found, double* best_z,
                            double* best_x, int* n_nodes, long long* n_lp_pivots, int* root_status,
                            lpx_bnb_node_fn on_node, void* user);
int lpx_bnb_simplex(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                    const lpx_options* opt, int flags, int* found, double* best_z, double* best_x, int* n_nodes,
                    long long* n_lp_pivots, int* root_status, lpx_bnb_node_fn on_node, void* user);
int lpx_bnb_instance(void);  /* inside on_node: which instance of the batch the record belongs to */

/* ---- Mode B: the pooled tree — NOT the reference's tree ------------------------------------------------ */
/* The reference's BranchAndBound never explores a '>=' child (R/Models/Branch&Bound.cs:157-161 rejects every
 * Dual Simplex result), so the tree lpx_bnb_simplex reproduces is one floor path.  lpx_bnb_pooled is the tree
 * the class's doc comment describes (R/Models/Branch&Bound.cs:9-19) with both children honoured: children are
 * warm-started from the parent's final tableau (bound row + Dual Simplex pivots with the reference's rules),
 * each round evaluates the `batch` open nodes with the largest bound, and after lpx_comm_init the nodes of a
 * round are dealt over the ranks (every rank makes the same call and returns the same result; children read
 * their parent's tableau from the owning GPU's memory over NVLink).  All rows must be '<=' with b >= 0.
 * Per evaluated node, in commit order (node 0 = root): id, outcome (LPX_BNB_*), dual pivots, z.
 * Checked against oracle/orc_pooled.cpp node for node and against an independent MILP solver. */
int lpx_bnb_pooled(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                   const lpx_options* opt, int batch, int* found, double* best_z, double* best_x,
                   long long* n_nodes, long long* n_pivots, long long* n_rounds, long long node_cap, int* node_id,
                   int* node_outcome, int* node_pivots, double* node_z);

/* ---- Branch & Bound Knapsack: BranchAndBoundKnapsack.Solve, R/Models/BranchAndBoundKnapsack.cs:58-407 */
typedef struct lpx_knap_eval {   /* one ComputeRelaxation of the root or of a child (:108,209,269) */
    int pop_index;        /* index of the expanded pop this evaluation belongs to, -1 = root */
    int child;            /* 0: x_k = 0 (left), 1: x_k = 1 (right) */
    int var;              /* original index of the fixed item */
    double bound, weight; /* relaxation profit (== bound) and weight */
    int frac_rank;        /* ratio rank of the fractional item, -1 if none */
    double frac;          /* its fraction */
    int break_rank;       /* first ratio rank not (fully) packed; n if all were */
    int decision;         /* LPX_KN_* */
    const signed char* assigned;  /* n entries: -1 undecided, 0, 1 */
} lpx_knap_eval;
typedef struct lpx_knap_pop {    /* one node taken from the heap and expanded or closed (:120-177) */
    int pop_index;
    int label_len;        /* hierarchical label, e.g. {1,2,1} = "1.2.1"; root = {0} */
    const int* label;
    lpx_knap_eval relax;  /* the recomputed relaxation of the popped node (:127) */
    int closed;           /* 0 expanded; 1 candidate/BEST CANDIDATE; 2 CANDIDATE (not better); 3 INFEASIBLE */
} lpx_knap_pop;
typedef void (*lpx_knap_pop_fn)(const lpx_knap_pop* pop, const lpx_knap_eval* left, const lpx_knap_eval* right,
                                void* user);

#define LPX_KN_ROOT          0
#define LPX_KN_INFEASIBLE    1
#define LPX_KN_CANDIDATE_INT 2   /* integral and feasible: incumbent candidate, not pushed */
#define LPX_KN_PUSHED        3
#define LPX_KN_DROPPED       4   /* bound <= best + 1e-9 */

/* rank_order (nullable, n ints) receives the ratio ordering (original index per rank, :75-79). */
int lpx_bnb_knapsack(int n, const double* profit, const double* weight, double capacity, const lpx_options* opt,
                     int* found, double* best_value, int* best_x, long long* n_evals, long long* n_pops,
                     int* rank_order, lpx_knap_pop_fn on_pop, void* user);
/* `count` independent instances, n items each: profit[count][n], weight[count][n], capacity[count]. */
int lpx_bnb_knapsack_batched(int count, int n, const double* profit, const double* weight, const double* capacity,
                             const lpx_options* opt, int* found, double* best_value, int* best_x,
                             long long* n_evals, long long* n_pops);

/* ---- multi-GPU incumbent sharing (one process per GPU) -------------------------------------- */
/* NCCL is loaded at run time (libnccl.so.2).  Rank 0 creates the id, the host application ships
 * the 128 bytes to every rank (any transport), every rank calls lpx_comm_init. */
int lpx_comm_unique_id(void* id128);
int lpx_comm_init(int world, int rank, const void* id128);
int lpx_comm_allreduce_max(double* values, int count);      /* host values, in place */
int lpx_comm_allgather(const void* send, void* recv, size_t bytes_per_rank);  /* host buffers */
void lpx_comm_destroy(void);
int lpx_comm_world(void);
int lpx_comm_rank(void);

/* ---- Revised Primal Simplex ------------------------------------------------------------------- */
/* RevisedPrimalSimplex.Solve (R/Models/RevisedPrimalSimplex.cs:17-145): price-out with the basis
 * inverse recomputed by Gauss-Jordan (partial pivoting) every iteration, left-to-right dot products,
 * ratio margin 1e-12.  status: LPX_OPTIMAL / LPX_UNBOUNDED / LPX_S_ITER_LIMIT / LPX_S_REV_UNSUPPORTED /
 * LPX_S_SINGULAR.  pivots: 2 ints per iteration (entering column, leaving ROW = basis position),
 * theta: the chosen ratio.  basis (m) = Bidx, nonbasic (n) = Nidx in the reference's list order,
 * xB (m), Binv (m x m, row-major), x (n) = basic decision variables mapped back.
 * history (nullable): history_cap records of lpx_revised_history_stride(m, n) doubles, one per
 * BuildIterationBlock call (:62, :134): [Binv m*m][x_B m][z][r_N n][d m][theta][Bidx m][Nidx n][entering],
 * integers stored as doubles; record 0 is the initial basis (r_N, d, theta unused, entering -1). */
int lpx_revised_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                      const lpx_options* opt, int* status, int* n_iters, int* pivots, double* theta, int pivots_cap,
                      int* basis, int* nonbasic, double* xB, double* Binv, double* x, double* history,
                      int history_cap);
size_t lpx_revised_history_stride(int m, int n);

/* ---- measurement helpers (bench.py) --------------------------------------------------------- */
/* Unfused FP64 rate (separate DMUL and DADD, the only arithmetic the bit-exactness contract
 * allows) in TFLOP/s: the roofline denominator of the on-chip batched kernels. */
int lpx_measure_fp64_rate(double* tflops);
/* Dependent-chain latencies in SM cycles, one warp alone: cycles8[0..7] = DADD, DMUL, DDIV, REDUX.MIN,
 * SHFL, shared-memory store+load round trip, BAR.SYNC of 13 warps, DSETP+select.  These, not the
 * FP64 rate, bound one pivot of a per-tableau kernel. */
int lpx_measure_latencies(double* cycles8);
/* Session kernels timed separately with CUDA events on the session stream, run back to back
 * (no overlap): us[0] = mean look-ahead / select time, us[1] = mean HBM pass time, us[2] = mean time
 * per block (blocked protocols) or per pivot (per-pivot protocols), over n blocks / pivots. */
int lpx_session_profile(lpx_session* s, int n, double* us);
/* Development aid: phase timestamps (ns, %globaltimer) of one look-ahead step. */
int lpx_session_debug_stamps(lpx_session* s, unsigned long long* out8);

/* Work of this thread's last lpx_bnb_simplex(_batched) call: stats4[0] = floating-point operations of all
 * node LPs (sum of pivots x (2 m_d (n+m_d+1) + (n+m_d+1)), m_d = constraint rows of that node's tableau:
 * SURVEY.md 8d), [1] = seconds inside the GPU evaluation calls (upload, kernels, download, sync),
 * [2] = seconds of the whole search, [3] = evaluation rounds (kernel launch groups). */
int lpx_bnb_last_stats(double* stats4);

/* ---- counters (for benchmarks: "how many of my kernels launched") --------------------------- */
long long lpx_kernel_launches(void);   /* since lpx_init / last reset */
void lpx_reset_counters(void);

#ifdef __cplusplus
}
#endif
#endif /* LPX_H_ */
