/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement of the reference's solvers
 * (Jellyman750/Linear_Programming_Solver_LPR381, R = Linear_Programming_Solver/):
 *   R/Models/LPParser.cs, R/Models/PrimalSimplex.cs, R/Models/DualSimplex.cs,
 *   R/Models/Branch&Bound.cs, R/Models/BranchAndBoundKnapsack.cs, R/Models/LPSolver.cs.
 * The reference cannot be compiled here (C#, net8.0-windows, no .NET toolchain), and it ships
 * no tests, fixtures or golden vectors, so PARITY IS UNPINNED by the reference itself; the
 * known-answer cases in tests/golden/ pin this restatement instead.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and --impl reference)
 * may load this library.  The product (liblpx.so and the host layer) never does.
 */
#ifndef ORACLE_H_
#define ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* status codes: 0 OPTIMAL, 1 UNBOUNDED, 2 INFEASIBLE(dual); <0 = the reference's exceptions:
 * -1 '>=' row, -2 negative RHS, -3 iteration limit, -4 bad arguments, -6 parse error,
 * -7 unsupported algorithm.  sense: 0 Max, 1 Min.  rel: 0 LE, 1 GE, 2 EQ. */

void orc_set_newline(const char* nl);
const char* orc_last_error(void);

/* Phase 1: A == NULL -> only *m,*n are written.  Phase 2: arrays sized from phase 1. */
int orc_parse_text(const char* text, int* sense, int* m, int* n, double* A, int* rel, double* b, double* c);

/* Tableau geometry after EQ expansion: rows = m' + 1, cols = n + m' + 1. */
void orc_tableau_dims(int m, int n, const int* rel, int* rows, int* cols);

/* PrimalSimplex.Solve.  history (nullable) receives the tableau after iteration 0..k
 * (at most history_cap tableaux).  pivots = (entering, leaving) pairs. */
int orc_primal_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                     int max_iterations, int* status, int* n_pivots, int* pivots, int pivots_cap, int* basis,
                     double* x, double* z, double* tableau, double* history, int history_cap);

/* DualSimplex.Solve; *silent = ForceDualFeasibility pivots (listed first in pivots). */
int orc_dual_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                   int* status, int* n_pivots, int* silent, int* pivots, int pivots_cap, int* basis, double* x,
                   double* z, double* tableau, double* history, int history_cap);

/* Timed CPU baseline: `count` all-<= Max problems of one shape, arithmetic loop
 * (ChooseEntering + ChooseLeaving + Pivot) on `threads` host threads.  with_format != 0 also
 * runs the reference's unconditional per-iteration tableau formatting.  Returns total pivots. */
long orc_primal_batch(int count, int m, int n, const double* A, const double* b, const double* c,
                      int max_iterations, int threads, int with_format, int* status, int* n_pivots, int* basis,
                      double* x, double* z, double* tableau);

/* Arithmetic-only pivots on a prebuilt (m+1) x width tableau, at most max_pivots of them. */
int orc_primal_core(double* T, int m, int width, int* basis, int max_pivots, int* n_pivots, int* pivots,
                    int pivots_cap);

/* BranchAndBound.Solve.  Node arrays are filled in solve order up to node_cap entries. */
int orc_bnb_simplex(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                    int* found, double* best_z, double* best_x, int* n_nodes, long* total_pivots, int node_cap,
                    int* node_outcome, int* node_algo, int* node_pivots, double* node_z, int* node_branch_var,
                    int* node_depth);

/* BranchAndBoundKnapsack.Solve.  Eval arrays: one entry per ComputeRelaxation of root/children. */
/* Mode B, the "pooled tree" — NOT the reference's tree (oracle/orc_pooled.cpp): both children honoured, warm
 * started from the parent's tableau, best-bound rounds of `batch` nodes.  Per evaluated node, in commit order:
 * id, outcome (BNB_* as above), dual pivots, z. */
int orc_bnb_pooled(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c, int batch,
                   long max_nodes /* stop after the round that passes this many nodes; 0 = none */, int* found, double* best_z, double* best_x, long* n_nodes, long* total_pivots, long* rounds,
                   long* skipped, long node_cap, int* node_id, int* node_outcome, int* node_pivots, double* node_z);

int orc_knapsack(int n, const double* profit, const double* weight, double capacity, int* found, double* best,
                 int* best_x, long* n_evals, long* n_pops, long eval_cap, int* ev_parent, int* ev_child,
                 int* ev_var, double* ev_bound, double* ev_weight, int* ev_frac, int* ev_decision);

/* RevisedPrimalSimplex.Solve numerics: per pivot the entering column, the leaving ROW and theta; the
 * basis state after the last completed iteration (Bidx, x_B, B^-1 row-major m x m); x and the z*
 * recomputed from the original objective.  status: 0 OPTIMAL, 1 UNBOUNDED, -3 iteration limit,
 * -10 unsupported model (not all <= with b >= -1e-9), -11 singular basis. */
int orc_revised_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                      int max_iterations, int* status, int* n_iters, int* enter, int* leave, double* theta, int cap,
                      int* basis, double* xB, double* Binv, double* x, double* z_original);

/* Full text of a headless solve: the updatePivot stream, Report and Summary.
 * algorithm: any LPSolver key, "knapsack" for BranchAndBoundKnapsack, "cutting plane" for
 * CuttingPlane (both constructed directly by Form1.btnSolve_Click, not through LPSolver). */
typedef struct orc_text orc_text;
orc_text* orc_solve_text(const char* input, const char* algorithm);
int orc_text_code(const orc_text* t);           /* 0 or the negative error code */
const char* orc_text_error(const orc_text* t);  /* exception message or "" */
const char* orc_text_log(const orc_text* t);
const char* orc_text_report(const orc_text* t);
const char* orc_text_summary(const orc_text* t);
int orc_text_masks(const orc_text* t);          /* number of callback chunks */
long orc_text_chunk_len(const orc_text* t, int k);   /* characters of chunk k */
/* bool[,] highlight of chunk k (PrimalSimplex.cs:117-119, DualSimplex.cs:65-69,104-106): returns 1 and the
 * rows x cols cells, or 0 for a null mask */
int orc_text_mask(const orc_text* t, int k, int* rows, int* cols, unsigned char* bits, long cap);
/* numeric part of the returned SimplexResult; returns 1 when Tableau != null */
int orc_text_result_dims(const orc_text* t, int* rows, int* cols, int* nx, int* nbasis);
const double* orc_text_tableau(const orc_text* t);
const double* orc_text_solution(const orc_text* t);
const int* orc_text_basis(const orc_text* t);
double orc_text_z(const orc_text* t);
/* Gomory cuts added by "cutting plane", in order */
int orc_text_cut_count(const orc_text* t);
int orc_text_cut(const orc_text* t, int k, int* frac_var, int* row, double* a, double* b);
int orc_text_cut_end(const orc_text* t);        /* 0 integer, 1 incomplete (50 rounds), 2 LP error, 3 non-basic */
void orc_text_free(orc_text* t);

/* .NET formatting restatements, exposed for unit tests.  Return pointers valid until the next call
 * on the same thread. */
const char* orc_fmt_custom(double v, int decimals);
const char* orc_fmt_fixed(double v, int decimals);
const char* orc_fmt_roundtrip(double v);
double orc_math_round(double v, int digits);

#ifdef __cplusplus
}
#endif
#endif
