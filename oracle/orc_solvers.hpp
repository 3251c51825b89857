// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).
#pragma once
#include "orc_model.hpp"

namespace orc {

struct PrimalOptions {
    int max_iterations = 10000;     // PrimalSimplex.MaxIterations (PrimalSimplex.cs:54)
    bool format_every_iteration = true;  // the reference formats every tableau even with a null sink
};

// R/Models/PrimalSimplex.cs:57-127
Outcome primal_simplex(const Problem& original, const Sink& sink, Trace* trace, const PrimalOptions& opt = {});
// R/Models/DualSimplex.cs:15-114
Outcome dual_simplex(const Problem& original, const Sink& sink, Trace* trace, bool format_every_iteration = true);
// R/Models/LPSolver.cs:16-76
std::string normalize_algorithm_key(const std::string& algorithm);
Outcome lp_solver_solve(const Problem& p, const std::string& algorithm, const Sink& sink, Trace* trace = nullptr,
                        bool format_every_iteration = true);

// Arithmetic-only primal loop on a prebuilt tableau (ChooseEntering + ChooseLeaving + Pivot,
// PrimalSimplex.cs:205-257).  Returns status; used as the timed CPU baseline.
int primal_core(double* T, int m, int ncols_total, int* basis, int max_iterations, int* n_pivots,
                int* pivots, int pivots_cap);
int choose_entering(const double* T, int m, int width);
int choose_leaving(const double* T, int m, int width, int entering, double margin);
void pivot(double* T, int rows, int width, int row, int col);

// One node of BranchAndBound.SolveNode as seen by the arithmetic (Branch&Bound.cs:128-258).
struct BnbNode {
    std::string name;
    int depth = 0;
    int algo = 0;         // 0 primal, 1 dual
    int outcome = 0;      // see BNB_* below
    int n_pivots = 0;
    double z = 0;
    std::vector<double> x;
    int branch_var = -1;
    int floor_val = 0, ceil_val = 0;
};
enum { BNB_ERROR = 0, BNB_INVALID = 1, BNB_INFEASIBLE = 2, BNB_PRUNED = 3, BNB_INCUMBENT = 4,
       BNB_BRANCHED = 5, BNB_NOFRAC = 6, BNB_DEPTH = 7 };
struct BnbTrace {
    std::vector<BnbNode> nodes;   // in the order the reference solves them (root LP first, then SolveNode calls)
    bool found = false;
    double best_z = 0;
    std::vector<double> best_x;
    long total_pivots = 0;
};
// R/Models/Branch&Bound.cs:30-123
Outcome branch_and_bound(const Problem& p, const Sink& sink, BnbTrace* trace, bool format_every_iteration = true);

// Mode B, the "pooled tree" (NOT the reference's tree: see orc_pooled.cpp).  One entry per evaluated node,
// in commit order (node 0 = root).
struct PooledTrace {
    std::vector<int> node_id, node_status, node_pivots, node_outcome;
    std::vector<double> node_z;
    bool found = false;
    double best_z = 0;
    std::vector<double> best_x;
    long pivots = 0, rounds = 0, skipped = 0;
};
int bnb_pooled(const Problem& p, int batch, long max_nodes, PooledTrace* t);

struct KnapEval {           // one ComputeRelaxation call for the root or a child
    int parent_pop = -1;    // index of the expanding pop (-1 root)
    int child = 0;          // 0 left (x=0), 1 right (x=1)
    int var = -1;           // original index fixed
    double bound = 0, weight = 0;
    int frac_sorted = -1;
    int decision = 0;       // KN_* below
};
enum { KN_ROOT = 0, KN_INFEASIBLE = 1, KN_CANDIDATE_INT = 2, KN_PUSHED = 3, KN_DROPPED = 4 };
struct KnapTrace {
    std::vector<KnapEval> evals;
    std::vector<std::string> pop_labels;  // label of every popped node that was expanded or closed
    long pops = 0;
    bool found = false;
    double best = 0;
    std::vector<int> best_x;
};
// R/Models/BranchAndBoundKnapsack.cs:58-407
Outcome knapsack_bnb(const Problem& p, const Sink& sink, KnapTrace* trace, bool build_text = true);

// RevisedPrimalSimplex: one entry per pivot; the basis state after the last completed iteration.
struct RevTrace {
    std::vector<int> enter, leave;   // entering column, leaving ROW (position in the basis)
    std::vector<double> theta;
    std::vector<int> basis;          // Bidx
    std::vector<double> xB, Binv;    // m, m x m
    double z = 0;                    // c_B . x_B of the standardized model
    int status = 0;                  // 0 OPTIMAL, 1 UNBOUNDED
};
// R/Models/RevisedPrimalSimplex.cs:17-145
Outcome revised_primal_simplex(const Problem& p, const Sink& sink, RevTrace* trace, int max_iterations = 10000);

// One entry per round of CuttingPlane.Solve (one PrimalSimplex solve, then at most one cut).
enum { CUT_INTEGER = 0, CUT_INCOMPLETE = 1, CUT_LP_ERROR = 2, CUT_NONBASIC = 3 };
struct CutTrace {
    std::vector<int> lp_pivots, lp_status;      // per round
    std::vector<int> frac_var, cut_row;         // per cut: first fractional x_j, tableau row the cut was read from
    std::vector<std::vector<double>> cut_a;     // per cut: coefficients over the decision variables
    std::vector<double> cut_b;
    int end = CUT_INTEGER;
};
// R/Models/CuttingPlane.cs:13-164
Outcome cutting_plane(const Problem& p, const Sink& sink, CutTrace* trace);
// R/Models/CuttingPlaneRevised.cs:14-111 (cut_row is -1: the cut is a variable bound, not a tableau row)
Outcome cutting_plane_revised(const Problem& p, const Sink& sink, CutTrace* trace);

}  // namespace orc
