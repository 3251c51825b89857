// Host-side BranchAndBound and BranchAndBoundKnapsack: the reference's log text
// (R/Models/Branch&Bound.cs:36-123,128-258; R/Models/BranchAndBoundKnapsack.cs:85-405) replayed
// from the records liblpx delivers in the reference's own order.  No arithmetic on the path
// happens here; the "Checking x.." lines only re-format numbers the engine returned.
#include <cmath>
#include <limits>

#include "../../include/lpx.h"
#include "dotnet_text.hpp"
#include "host_util.hpp"
#include "lp_model.hpp"

namespace lpr381 {

using text::custom_hash;
using text::fixed;

namespace {

const char* rel_name(int r) { return r == 0 ? "LE" : r == 1 ? "GE" : "EQ"; }

std::string vec_f3(const double* x, int n) {
    std::string s;
    for (int i = 0; i < n; i++) s += (i ? ", " : "") + fixed(x[i], 3);
    return s;
}

std::string row_text(const double* a, int n, int rel, double b) {
    std::string s;
    bool first = true;
    for (int j = 0; j < n; j++)
        if (a[j] != 0) {
            s += (first ? "" : " + ") + fixed(a[j], 3) + "x" + std::to_string(j + 1);
            first = false;
        }
    return s + " " + rel_name(rel) + " " + fixed(b, 3);
}

struct BnbReplay {
    const LPProblem& problem;
    Flat flat;
    UpdatePivot cb;
    double best = -std::numeric_limits<double>::infinity();
    // per record: the unit row that created the node and the parent, to rebuild its constraint list
    struct Rec {
        int parent, var, rel;
        double rhs;
        std::string name;
    };
    std::vector<Rec> recs;
    SimplexResult root;  // Tableau/Basis/VarNames of the root LP for BuildReport
    bool have_root = false;

    void log(const std::string& msg) const {
        if (cb) cb(msg + NewLine(), Highlight());
    }

    std::string constraints_text(int rec) const {
        std::vector<const Rec*> chain;
        for (int k = rec; k >= 0 && recs[k].var >= 0; k = recs[k].parent) chain.push_back(&recs[k]);
        std::string s;
        for (int i = 0; i < flat.m; i++)
            s += (i ? "; " : "") + row_text(flat.A.data() + (size_t)i * flat.n, flat.n, flat.rel[i], flat.b[i]);
        std::vector<double> unit(flat.n);
        for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
            std::fill(unit.begin(), unit.end(), 0.0);
            unit[(*it)->var] = 1.0;
            s += "; " + row_text(unit.data(), flat.n, (*it)->rel, (*it)->rhs);
        }
        return s;
    }

    // the iteration text of one node's LP, from its history and pivots
    void lp_text(const lpx_bnb_node& nd) const {
        if (!cb || nd.n_history <= 0) return;
        const int rows = nd.rows, cols = nd.cols, n = flat.n;
        const std::vector<std::string> names = var_names(n, rows - 1);
        std::vector<int> basis(rows - 1);
        for (int i = 0; i < rows - 1; i++) basis[i] = n + i;
        for (int k = 0; k < nd.silent_pivots; k++) basis[nd.pivots[2 * k + 1]] = nd.pivots[2 * k];
        const char* title = nd.algo == 1 ? "DUAL SIMPLEX TABLEAU Iteration" : "TABLEAU Iteration";
        const size_t tsize = (size_t)rows * cols;
        for (int k = 0; k < nd.n_history; k++) {
            Highlight hl;
            if (k > 0) {
                const int e = nd.pivots[2 * (nd.silent_pivots + k - 1)], l = nd.pivots[2 * (nd.silent_pivots + k - 1) + 1];
                basis[l] = e;
                hl = cross(rows, cols, l, e);
            }
            cb(tableau_text(title, nd.history + (size_t)k * tsize, rows, cols, basis, names, k), hl);
        }
        if (nd.algo == 1 && nd.lp_status == LPX_OPTIMAL) {
            Highlight hl;
            hl.rows = rows;
            hl.cols = cols;
            hl.v.assign(tsize, 0);
            for (int j = 0; j < cols; j++) hl.v[j] = 1;
            cb(tableau_text(title, nd.history + (size_t)(nd.n_history - 1) * tsize, rows, cols, basis, names,
                            nd.n_history),
               hl);
        }
    }

    std::string lp_error(const lpx_bnb_node& nd) const {
        if (nd.algo == 1 && nd.lp_status == LPX_S_ITER_LIMIT) return "Iteration limit exceeded (Dual Simplex).";
        return lpx_status_message(nd.lp_status);
    }

    void keep_root(const lpx_bnb_node& nd) {
        if (nd.algo != 0 || nd.lp_status < 0) return;
        root.Tableau.rows = nd.rows;
        root.Tableau.cols = nd.cols;
        if (nd.n_history > 0)
            root.Tableau.v.assign(nd.history + (size_t)(nd.n_history - 1) * nd.rows * nd.cols,
                                  nd.history + (size_t)nd.n_history * nd.rows * nd.cols);
        root.Basis.resize(nd.rows - 1);
        for (int i = 0; i < nd.rows - 1; i++) root.Basis[i] = flat.n + i;
        for (int k = 0; k < nd.n_pivots; k++) root.Basis[nd.pivots[2 * k + 1]] = nd.pivots[2 * k];
        root.VarNames = var_names(flat.n, nd.rows - 1);
        have_root = true;
    }

    void on_node(const lpx_bnb_node& nd) {
        const int n = flat.n;
        Rec rec{nd.parent, nd.bound_var, nd.is_ceil_child ? 1 : 0, (double)nd.bound_val, ""};
        if (nd.index == 0) {  // root LP of Solve (Branch&Bound.cs:53-95)
            rec.var = -1;
            recs.push_back(rec);
            lp_text(nd);
            keep_root(nd);
            if (nd.outcome == LPX_BNB_ERROR) {
                log("Root Problem: LP relaxation infeasible or error: " + lp_error(nd));
                return;
            }
            if (nd.outcome == LPX_BNB_INVALID) {
                log("Root Problem: Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames).");
                return;
            }
            log("Root Problem LP solution: z* = " + fixed(nd.z, 3) + ", x* = [" + vec_f3(nd.x, n) + "]");
            log("Root Problem optimal tableau displayed above.");
            if (nd.outcome == LPX_BNB_INCUMBENT) {
                best = nd.z;
                log("Root Problem is already integral and feasible. Branch & Bound not required.");
            } else {
                log("Root solution is fractional \xE2\x86\x92 starting Branch & Bound.");
            }
            return;
        }
        // SolveNode (Branch&Bound.cs:128-258)
        std::string name = "Root Problem";
        if (nd.bound_var >= 0) {
            std::string id;
            for (int k = 0; k < nd.id_path_len; k++) id += (k ? "." : "") + std::to_string(nd.id_path[k]);
            name = "Subproblem " + id + ": x" + std::to_string(nd.bound_var + 1) + (nd.is_ceil_child ? " >= " : " <= ") +
                   std::to_string(nd.bound_val);
        } else {
            rec.var = -1;
        }
        rec.name = name;
        recs.push_back(rec);
        const int me = (int)recs.size() - 1;
        if (nd.outcome == LPX_BNB_DEPTH) {
            log(name + ": Maximum recursion depth reached \xE2\x86\x92 prune.");
            return;
        }
        log(name + ": Constraints: " + constraints_text(me));
        log(name + ": Solving LP relaxation with " + (nd.algo == 1 ? "Dual Simplex" : "Primal Simplex") + "...");
        lp_text(nd);
        if (nd.outcome == LPX_BNB_ERROR) {
            log(name + ": LP relaxation infeasible or error: " + lp_error(nd));
            return;
        }
        if (nd.outcome == LPX_BNB_INVALID) {
            log(name + ": Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames).");
            return;
        }
        log(name + " LP solution: z* = " + fixed(nd.z, 3) + ", x* = [" + vec_f3(nd.x, n) + "]");
        if (nd.outcome == LPX_BNB_INFEASIBLE) {
            log(name + ": Solution x* = [" + vec_f3(nd.x, n) + "] is infeasible for constraints.");
            return;
        }
        if (nd.outcome == LPX_BNB_PRUNED) {
            log(name + ": Pruned by bound (z* \xE2\x89\xA4 current best " + fixed(best, 3) + ").");
            return;
        }
        if (nd.outcome == LPX_BNB_INCUMBENT) {
            best = nd.z;
            log(name + " is integer feasible. Updated BestObjective = " + fixed(best, 3));
            return;
        }
        for (int i = 0; i < n; i++) {  // the "Checking" lines of the branching-variable scan (:200-213)
            const double frac = nd.x[i] - std::floor(nd.x[i]);
            if (frac > 1e-6 && (1 - frac) > 1e-6)
                log("Checking x" + std::to_string(i + 1) + " = " + fixed(nd.x[i], 6) + ", fracPart = " + fixed(frac, 6) +
                    ", distance to 0.5 = " + fixed(std::fabs(frac - 0.5), 6));
        }
        if (nd.outcome == LPX_BNB_NOFRAC) {
            log(name + ": No fractional variable found but solution not integral \xE2\x86\x92 prune.");
            return;
        }
        const std::string xn = "x" + std::to_string(nd.branch_var + 1);
        log(name + ": Branching on " + xn + " = " + fixed(nd.x[nd.branch_var], 3) + " (floor=" +
            std::to_string(nd.floor_val) + ", ceil=" + std::to_string(nd.ceil_val) + ")");
        // child ids: the engine numbers them exactly like _subProblemCounter; the children's own
        // records carry them, but the two arrows are logged here, before the children run
        std::string prefix;
        for (int k = 0; k < nd.id_path_len; k++) prefix += std::to_string(nd.id_path[k]) + ".";
        const int ceil_id = next_counter, floor_id = next_counter + 1;
        next_counter += 2;
        log(name + ": \xE2\x86\x92 Subproblem " + prefix + std::to_string(ceil_id) + ": " + xn + " >= " +
            std::to_string(nd.ceil_val) + " (ceil first)");
        log(name + ": \xE2\x86\x92 Subproblem " + prefix + std::to_string(floor_id) + ": " + xn + " <= " +
            std::to_string(nd.floor_val));
    }
    int next_counter = 1;
};

void bnb_trampoline(const lpx_bnb_node* nd, void* user) { static_cast<BnbReplay*>(user)->on_node(*nd); }

}  // namespace

SimplexResult BranchAndBound::Solve(const LPProblem& problem, UpdatePivot updatePivot) {
    const std::string& nl = NewLine();
    BnbReplay rp{problem, flatten(problem), updatePivot, -std::numeric_limits<double>::infinity(), {}, {}, false};
    const Flat& f = rp.flat;
    rp.log("=== Branch & Bound Algorithm ===");
    {
        std::string s = "Objective: Maximize ";
        for (int i = 0; i < f.n; i++) s += (i ? " + " : "") + fixed(f.c[i], 3) + "x" + std::to_string(i + 1);
        rp.log(s);
    }
    rp.log("Subject to:");
    bool dual_root = false;
    for (int i = 0; i < f.m; i++) {
        rp.log(row_text(f.A.data() + (size_t)i * f.n, f.n, f.rel[i], f.b[i]));
        dual_root = dual_root || f.rel[i] != 0;
    }
    rp.log("x_j >= 0, integer");
    rp.log(std::string("Branch & Bound: Using ") + (dual_root ? "Dual Simplex" : "Primal Simplex") +
           " for the ROOT LP relaxation.");

    lpx_options opt;
    lpx_default_options(&opt);
    int found = 0, n_nodes = 0, root_status = 0;
    long long lp_pivots = 0;
    double best_z = 0;
    std::vector<double> best_x(f.n);
    // with a sink attached every node's iteration tableaux are needed for the log; without one the
    // tree runs without records and only the root's tableau/basis (part of the result, :118-120)
    // are fetched afterwards
    throw_on(lpx_bnb_simplex(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt,
                             updatePivot ? LPX_BNB_WANT_HISTORY : 0, &found, &best_z, best_x.data(), &n_nodes,
                             &lp_pivots, &root_status, updatePivot ? bnb_trampoline : nullptr, &rp));
    auto text_only = [](const char* report, const char* summary) {
        SimplexResult r;
        r.Report = report;
        r.Summary = summary;
        return r;
    };
    if (root_status < 0) return text_only("LP relaxation infeasible", "Error: Infeasible");
    if (dual_root) return text_only("Invalid Simplex result", "Error: Invalid result");
    if (!rp.have_root) {
        int rows = 0, cols = 0, st = 0, np = 0;
        throw_on(lpx_tableau_dims(f.m, f.n, f.rel.data(), &rows, &cols));
        rp.root.Tableau.rows = rows;
        rp.root.Tableau.cols = cols;
        rp.root.Tableau.v.resize((size_t)rows * cols);
        rp.root.Basis.resize(rows - 1);
        throw_on(lpx_primal_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &st, &np,
                                  nullptr, 0, rp.root.Basis.data(), nullptr, nullptr, rp.root.Tableau.v.data(), nullptr,
                                  0));
        rp.root.VarNames = var_names(f.n, rows - 1);
    }

    std::string sb = "Branch & Bound Finished." + nl;
    HasBest = found != 0;
    BestObjective = best_z;
    BestSolution = found ? best_x : std::vector<double>();
    if (!found) sb += "No integer-feasible solution found." + nl;
    else {
        sb += "Best integer z* = " + fixed(best_z, 3) + nl;
        sb += "Best integer x* = [" + vec_f3(best_x.data(), f.n) + "]" + nl;
    }
    SimplexResult r;
    r.Report = r.Summary = sb;
    r.OptimalValue = best_z;
    r.HasSolution = found != 0;
    r.Solution = BestSolution;
    r.Tableau = rp.root.Tableau;
    r.Basis = rp.root.Basis;
    r.VarNames = rp.root.VarNames;
    return r;
}

// ================================================================================================
// BranchAndBoundKnapsack
// ================================================================================================
namespace {

struct KnapReplay {
    int n;
    UpdatePivot cb;
    std::vector<int> rank_order;  // filled after the call for the header; the body needs it live
    std::string report;
    size_t flush_pos = 0;
    const double* profit;
    const double* weight;
    double capacity;

    void line(const std::string& s) { report += s + NewLine(); }
    void flush() {
        if (!cb) return;
        if (report.size() > flush_pos) {
            std::string delta = report.substr(flush_pos);
            flush_pos = report.size();
            cb(delta, Highlight());
        }
    }
    // relaxed[] of ComputeRelaxation from the assignment and where the greedy pass stopped
    std::vector<double> relaxed(const lpx_knap_eval& e) const {
        std::vector<double> r(n, 0.0);
        double fixed_w = 0.0;
        for (int i = 0; i < n; i++)
            if (e.assigned[i] == 1) {
                r[i] = 1.0;
                fixed_w += weight[i];
            }
        if (fixed_w > capacity + 1e-9) return r;  // early return: only the fixed items (:455-456)
        for (int s = 0; s < e.break_rank && s < n; s++)
            if (e.assigned[rank_order[s]] < 0) r[rank_order[s]] = 1.0;
        if (e.frac_rank >= 0) r[rank_order[e.frac_rank]] = e.frac;
        return r;
    }
    void vector_lines(const lpx_knap_eval& e) {
        const std::vector<double> r = relaxed(e);
        const int frac_orig = e.frac_rank >= 0 ? rank_order[e.frac_rank] : -1;
        for (int i = 0; i < n; i++)
            report += std::string(i == frac_orig ? ">" : " ") + "\tx" + std::to_string(i + 1) + "\t=\t" + custom_hash(r[i]) +
                      NewLine();
    }
    static std::string label_text(const int* l, int len) {
        std::string s;
        for (int k = 0; k < len; k++) s += (k ? "." : "") + std::to_string(l[k]);
        return s;
    }
    void on_pop(const lpx_knap_pop& p, const lpx_knap_eval* left, const lpx_knap_eval* right) {
        const std::string label = label_text(p.label, p.label_len);
        line(label == "0" ? "Sub-Problem 0" : "Sub-Problem " + label);
        line("");
        vector_lines(p.relax);
        line("");
        if (p.closed != 0) {
            if (p.closed == 3) line("\tINFEASIBLE");
            else {
                line("\tz = " + custom_hash(text::math_round(p.relax.bound, 6)));
                line(p.closed == 1 ? "\tBEST CANDIDATE" : "\tCANDIDATE");
            }
            line("------------------------------------------------");
            flush();
            return;
        }
        line("------------------------------------------------");
        line("");
        flush();
        const lpx_knap_eval* ch[2] = {left, right};
        for (int side = 0; side < 2; side++) {
            const lpx_knap_eval& e = *ch[side];
            const std::string cl = label == "0" ? std::to_string(side + 1) : label + "." + std::to_string(side + 1);
            line("-- Node " + cl + " branching (x" + std::to_string(e.var + 1) + "=" + std::to_string(side) + "):");
            vector_lines(e);
            line("\tBound = " + custom_hash(e.bound) + ", Capacity = " + custom_hash(e.weight));
            if (e.decision == LPX_KN_INFEASIBLE) line(side == 0 ? "\tINFEASIBLE" : "\tINFEASIBLE ");
            else if (e.decision == LPX_KN_CANDIDATE_INT || e.decision == LPX_KN_DROPPED) line("\tCANDIDATE " + cl);
            line("------------------------------------------------");
            line("");
            flush();
        }
    }
};

void knap_trampoline(const lpx_knap_pop* p, const lpx_knap_eval* l, const lpx_knap_eval* r, void* user) {
    static_cast<KnapReplay*>(user)->on_pop(*p, l, r);
}

}  // namespace

SimplexResult BranchAndBoundKnapsack::Solve(const LPProblem& problem, UpdatePivot updatePivot) {
    const std::string& nl = NewLine();
    if (problem.Constraints.size() != 1)
        throw LpException("Knapsack solver requires exactly one constraint (weights and capacity).");
    const Constraint& cons = problem.Constraints[0];
    if (cons.Relation != Rel::LE) throw LpException("Knapsack solver requires a <= constraint.");
    const int n = problem.NumVars();
    if ((int)cons.A.size() < n) throw LpException("Index was outside the bounds of the array.");

    KnapReplay rp;
    rp.n = n;
    rp.cb = updatePivot;
    rp.profit = problem.C.data();
    rp.weight = cons.A.data();
    rp.capacity = cons.B;
    // ratio ordering for the header (:75-94); the engine returns the same ordering (rank_order)
    rp.rank_order.resize(n);
    {
        // a zero-cost first call would be wasteful; the ordering is cheap host logic and is
        // cross-checked against the engine's below
        std::vector<int> idx(n);
        for (int i = 0; i < n; i++) idx[i] = i;
        auto ratio = [&](int i) { return cons.A[i] > 0 ? problem.C[i] / cons.A[i] : std::numeric_limits<double>::infinity(); };
        auto cmp = [](double a, double b) { return a < b ? -1 : a > b ? 1 : a == b ? 0 : (std::isnan(a) ? (std::isnan(b) ? 0 : -1) : 1); };
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            const int c = cmp(ratio(a), ratio(b));
            return c != 0 ? c > 0 : cmp(problem.C[a], problem.C[b]) > 0;
        });
        rp.rank_order = idx;
        rp.line("Branch and Bound Knapsack Algorithm");
        rp.line("=================================================");
        rp.line("Ratio Test:");
        rp.line("Item\tci/ai\tRank");
        for (int s = 0; s < n; s++)
            rp.line(std::to_string(idx[s] + 1) + "\t" + custom_hash(ratio(idx[s])) + "\t" + std::to_string(s + 1));
        rp.line("");
        rp.line("-------------------------------------------------");
    }
    lpx_options opt;
    lpx_default_options(&opt);
    int found = 0;
    double best = 0;
    std::vector<int> best_x(n), engine_order(n);
    long long evals = 0, pops = 0;
    throw_on(lpx_bnb_knapsack(n, problem.C.data(), cons.A.data(), cons.B, &opt, &found, &best, best_x.data(), &evals,
                              &pops, engine_order.data(), knap_trampoline, &rp));
    if (engine_order != rp.rank_order) throw LpException("internal: ratio ordering mismatch between host and engine");

    auto join_x = [&]() {
        std::string s;
        for (int i = 0; i < n; i++) s += (i ? ", " : "") + std::to_string(best_x[i]);
        return s;
    };
    rp.line("");
    rp.line("Final Report:");
    rp.line("Branch & Bound Knapsack Finished.");
    rp.line("");
    if (!found) rp.line("Status: NO FEASIBLE CANDIDATE");
    else {
        rp.line("Status: BEST CANDIDATE FOUND");
        for (int j = 0; j < n; j++) rp.line("  x" + std::to_string(j + 1) + " = " + std::to_string(best_x[j]));
        rp.line("  z* = " + custom_hash(text::math_round(best, 6)));
    }
    rp.line("");
    rp.line("");
    rp.line("Summary:");
    if (!found) rp.line("No feasible candidate found.");
    else {
        rp.line("Best Candidate = " + custom_hash(text::math_round(best, 6)));
        rp.line("Best x* = [" + join_x() + "]");
    }
    rp.flush();

    std::string fin = "Final Report:" + nl + "Branch & Bound Knapsack Finished." + nl + nl;
    if (!found) fin += "Status: INFEASIBLE" + nl;
    else {
        fin += "Status: BEST CANDIDATE FOUND" + nl;
        for (int i = 0; i < n; i++) fin += "  x" + std::to_string(i + 1) + " = " + std::to_string(best_x[i]) + nl;
        fin += "  z* = " + custom_hash(best) + nl;
    }
    fin += nl + "Summary:" + nl;
    if (!found) fin += "No feasible candidate found." + nl;
    else {
        fin += "Best Candidate = " + custom_hash(best) + nl;
        fin += "Best x* = [" + join_x() + "]" + nl;
    }
    SimplexResult r;
    r.Report = fin;
    r.Summary = "";
    return r;
}

}  // namespace lpr381
