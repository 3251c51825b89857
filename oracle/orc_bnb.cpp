// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (no tests/fixtures upstream); pinned by tests/golden/ known-answer cases.
//
// CPU restatement of
//   R/Models/Branch&Bound.cs:30-303            (BranchAndBound: recursive DFS, ceil child first)
//   R/Models/BranchAndBoundKnapsack.cs:58-547  (best-first 0/1 knapsack with its own max-heap)
// Quirks are kept on purpose: every ">=" child is routed to Dual Simplex whose
// result carries no Solution/Tableau, so it is rejected as "Invalid Simplex result".
#include "orc_solvers.hpp"
#include "orc_dotnet.hpp"

#include <algorithm>
#include <cmath>
#include <limits>

namespace orc {

namespace {

const double BB_EPS = 1e-6;  // BranchAndBound.EPS
const int BB_MAX_DEPTH = 200;

std::string f3(double v) { return fmt_fixed(v, 3); }
std::string f6(double v) { return fmt_fixed(v, 6); }

const char* rel_name(int rel) { return rel == LE ? "LE" : rel == GE ? "GE" : "EQ"; }

std::string row_text(const Row& c) {
    std::string s;
    bool first = true;
    for (size_t j = 0; j < c.a.size(); j++) {
        if (c.a[j] != 0) {
            if (!first) s += " + ";
            first = false;
            s += f3(c.a[j]) + "x" + std::to_string(j + 1);
        }
    }
    s += std::string(" ") + rel_name(c.rel) + " " + f3(c.b);
    return s;
}

std::string vec_f3(const std::vector<double>& x) {
    std::string s;
    for (size_t i = 0; i < x.size(); i++) {
        if (i) s += ", ";
        s += f3(x[i]);
    }
    return s;
}

bool is_integral(const std::vector<double>& x) {
    for (double v : x)
        if (std::fabs(v - math_round0(v)) > BB_EPS) return false;
    return true;
}

bool is_feasible(const std::vector<double>& x, const Problem& p) {
    for (const Row& c : p.rows) {
        double sum = 0;
        for (size_t i = 0; i < x.size(); i++) sum += c.a[i] * x[i];
        if (c.rel == LE && sum > c.b + BB_EPS) return false;
        if (c.rel == GE && sum < c.b - BB_EPS) return false;
        if (c.rel == EQ && std::fabs(sum - c.b) > BB_EPS) return false;
    }
    for (double v : x)
        if (v < -BB_EPS) return false;
    return true;
}

const char* choose_algorithm(const Problem& p) {
    for (const Row& c : p.rows)
        if (c.rel == GE || c.rel == EQ) return "Dual Simplex";
    return "Primal Simplex";
}

struct BnbState {
    const Sink& sink;
    BnbTrace* trace;
    bool fmt_every;
    int counter = 1;
    double best = -std::numeric_limits<double>::infinity();
    bool have_best = false;
    std::vector<double> best_x;

    void log(const std::string& msg) const {
        if (sink) sink(msg + g_newline, Mask{});
    }

    // returns false when the solve threw
    bool run_lp(const Problem& p, const char* algo, Outcome& res, std::string& err, BnbNode& node) {
        Trace t;
        node.algo = algo[0] == 'D' ? 1 : 0;
        try {
            res = lp_solver_solve(p, algo, sink, &t, fmt_every);
        } catch (const SolveError& ex) {
            err = ex.what();
            node.n_pivots = (int)t.enter.size();
            if (trace) trace->total_pivots += node.n_pivots;
            return false;
        }
        node.n_pivots = (int)t.enter.size();
        if (trace) trace->total_pivots += node.n_pivots;
        return true;
    }

    void solve_node(const Problem& problem, const std::string& name, const std::string& parent_id, int depth) {
        BnbNode node;
        node.name = name;
        node.depth = depth;
        auto done = [&](int outcome) {
            node.outcome = outcome;
            if (trace) trace->nodes.push_back(node);
        };
        if (depth > BB_MAX_DEPTH) {
            log(name + ": Maximum recursion depth reached \xE2\x86\x92 prune.");
            done(BNB_DEPTH);
            return;
        }
        {
            std::string s = name + ": Constraints: ";
            for (size_t i = 0; i < problem.rows.size(); i++) {
                if (i) s += "; ";
                s += row_text(problem.rows[i]);
            }
            log(s);
        }
        const char* algo = choose_algorithm(problem);
        log(name + ": Solving LP relaxation with " + algo + "...");
        Outcome res;
        std::string err;
        if (!run_lp(problem, algo, res, err, node)) {
            log(name + ": LP relaxation infeasible or error: " + err);
            done(BNB_ERROR);
            return;
        }
        if (!res.has_x || !res.has_tableau) {
            log(name + ": Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames).");
            done(BNB_INVALID);
            return;
        }
        std::vector<double> x(res.x.begin(), res.x.begin() + std::min<size_t>(res.x.size(), problem.nvars()));
        double z = res.z;
        node.z = z;
        node.x = x;
        log(name + " LP solution: z* = " + f3(z) + ", x* = [" + vec_f3(x) + "]");
        if (!is_feasible(x, problem)) {
            log(name + ": Solution x* = [" + vec_f3(x) + "] is infeasible for constraints.");
            done(BNB_INFEASIBLE);
            return;
        }
        if (z <= best + BB_EPS) {
            log(name + ": Pruned by bound (z* \xE2\x89\xA4 current best " + f3(best) + ").");
            done(BNB_PRUNED);
            return;
        }
        if (is_integral(x)) {
            best = z;
            have_best = true;
            best_x.clear();
            for (double v : x) best_x.push_back(math_round0(v));
            log(name + " is integer feasible. Updated BestObjective = " + f3(best));
            done(BNB_INCUMBENT);
            return;
        }
        int frac_index = -1;
        double min_dist = std::numeric_limits<double>::max();
        for (int i = 0; i < (int)x.size(); i++) {
            double frac = x[i] - std::floor(x[i]);
            if (frac > BB_EPS && (1 - frac) > BB_EPS) {
                double dist = std::fabs(frac - 0.5);
                log("Checking x" + std::to_string(i + 1) + " = " + f6(x[i]) + ", fracPart = " + f6(frac) +
                    ", distance to 0.5 = " + f6(dist));
                if (dist < min_dist || (dist == min_dist && i < frac_index)) {
                    min_dist = dist;
                    frac_index = i;
                }
            }
        }
        if (frac_index == -1) {
            log(name + ": No fractional variable found but solution not integral \xE2\x86\x92 prune.");
            done(BNB_NOFRAC);
            return;
        }
        double frac_val = x[frac_index];
        int floor_val = (int)std::floor(frac_val);
        int ceil_val = (int)std::ceil(frac_val);
        node.branch_var = frac_index;
        node.floor_val = floor_val;
        node.ceil_val = ceil_val;
        std::string xn = "x" + std::to_string(frac_index + 1);
        log(name + ": Branching on " + xn + " = " + f3(frac_val) + " (floor=" + std::to_string(floor_val) +
            ", ceil=" + std::to_string(ceil_val) + ")");
        std::string ceil_id = parent_id.empty() ? std::to_string(counter) : parent_id + "." + std::to_string(counter);
        std::string floor_id =
            parent_id.empty() ? std::to_string(counter + 1) : parent_id + "." + std::to_string(counter + 1);
        counter += 2;

        Problem left = problem, right = problem;
        Row unit;
        unit.a.assign(problem.nvars(), 0.0);
        unit.a[frac_index] = 1.0;
        unit.rel = LE;
        unit.b = floor_val;
        left.rows.push_back(unit);
        unit.rel = GE;
        unit.b = ceil_val;
        right.rows.push_back(unit);

        std::string left_name = "Subproblem " + floor_id + ": " + xn + " <= " + std::to_string(floor_val);
        std::string right_name = "Subproblem " + ceil_id + ": " + xn + " >= " + std::to_string(ceil_val);
        log(name + ": \xE2\x86\x92 " + right_name + " (ceil first)");
        log(name + ": \xE2\x86\x92 " + left_name);
        done(BNB_BRANCHED);
        solve_node(right, right_name, ceil_id, depth + 1);
        solve_node(left, left_name, floor_id, depth + 1);
    }
};

}  // namespace

Outcome branch_and_bound(const Problem& problem, const Sink& sink, BnbTrace* trace, bool fmt_every) {
    BnbState st{sink, trace, fmt_every, 1, -std::numeric_limits<double>::infinity(), false, {}};
    st.log("=== Branch & Bound Algorithm ===");
    {
        std::string s = "Objective: Maximize ";
        for (size_t i = 0; i < problem.c.size(); i++) {
            if (i) s += " + ";
            s += f3(problem.c[i]) + "x" + std::to_string(i + 1);
        }
        st.log(s);
    }
    st.log("Subject to:");
    for (const Row& c : problem.rows) st.log(row_text(c));
    st.log("x_j >= 0, integer");
    const char* root_algo = choose_algorithm(problem);
    st.log(std::string("Branch & Bound: Using ") + root_algo + " for the ROOT LP relaxation.");

    Outcome root;
    std::string err;
    BnbNode rnode;
    rnode.name = "ROOT LP";
    bool ok = st.run_lp(problem, root_algo, root, err, rnode);
    auto error_result = [&](const char* rep, const char* sum) {
        Outcome o;
        o.report = rep;
        o.summary = sum;
        if (trace) {
            trace->nodes.push_back(rnode);
            trace->found = false;
        }
        return o;
    };
    if (!ok) {
        st.log("Root Problem: LP relaxation infeasible or error: " + err);
        rnode.outcome = BNB_ERROR;
        return error_result("LP relaxation infeasible", "Error: Infeasible");
    }
    if (!root.has_x || !root.has_tableau) {
        st.log("Root Problem: Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames).");
        rnode.outcome = BNB_INVALID;
        return error_result("Invalid Simplex result", "Error: Invalid result");
    }
    std::vector<double> x_root(root.x.begin(), root.x.begin() + std::min<size_t>(root.x.size(), problem.nvars()));
    double z_root = root.z;
    rnode.z = z_root;
    rnode.x = x_root;
    st.log("Root Problem LP solution: z* = " + f3(z_root) + ", x* = [" + vec_f3(x_root) + "]");
    st.log("Root Problem optimal tableau displayed above.");

    auto build_report = [&]() {
        std::string sb = "Branch & Bound Finished." + g_newline;
        if (!st.have_best) {
            sb += "No integer-feasible solution found." + g_newline;
        } else {
            sb += "Best integer z* = " + f3(st.best) + g_newline;
            sb += "Best integer x* = [" + vec_f3(st.best_x) + "]" + g_newline;
        }
        Outcome o;
        o.report = sb;
        o.summary = sb;
        o.z = st.best;
        o.has_x = st.have_best;
        o.x = st.best_x;
        o.has_tableau = true;
        o.T = root.T;
        o.rows = root.rows;
        o.cols = root.cols;
        o.basis = root.basis;
        o.names = root.names;
        if (trace) {
            trace->found = st.have_best;
            trace->best_z = st.best;
            trace->best_x = st.best_x;
        }
        return o;
    };

    if (is_integral(x_root) && is_feasible(x_root, problem)) {
        st.best = z_root;
        st.have_best = true;
        for (double v : x_root) st.best_x.push_back(math_round0(v));
        st.log("Root Problem is already integral and feasible. Branch & Bound not required.");
        rnode.outcome = BNB_INCUMBENT;
        if (trace) trace->nodes.push_back(rnode);
        return build_report();
    }
    rnode.outcome = BNB_BRANCHED;
    if (trace) trace->nodes.push_back(rnode);
    st.log("Root solution is fractional \xE2\x86\x92 starting Branch & Bound.");
    st.solve_node(problem, "Root Problem", "", 0);
    return build_report();
}

// ================================================================================================
// BranchAndBoundKnapsack
// ================================================================================================

namespace {

const double KN_EPS = 1e-9;

struct Item {
    int index;
    double profit, weight;
    double ratio() const { return weight > 0 ? profit / weight : std::numeric_limits<double>::infinity(); }
};

// double.CompareTo
int cmp_double(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (std::isnan(a)) return std::isnan(b) ? 0 : -1;
    return 1;
}

struct KNode {
    std::vector<int> assigned;  // -1 undecided, 0, 1
    std::string label;
    double bound = 0, relax_profit = 0, relax_weight = 0;
    int pop_id = -1;
};

// SimpleMaxHeap<T> (BranchAndBoundKnapsack.cs:494-547)
struct MaxHeap {
    std::vector<KNode> data;
    static int cmp(const KNode& a, const KNode& b) { return cmp_double(a.bound, b.bound); }
    void push(KNode item) {
        data.push_back(std::move(item));
        int ci = (int)data.size() - 1;
        while (ci > 0) {
            int pi = (ci - 1) / 2;
            if (cmp(data[ci], data[pi]) <= 0) break;
            std::swap(data[ci], data[pi]);
            ci = pi;
        }
    }
    KNode pop() {
        int li = (int)data.size() - 1;
        std::swap(data[0], data[li]);
        KNode ret = std::move(data[li]);
        data.pop_back();
        int i = 0;
        li = (int)data.size() - 1;
        while (true) {
            int l = 2 * i + 1, r = 2 * i + 2, largest = i;
            if (l <= li && cmp(data[l], data[largest]) > 0) largest = l;
            if (r <= li && cmp(data[r], data[largest]) > 0) largest = r;
            if (largest == i) break;
            std::swap(data[i], data[largest]);
            i = largest;
        }
        return ret;
    }
};

struct Relax {
    std::vector<double> relaxed;
    double bound = 0;
    int frac_sorted = -1;
    double profit = 0, weight = 0;
};

struct KnapState {
    int n = 0;
    double capacity = 0;
    std::vector<Item> by_ratio;
    std::vector<int> rank_of;  // original index -> position in by_ratio

    // ComputeRelaxation (BranchAndBoundKnapsack.cs:431-491)
    Relax relax(const std::vector<int>& assigned) const {
        Relax r;
        r.relaxed.assign(n, 0.0);
        double weight = 0.0, profit = 0.0;
        int frac_sorted = -1;
        for (int i = 0; i < n; i++) {
            if (assigned[i] == 1) {
                const Item& it = by_ratio[rank_of[i]];
                r.relaxed[i] = 1.0;
                weight += it.weight;
                profit += it.profit;
            }
        }
        if (weight > capacity + KN_EPS) {
            r.bound = profit;
            r.profit = profit;
            r.weight = weight;
            return r;
        }
        for (int s = 0; s < n; s++) {
            const Item& it = by_ratio[s];
            int orig = it.index;
            if (assigned[orig] == 1) continue;
            if (assigned[orig] == 0) {
                r.relaxed[orig] = 0.0;
                continue;
            }
            if (weight + it.weight <= capacity + KN_EPS) {
                r.relaxed[orig] = 1.0;
                weight += it.weight;
                profit += it.profit;
            } else {
                double remain = capacity - weight;
                if (remain > KN_EPS && it.weight > KN_EPS) {
                    double frac = remain / it.weight;
                    r.relaxed[orig] = frac;
                    profit += it.profit * frac;
                    weight += it.weight * frac;
                    frac_sorted = s;
                }
                break;
            }
        }
        r.bound = profit;
        r.frac_sorted = frac_sorted;
        r.profit = profit;
        r.weight = weight;
        return r;
    }
};

}  // namespace

Outcome knapsack_bnb(const Problem& problem, const Sink& sink_in, KnapTrace* trace, bool build_text) {
    if (problem.rows.size() != 1)
        throw SolveError(ERR_BAD_ARGS, "Knapsack solver requires exactly one constraint (weights and capacity).");
    const Row& cons = problem.rows[0];
    if (cons.rel != LE) throw SolveError(ERR_BAD_ARGS, "Knapsack solver requires a <= constraint.");

    KnapState ks;
    ks.n = problem.nvars();
    ks.capacity = cons.b;
    const int n = ks.n;
    for (int i = 0; i < n; i++) ks.by_ratio.push_back(Item{i, problem.c[i], cons.a[i]});
    // OrderByDescending(Ratio).ThenByDescending(Profit): stable
    std::stable_sort(ks.by_ratio.begin(), ks.by_ratio.end(), [](const Item& a, const Item& b) {
        int c = cmp_double(a.ratio(), b.ratio());
        if (c != 0) return c > 0;
        return cmp_double(a.profit, b.profit) > 0;
    });
    ks.rank_of.assign(n, 0);
    for (int s = 0; s < n; s++) ks.rank_of[ks.by_ratio[s].index] = s;

    double best_value = -std::numeric_limits<double>::infinity();
    std::vector<int> best_x(n, 0);

    std::string report;
    size_t flush_pos = 0;
    const std::string& NL = g_newline;
    auto line = [&](const std::string& s) {
        if (build_text) report += s + NL;
    };
    auto flush = [&]() {
        if (!sink_in || !build_text) return;
        if (report.size() > flush_pos) {
            std::string delta = report.substr(flush_pos);
            flush_pos = report.size();
            sink_in(delta, Mask{});
        }
    };
    auto log_vector = [&](const std::vector<double>& relaxed, int frac_sorted) {
        if (!build_text) return;
        int frac_orig = frac_sorted >= 0 ? ks.by_ratio[frac_sorted].index : -1;
        for (int i = 0; i < n; i++)
            report += std::string(i == frac_orig ? ">" : " ") + "\tx" + std::to_string(i + 1) + "\t=\t" +
                      fmt_custom(relaxed[i]) + NL;
    };

    line("Branch and Bound Knapsack Algorithm");
    line("=================================================");
    line("Ratio Test:");
    line("Item\tci/ai\tRank");
    for (int i = 0; i < n; i++) {
        const Item& it = ks.by_ratio[i];
        line(std::to_string(it.index + 1) + "\t" + fmt_custom(it.ratio()) + "\t" + std::to_string(i + 1));
    }
    line("");
    line("-------------------------------------------------");

    MaxHeap pq;
    KNode root;
    root.assigned.assign(n, -1);
    root.label = "0";
    {
        Relax rr = ks.relax(root.assigned);
        root.bound = rr.bound;
        root.relax_profit = rr.profit;
        root.relax_weight = rr.weight;
        if (trace) trace->evals.push_back(KnapEval{-1, 0, -1, rr.bound, rr.weight, rr.frac_sorted, KN_ROOT});
    }
    pq.push(root);
    int pop_counter = 0;

    while (!pq.data.empty()) {
        KNode node = pq.pop();
        if (trace) trace->pops++;
        if (node.bound <= best_value + KN_EPS) continue;
        Relax cur = ks.relax(node.assigned);
        int this_pop = pop_counter++;
        if (trace) trace->pop_labels.push_back(node.label);

        line(node.label == "0" ? "Sub-Problem 0" : "Sub-Problem " + node.label);
        line("");
        log_vector(cur.relaxed, cur.frac_sorted);
        line("");

        if (cur.frac_sorted == -1) {
            if (cur.weight <= ks.capacity + KN_EPS) {
                double cand = cur.profit;
                line("\tz = " + fmt_custom(math_round(cand, 6)));
                if (cand > best_value + KN_EPS) {
                    best_value = cand;
                    for (int i = 0; i < n; i++) best_x[i] = cur.relaxed[i] >= 0.5 ? 1 : 0;
                    line("\tBEST CANDIDATE");
                } else {
                    line("\tCANDIDATE");
                }
            } else {
                line("\tINFEASIBLE");
            }
            line("------------------------------------------------");
            flush();
            continue;
        }
        int orig = ks.by_ratio[cur.frac_sorted].index;
        std::string left_label = node.label == "0" ? "1" : node.label + ".1";
        std::string right_label = node.label == "0" ? "2" : node.label + ".2";
        line("------------------------------------------------");
        line("");
        flush();

        for (int side = 0; side < 2; side++) {
            const std::string& label = side == 0 ? left_label : right_label;
            std::vector<int> assigned = node.assigned;
            assigned[orig] = side;
            Relax cr = ks.relax(assigned);
            line("-- Node " + label + " branching (x" + std::to_string(orig + 1) + "=" + std::to_string(side) + "):");
            log_vector(cr.relaxed, cr.frac_sorted);
            line("\tBound = " + fmt_custom(cr.bound) + ", Capacity = " + fmt_custom(cr.weight));
            int decision;
            if (cr.weight > ks.capacity + KN_EPS) {
                line(side == 0 ? "\tINFEASIBLE" : "\tINFEASIBLE ");
                decision = KN_INFEASIBLE;
            } else if (cr.bound > best_value + KN_EPS) {
                bool all_int = true;
                for (double v : cr.relaxed)
                    if (!(std::fabs(v - math_round0(v)) < KN_EPS)) all_int = false;
                if (all_int) {
                    line("\tCANDIDATE " + label);
                    if (cr.profit > best_value + KN_EPS) {
                        best_value = cr.profit;
                        for (int i = 0; i < n; i++) best_x[i] = (int)math_round0(cr.relaxed[i]);
                    }
                    decision = KN_CANDIDATE_INT;
                } else {
                    KNode child;
                    child.assigned = std::move(assigned);
                    child.label = label;
                    child.bound = cr.bound;
                    child.relax_profit = cr.profit;
                    child.relax_weight = cr.weight;
                    pq.push(std::move(child));
                    decision = KN_PUSHED;
                }
            } else {
                line("\tCANDIDATE " + label);
                decision = KN_DROPPED;
            }
            line("------------------------------------------------");
            line("");
            flush();
            if (trace) trace->evals.push_back(KnapEval{this_pop, side, orig, cr.bound, cr.weight, cr.frac_sorted, decision});
        }
    }

    const bool none = std::isinf(best_value) && best_value < 0;
    auto join_x = [&]() {
        std::string s;
        for (int i = 0; i < n; i++) {
            if (i) s += ", ";
            s += std::to_string(best_x[i]);
        }
        return s;
    };
    line("");
    line("Final Report:");
    line("Branch & Bound Knapsack Finished.");
    line("");
    if (none) {
        line("Status: NO FEASIBLE CANDIDATE");
    } else {
        line("Status: BEST CANDIDATE FOUND");
        for (int j = 0; j < n; j++) line("  x" + std::to_string(j + 1) + " = " + std::to_string(best_x[j]));
        line("  z* = " + fmt_custom(math_round(best_value, 6)));
    }
    line("");
    line("");
    line("Summary:");
    if (none) {
        line("No feasible candidate found.");
    } else {
        line("Best Candidate = " + fmt_custom(math_round(best_value, 6)));
        line("Best x* = [" + join_x() + "]");
    }
    flush();

    std::string fin;
    fin += "Final Report:" + NL + "Branch & Bound Knapsack Finished." + NL + NL;
    if (none) {
        fin += "Status: INFEASIBLE" + NL;
    } else {
        fin += "Status: BEST CANDIDATE FOUND" + NL;
        for (int i = 0; i < n; i++) fin += "  x" + std::to_string(i + 1) + " = " + std::to_string(best_x[i]) + NL;
        fin += "  z* = " + fmt_custom(best_value) + NL;
    }
    fin += NL + "Summary:" + NL;
    if (none) {
        fin += "No feasible candidate found." + NL;
    } else {
        fin += "Best Candidate = " + fmt_custom(best_value) + NL;
        fin += "Best x* = [" + join_x() + "]" + NL;
    }
    if (trace) {
        trace->found = !none;
        trace->best = best_value;
        trace->best_x = best_x;
    }
    Outcome o;
    o.report = fin;
    o.summary = "";
    return o;
}

}  // namespace orc
