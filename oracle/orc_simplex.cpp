// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (no tests/fixtures upstream); pinned by tests/golden/ known-answer cases.
//
// CPU restatement of the dense-tableau simplex solvers:
//   R/Models/PrimalSimplex.cs:57-304, R/Models/DualSimplex.cs:15-311, R/Models/LPSolver.cs:16-76
// Build: g++ -O2 -ffp-contract=off (separate multiply and subtract, true division — what
// RyuJIT x64 emits for the C# loops).
#include "orc_solvers.hpp"
#include "orc_dotnet.hpp"

#include <algorithm>
#include <cmath>
#include <limits>

namespace orc {

static const double kEps = 1e-9;  // PrimalSimplex.Eps, DualSimplex.Eps

// ---- arithmetic core -------------------------------------------------------------------------

// PrimalSimplex.ChooseEntering (PrimalSimplex.cs:205-220); width = total columns incl. RHS.
int choose_entering(const double* T, int m, int width) {
    const double* zr = T + (size_t)m * width;
    int best = -1;
    double min_val = -kEps;
    for (int j = 0; j < width - 1; j++) {
        if (zr[j] < min_val) {
            min_val = zr[j];
            best = j;
        }
    }
    return best;
}

// PrimalSimplex.ChooseLeaving (PrimalSimplex.cs:222-243); margin 1e-9 there,
// 1e-12 in DualSimplex.ForceDualFeasibility (DualSimplex.cs:212-222).
int choose_leaving(const double* T, int m, int width, int entering, double margin) {
    double best_ratio = std::numeric_limits<double>::infinity();
    int best_row = -1;
    for (int i = 0; i < m; i++) {
        double a = T[(size_t)i * width + entering];
        if (a > kEps) {
            double ratio = T[(size_t)i * width + width - 1] / a;
            if (ratio < best_ratio - margin) {
                best_ratio = ratio;
                best_row = i;
            }
        }
    }
    return best_row;
}

// PrimalSimplex.Pivot (PrimalSimplex.cs:245-257) == DualSimplex.Pivot (DualSimplex.cs:232-246).
void pivot(double* T, int rows, int width, int row, int col) {
    double* pr = T + (size_t)row * width;
    double piv = pr[col];
    for (int j = 0; j < width; j++) pr[j] /= piv;
    for (int i = 0; i < rows; i++) {
        if (i == row) continue;
        double* ri = T + (size_t)i * width;
        double f = ri[col];
        for (int j = 0; j < width; j++) ri[j] -= f * pr[j];
    }
}

int primal_core(double* T, int m, int width, int* basis, int max_iterations, int* n_pivots, int* pivots,
                int pivots_cap) {
    int iter = 1;
    *n_pivots = 0;
    while (true) {
        if (iter > max_iterations) return ERR_ITER_LIMIT;
        int e = choose_entering(T, m, width);
        if (e == -1) return 0;
        int l = choose_leaving(T, m, width, e, kEps);
        if (l == -1) return 1;
        pivot(T, m + 1, width, l, e);
        basis[l] = e;
        if (pivots && *n_pivots < pivots_cap) {
            pivots[2 * *n_pivots] = e;
            pivots[2 * *n_pivots + 1] = l;
        }
        ++*n_pivots;
        iter++;
    }
}

// ---- model preparation -------------------------------------------------------------------------

// PrimalSimplex.ExpandEqualitiesToInequalities (PrimalSimplex.cs:161-177)
static Problem expand_equalities(const Problem& model) {
    Problem ex;
    ex.c = model.c;  // sense resets to the default Max
    for (const Row& r : model.rows) {
        if (r.rel == EQ) {
            Row pos = r;
            pos.rel = LE;
            ex.rows.push_back(pos);
            Row neg = r;
            neg.rel = LE;
            for (double& v : neg.a) v *= -1;
            neg.b *= -1;
            ex.rows.push_back(neg);
        } else {
            ex.rows.push_back(r);
        }
    }
    return ex;
}

// BuildTableau (PrimalSimplex.cs:179-203, DualSimplex.cs:160-189): z-row is the LAST row.
static std::vector<double> build_tableau(const Problem& model, std::vector<int>& basis,
                                         std::vector<std::string>& names, int& rows, int& cols) {
    int m = (int)model.rows.size(), n = model.nvars(), s = m;
    rows = m + 1;
    cols = n + s + 1;
    std::vector<double> T((size_t)rows * cols, 0.0);
    for (int i = 0; i < m; i++) {
        const Row& r = model.rows[i];
        if ((int)r.a.size() < n) throw SolveError(ERR_BAD_ARGS, "Index was outside the bounds of the array.");
        for (int j = 0; j < n; j++) T[(size_t)i * cols + j] = r.a[j];
        T[(size_t)i * cols + n + i] = 1.0;
        T[(size_t)i * cols + n + s] = r.b;
    }
    for (int j = 0; j < n; j++) T[(size_t)m * cols + j] = -model.c[j];
    basis.resize(s);
    for (int i = 0; i < s; i++) basis[i] = n + i;
    names.resize(n + s);
    for (int j = 0; j < n; j++) names[j] = "x" + std::to_string(j + 1);
    for (int j = 0; j < s; j++) names[n + j] = "c" + std::to_string(j + 1);
    return T;
}

// ---- text ----------------------------------------------------------------------------------------

static std::string signed_terms(const std::vector<double>& v) {
    std::string out;
    for (size_t j = 0; j < v.size(); j++) {
        if (j) out += " ";
        out += (v[j] >= 0 ? "+" : "-");
        out += fmt_custom(std::fabs(v[j]));
        out += "x" + std::to_string(j + 1);
    }
    return out;
}

// PrimalSimplex.AppendCanonicalForm (PrimalSimplex.cs:259-270)
static void append_canonical(std::string& sb, const Problem& model) {
    sb += "Objective: max " + signed_terms(model.c) + g_newline;
    sb += "Subject to:" + g_newline;
    for (const Row& r : model.rows) {
        const char* rel = r.rel == LE ? "<=" : r.rel == GE ? ">=" : "=";
        sb += "  " + signed_terms(r.a) + " " + rel + " " + fmt_custom(r.b) + g_newline;
    }
    sb += "x >= 0" + g_newline;
}

// AppendTableau (PrimalSimplex.cs:272-304 / DualSimplex.cs:248-281); title differs.
static void append_tableau(std::string& sb, const char* title, const std::vector<double>& T, int rows, int cols,
                           const std::vector<int>& basis, const std::vector<std::string>& names, int iter) {
    const int m = rows - 1, ns = cols - 1, W = 12;
    sb += std::string(title) + " " + std::to_string(iter) + g_newline;
    sb += pad_left("Basis", W);
    for (int j = 0; j < ns; j++) sb += pad_left(names[j], W);
    sb += pad_left("RHS", W);
    sb += g_newline;
    sb += std::string((size_t)W * (ns + 2), '-') + g_newline;
    sb += pad_left("z", W);
    for (int j = 0; j <= ns; j++) sb += pad_left(fmt_custom(T[(size_t)m * cols + j]), W);
    sb += g_newline;
    for (int i = 0; i < m; i++) {
        sb += pad_left(names[basis[i]], W);
        for (int j = 0; j <= ns; j++) sb += pad_left(fmt_custom(T[(size_t)i * cols + j]), W);
        sb += g_newline;
    }
}

static int count_x(const std::vector<std::string>& names) {
    int n = 0;
    for (const auto& s : names)
        if (!s.empty() && s[0] == 'x') n++;
    return n;
}

// FinalizeReport (PrimalSimplex.cs:130-159; DualSimplex.cs:283-311 keeps only Report/Summary).
static Outcome finalize(std::string sb, std::vector<double>& T, int rows, int cols, std::vector<int>& basis,
                        std::vector<std::string>& names, const char* status, bool full) {
    int m = rows - 1, n = count_x(names);
    std::vector<double> x(n, 0.0);
    for (int i = 0; i < m; i++)
        if (basis[i] < n) x[basis[i]] = T[(size_t)i * cols + cols - 1];
    double z = T[(size_t)m * cols + cols - 1];
    sb += "\nStatus: " + std::string(status) + g_newline;
    for (int j = 0; j < n; j++)
        sb += "  x" + std::to_string(j + 1) + " = " + fmt_custom(math_round(x[j], 3)) + g_newline;
    sb += "  z* = " + fmt_custom(math_round(z, 3)) + g_newline;
    std::string summary = "Status: " + std::string(status) + g_newline;
    summary += "z* = " + fmt_custom(math_round(z, 3)) + g_newline;
    summary += "x* = [";
    for (int j = 0; j < n; j++) {
        if (j) summary += ", ";
        summary += fmt_roundtrip(math_round(x[j], 3));
    }
    summary += "]" + g_newline;
    Outcome o;
    o.report = sb;
    o.summary = summary;
    if (full) {
        o.z = z;
        o.has_x = true;
        o.x = x;
        o.has_tableau = true;
        o.T = T;
        o.rows = rows;
        o.cols = cols;
        o.basis = basis;
        o.names = names;
    }
    return o;
}

static Mask cross_mask(int rows, int cols, int row, int col) {
    Mask mk;
    mk.rows = rows;
    mk.cols = cols;
    mk.bits.assign((size_t)rows * cols, 0);
    for (int j = 0; j < cols; j++) mk.bits[(size_t)row * cols + j] = 1;
    for (int i = 0; i < rows; i++) mk.bits[(size_t)i * cols + col] = 1;
    return mk;
}

// ---- PrimalSimplex.Solve ---------------------------------------------------------------------

Outcome primal_simplex(const Problem& original, const Sink& sink, Trace* trace, const PrimalOptions& opt) {
    Problem model = original;
    if (model.sense == MIN)
        for (double& v : model.c) v = -v;
    for (const Row& r : model.rows) {
        if (r.rel == GE)
            throw SolveError(ERR_GE_ROW,
                             "Constraint contains '>=' sign. The Primal Simplex method cannot handle this. Please try "
                             "the Dual Simplex algorithm instead.");
        if (r.b < -1e-9)
            throw SolveError(ERR_NEG_RHS,
                             "Constraint has a negative RHS value. The Primal Simplex method cannot handle this. "
                             "Please try the Dual Simplex algorithm instead.");
    }
    Problem tm = expand_equalities(model);
    std::string report;
    append_canonical(report, tm);

    std::vector<int> basis;
    std::vector<std::string> names;
    int rows, cols;
    std::vector<double> T = build_tableau(tm, basis, names, rows, cols);
    const int m = rows - 1;
    const bool want_text = opt.format_every_iteration || (bool)sink;

    if (want_text) {
        std::string it;
        append_tableau(it, "TABLEAU Iteration", T, rows, cols, basis, names, 0);
        if (sink) sink(it, Mask{});
    }
    if (trace && trace->keep_history) trace->history.push_back(T);

    int iter = 1;
    while (true) {
        if (iter > opt.max_iterations) throw SolveError(ERR_ITER_LIMIT, "Iteration limit exceeded.");
        int entering = choose_entering(T.data(), m, cols);
        if (entering == -1) break;
        int leaving = choose_leaving(T.data(), m, cols, entering, kEps);
        if (leaving == -1) {
            report += "UNBOUNDED" + g_newline;
            if (trace) trace->status = 1;
            return finalize(report, T, rows, cols, basis, names, "UNBOUNDED", true);
        }
        pivot(T.data(), rows, cols, leaving, entering);
        basis[leaving] = entering;
        if (trace) {
            trace->enter.push_back(entering);
            trace->leave.push_back(leaving);
            if (trace->keep_history) trace->history.push_back(T);
        }
        if (want_text) {
            std::string it;
            append_tableau(it, "TABLEAU Iteration", T, rows, cols, basis, names, iter);
            Mask mk = cross_mask(rows, cols, leaving, entering);
            if (sink) sink(it, mk);
        }
        iter++;
    }
    if (trace) trace->status = 0;
    return finalize(report, T, rows, cols, basis, names, "OPTIMAL", true);
}

// ---- DualSimplex.Solve -----------------------------------------------------------------------

// DualSimplex.PrepareForTableau (DualSimplex.cs:117-158)
static Problem dual_prepare(const Problem& original) {
    Problem model = original;
    if (model.sense == MIN)
        for (double& v : model.c) v = -v;
    Problem ex;
    ex.c = model.c;
    ex.sense = MAX;
    for (const Row& cons : model.rows) {
        if (cons.rel == EQ) {
            Row pos = cons;
            pos.rel = LE;
            ex.rows.push_back(pos);
            Row neg = cons;
            neg.rel = LE;
            neg.b = -cons.b;
            for (double& v : neg.a) v *= -1;
            ex.rows.push_back(neg);
        } else {
            Row row = cons;
            if (row.rel == GE) {
                for (double& v : row.a) v *= -1;
                row.b *= -1;
                row.rel = LE;
            }
            if (row.b < -kEps) {  // flips a second time: "x >= c" ends up as "x <= c"
                for (double& v : row.a) v *= -1;
                row.b *= -1;
            }
            ex.rows.push_back(row);
        }
    }
    return ex;
}

// DualSimplex.ForceDualFeasibility (DualSimplex.cs:195-228)
static int force_dual_feasibility(std::vector<double>& T, int rows, int cols, std::vector<int>& basis,
                                  Trace* trace) {
    int m = rows - 1, done = 0;
    for (int guard = 0; guard < 100; guard++) {
        int entering = choose_entering(T.data(), m, cols);
        if (entering == -1) return done;
        int leave = choose_leaving(T.data(), m, cols, entering, 1e-12);
        if (leave == -1) return done;
        pivot(T.data(), rows, cols, leave, entering);
        basis[leave] = entering;
        done++;
        if (trace) {
            trace->enter.push_back(entering);
            trace->leave.push_back(leave);
        }
    }
    return done;
}

Outcome dual_simplex(const Problem& original, const Sink& sink, Trace* trace, bool format_every_iteration) {
    Problem model = dual_prepare(original);
    std::vector<int> basis;
    std::vector<std::string> names;
    int rows, cols;
    std::vector<double> T = build_tableau(model, basis, names, rows, cols);
    int silent = force_dual_feasibility(T, rows, cols, basis, trace);
    if (trace) trace->silent_pivots = silent;
    const bool want_text = format_every_iteration || (bool)sink;
    const char* title = "DUAL SIMPLEX TABLEAU Iteration";
    if (want_text) {
        std::string it;
        append_tableau(it, title, T, rows, cols, basis, names, 0);
        if (sink) sink(it, Mask{});
    }
    if (trace && trace->keep_history) trace->history.push_back(T);

    const int m = rows - 1, ns = cols - 1;
    int iter = 1;
    while (true) {
        if (iter > 10000) throw SolveError(ERR_ITER_LIMIT, "Iteration limit exceeded (Dual Simplex).");
        int leave = -1;
        double most_neg = -kEps;
        for (int i = 0; i < m; i++) {
            double rhs = T[(size_t)i * cols + ns];
            if (rhs < most_neg) {
                most_neg = rhs;
                leave = i;
            }
        }
        if (leave == -1) {
            if (want_text) {
                std::string it;
                append_tableau(it, title, T, rows, cols, basis, names, iter);
                Mask mk;
                mk.rows = rows;
                mk.cols = cols;
                mk.bits.assign((size_t)rows * cols, 0);
                for (int j = 0; j < cols; j++) mk.bits[j] = 1;  // "z row" assumed first (DualSimplex.cs:67)
                if (sink) sink(it, mk);
            }
            if (trace) trace->status = 0;
            return finalize("", T, rows, cols, basis, names, "OPTIMAL", false);
        }
        int enter = -1;
        double best_ratio = std::numeric_limits<double>::infinity();
        for (int j = 0; j < ns; j++) {
            double a = T[(size_t)leave * cols + j];
            if (a < -kEps) {
                double ratio = T[(size_t)m * cols + j] / (-a);
                if (ratio < best_ratio - 1e-12) {
                    best_ratio = ratio;
                    enter = j;
                }
            }
        }
        if (enter == -1) {
            if (trace) trace->status = 2;
            return finalize("INFEASIBLE (no entering column found)" + g_newline, T, rows, cols, basis, names,
                            "INFEASIBLE", false);
        }
        pivot(T.data(), rows, cols, leave, enter);
        basis[leave] = enter;
        if (trace) {
            trace->enter.push_back(enter);
            trace->leave.push_back(leave);
            if (trace->keep_history) trace->history.push_back(T);
        }
        if (want_text) {
            std::string it;
            append_tableau(it, title, T, rows, cols, basis, names, iter);
            Mask mk = cross_mask(rows, cols, leave, enter);
            if (sink) sink(it, mk);
        }
        iter++;
    }
}

// ---- LPSolver ------------------------------------------------------------------------------------

// LPSolver.NormalizeAlgorithmKey (LPSolver.cs:61-76)
std::string normalize_algorithm_key(const std::string& algorithm) {
    bool blank = true;
    for (char ch : algorithm)
        if (!std::isspace((unsigned char)ch)) blank = false;
    if (blank) throw SolveError(ERR_UNSUPPORTED_ALGO, "No algorithm selected.");
    std::string key;
    for (char ch : algorithm) key.push_back((char)std::tolower((unsigned char)ch));
    auto trim_ws = [](std::string s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) a++;
        while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    };
    key = trim_ws(key);
    {   // string.Replace: one left-to-right pass over non-overlapping matches
        std::string out;
        size_t from = 0, p;
        while ((p = key.find("algorithm", from)) != std::string::npos) {
            out.append(key, from, p - from);
            from = p + 9;
        }
        out.append(key, from, std::string::npos);
        key = out;
    }
    std::string collapsed;
    bool in_ws = false;
    for (char ch : key) {
        if (std::isspace((unsigned char)ch)) {
            in_ws = true;
        } else {
            if (in_ws) collapsed.push_back(' ');
            in_ws = false;
            collapsed.push_back(ch);
        }
    }
    // a leading whitespace run collapses to one space which Trim() then removes
    return trim_ws(collapsed);
}

Outcome lp_solver_solve(const Problem& p, const std::string& algorithm, const Sink& sink, Trace* trace,
                        bool format_every_iteration) {
    std::string key = normalize_algorithm_key(algorithm);
    if (key == "primal simplex" || key == "primal") {
        PrimalOptions o;
        o.format_every_iteration = format_every_iteration;
        return primal_simplex(p, sink, trace, o);
    }
    if (key == "dual simplex" || key == "dual") return dual_simplex(p, sink, trace, format_every_iteration);
    if (key == "branch and bound simplex" || key == "branch and bound" || key == "bnb")
        return branch_and_bound(p, sink, nullptr, format_every_iteration);
    if (key == "revised primal simplex" || key == "revised primal") return revised_primal_simplex(p, sink, nullptr);
    throw SolveError(ERR_UNSUPPORTED_ALGO,
                     "Algorithm not supported: '" + algorithm +
                         "'. Try one of: Primal Simplex, Revised Primal Simplex, Dual Simplex, Branch and Bound "
                         "Simplex.");
}

}  // namespace orc
