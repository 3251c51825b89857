// Host-side PrimalSimplex / DualSimplex / LPSolver / LPController: the reference's interface
// (R/Models/PrimalSimplex.cs, DualSimplex.cs, LPSolver.cs, R/Controllers/LPController.cs) with the
// arithmetic delegated to liblpx.so.  What stays here is exactly what stays in C# after the
// drop-in: marshalling and the text (AppendCanonicalForm, AppendTableau, FinalizeReport).
#include <algorithm>
#include <cctype>
#include <cmath>

#include "../../include/lpx.h"
#include "dotnet_text.hpp"
#include "host_util.hpp"
#include "lp_model.hpp"

namespace lpr381 {

using text::custom_hash;
using text::pad_left;

// ---- marshalling ---------------------------------------------------------------------------------
Flat flatten(const LPProblem& p) {
    Flat f;
    f.m = (int)p.Constraints.size();
    f.n = p.NumVars();
    f.sense = (int)p.ObjectiveSense;
    f.A.resize((size_t)f.m * f.n);
    f.rel.resize(f.m);
    f.b.resize(f.m);
    for (int i = 0; i < f.m; i++) {
        const Constraint& c = p.Constraints[i];
        // BuildTableau indexes row.A[j] for j < n: a shorter row is an IndexOutOfRangeException
        if ((int)c.A.size() < f.n) throw LpException("Index was outside the bounds of the array.");
        std::copy(c.A.begin(), c.A.begin() + f.n, f.A.begin() + (size_t)i * f.n);
        f.rel[i] = (int)c.Relation;
        f.b[i] = c.B;
    }
    f.c = p.C;
    return f;
}

void throw_on(int rc) {
    if (rc != LPX_OK) throw LpException(lpx_last_error(), rc);
}

// ---- text helpers --------------------------------------------------------------------------------
static std::string signed_terms(const std::vector<double>& v) {
    std::string s;
    for (size_t j = 0; j < v.size(); j++) {
        if (j) s += ' ';
        s += v[j] >= 0 ? '+' : '-';
        s += custom_hash(std::fabs(v[j])) + "x" + std::to_string(j + 1);
    }
    return s;
}

// PrimalSimplex.AppendCanonicalForm (PrimalSimplex.cs:259-270) on the EQ-expanded model
static std::string canonical_form(const LPProblem& original) {
    const std::string& nl = NewLine();
    std::vector<double> c = original.C;
    if (original.ObjectiveSense == Sense::Min)
        for (double& v : c) v = -v;
    std::string sb = "Objective: max " + signed_terms(c) + nl + "Subject to:" + nl;
    for (const Constraint& k : original.Constraints) {
        if (k.Relation == Rel::EQ) {
            sb += "  " + signed_terms(k.A) + " <= " + custom_hash(k.B) + nl;
            std::vector<double> neg = k.A;
            for (double& v : neg) v *= -1;
            sb += "  " + signed_terms(neg) + " <= " + custom_hash(k.B * -1) + nl;
        } else {
            sb += "  " + signed_terms(k.A) + (k.Relation == Rel::LE ? " <= " : " >= ") + custom_hash(k.B) + nl;
        }
    }
    return sb + "x >= 0" + nl;
}

std::vector<std::string> var_names(int n, int m) {
    std::vector<std::string> names;
    for (int j = 0; j < n; j++) names.push_back("x" + std::to_string(j + 1));
    for (int j = 0; j < m; j++) names.push_back("c" + std::to_string(j + 1));
    return names;
}

// AppendTableau (PrimalSimplex.cs:272-304, DualSimplex.cs:248-281)
std::string tableau_text(const char* title, const double* T, int rows, int cols, const std::vector<int>& basis,
                         const std::vector<std::string>& names, int iter) {
    const std::string& nl = NewLine();
    const int m = rows - 1, ns = cols - 1, W = 12;
    std::string sb = std::string(title) + " " + std::to_string(iter) + nl;
    sb += pad_left("Basis", W);
    for (int j = 0; j < ns; j++) sb += pad_left(names[j], W);
    sb += pad_left("RHS", W) + nl;
    sb += std::string((size_t)W * (ns + 2), '-') + nl;
    auto row = [&](const std::string& label, const double* r) {
        sb += pad_left(label, W);
        for (int j = 0; j <= ns; j++) sb += pad_left(custom_hash(r[j]), W);
        sb += nl;
    };
    row("z", T + (size_t)m * cols);
    for (int i = 0; i < m; i++) row(names[basis[i]], T + (size_t)i * cols);
    return sb;
}

Highlight cross(int rows, int cols, int row, int col) {
    Highlight h;
    h.rows = rows;
    h.cols = cols;
    h.v.assign((size_t)rows * cols, 0);
    for (int j = 0; j < cols; j++) h.v[(size_t)row * cols + j] = 1;
    for (int i = 0; i < rows; i++) h.v[(size_t)i * cols + col] = 1;
    return h;
}

// FinalizeReport (PrimalSimplex.cs:130-159).  full = false is DualSimplex's variant, which returns
// Report and Summary only (DualSimplex.cs:310).
static SimplexResult finalize_report(std::string sb, const std::vector<double>& T, int rows, int cols,
                                     const std::vector<int>& basis, const std::vector<double>& x, double z,
                                     const std::vector<std::string>& names, const char* status, bool full) {
    const std::string& nl = NewLine();
    const int n = (int)x.size();
    sb += "\nStatus: " + std::string(status) + nl;
    for (int j = 0; j < n; j++)
        sb += "  x" + std::to_string(j + 1) + " = " + custom_hash(text::math_round(x[j], 3)) + nl;
    sb += "  z* = " + custom_hash(text::math_round(z, 3)) + nl;
    std::string summary = "Status: " + std::string(status) + nl + "z* = " + custom_hash(text::math_round(z, 3)) + nl;
    summary += "x* = [";
    for (int j = 0; j < n; j++) summary += (j ? ", " : "") + text::round_trip(text::math_round(x[j], 3));
    summary += "]" + nl;
    SimplexResult r;
    r.Report = sb;
    r.Summary = summary;
    if (full) {
        r.OptimalValue = z;
        r.HasSolution = true;
        r.Solution = x;
        r.Tableau.rows = rows;
        r.Tableau.cols = cols;
        r.Tableau.v = T;
        r.Basis = basis;
        r.VarNames = names;
    }
    return r;
}

// Replays the iteration text from the engine's per-iteration tableaux and pivot list.
static void emit_iterations(const UpdatePivot& cb, const char* title, const std::vector<double>& hist, int n_hist,
                            const std::vector<int>& pivots, int first_pivot, int rows, int cols, int n,
                            const std::vector<std::string>& names, std::vector<int>& basis) {
    const size_t tsize = (size_t)rows * cols;
    for (int k = 0; k < n_hist; k++) {
        Highlight hl;
        if (k > 0) {
            const int e = pivots[2 * (first_pivot + k - 1)], l = pivots[2 * (first_pivot + k - 1) + 1];
            basis[l] = e;
            hl = cross(rows, cols, l, e);
        }
        cb(tableau_text(title, hist.data() + (size_t)k * tsize, rows, cols, basis, names, k), hl);
    }
    (void)n;
}

// ---- PrimalSimplex.Solve (PrimalSimplex.cs:57-127) --------------------------------------------
SimplexResult PrimalSimplex::Solve(const LPProblem& original, UpdatePivot updatePivot) {
    Flat f = flatten(original);
    int rows = 0, cols = 0;
    throw_on(lpx_tableau_dims(f.m, f.n, f.rel.data(), &rows, &cols));
    lpx_options opt;
    lpx_default_options(&opt);
    const size_t tsize = (size_t)rows * cols;
    std::vector<int> pivots((size_t)opt.max_iterations * 2), basis(rows - 1);
    std::vector<double> x(f.n), T(tsize), hist;
    double z = 0;
    int status = 0, npiv = 0;
    int hist_cap = 0;
    if (updatePivot) {
        // the engine returns every iteration's tableau in one call; a first pass sizes the buffer
        throw_on(lpx_primal_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status,
                                  &npiv, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0));
        if (status >= 0 || status == LPX_S_ITER_LIMIT) {
            hist_cap = npiv + 1;
            hist.resize(tsize * hist_cap);
        }
    }
    throw_on(lpx_primal_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status, &npiv,
                              pivots.data(), opt.max_iterations, basis.data(), x.data(), &z, T.data(),
                              hist_cap ? hist.data() : nullptr, hist_cap));
    if (status == LPX_S_GE_ROW || status == LPX_S_NEG_RHS) throw LpException(lpx_status_message(status), status);

    const std::vector<std::string> names = var_names(f.n, rows - 1);
    if (updatePivot) {
        std::vector<int> bs(rows - 1);
        for (int i = 0; i < rows - 1; i++) bs[i] = f.n + i;
        emit_iterations(updatePivot, "TABLEAU Iteration", hist, std::min(hist_cap, npiv + 1), pivots, 0, rows, cols, f.n,
                        names, bs);
    }
    if (status == LPX_S_ITER_LIMIT) throw LpException(lpx_status_message(status), status);
    std::string report = canonical_form(original);
    if (status == LPX_UNBOUNDED) report += "UNBOUNDED" + NewLine();
    return finalize_report(report, T, rows, cols, basis, x, z, names, status == LPX_UNBOUNDED ? "UNBOUNDED" : "OPTIMAL",
                           true);
}

// ---- DualSimplex.Solve (DualSimplex.cs:15-114) -----------------------------------------------
SimplexResult DualSimplex::Solve(const LPProblem& original, UpdatePivot updatePivot) {
    Flat f = flatten(original);
    int rows = 0, cols = 0;
    throw_on(lpx_tableau_dims(f.m, f.n, f.rel.data(), &rows, &cols));
    lpx_options opt;
    lpx_default_options(&opt);
    const size_t tsize = (size_t)rows * cols;
    const int cap = opt.max_iterations + 128;
    std::vector<int> pivots((size_t)cap * 2), basis(rows - 1);
    std::vector<double> x(f.n), T(tsize), hist;
    double z = 0;
    int status = 0, npiv = 0, silent = 0, hist_cap = 0;
    if (updatePivot) {
        throw_on(lpx_dual_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status, &npiv,
                                &silent, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0));
        hist_cap = npiv - silent + 1;
        hist.resize(tsize * hist_cap);
    }
    throw_on(lpx_dual_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status, &npiv,
                            &silent, pivots.data(), cap, basis.data(), x.data(), &z, T.data(),
                            hist_cap ? hist.data() : nullptr, hist_cap));
    const std::vector<std::string> names = var_names(f.n, rows - 1);
    const char* title = "DUAL SIMPLEX TABLEAU Iteration";
    if (updatePivot) {
        // basis after the silent ForceDualFeasibility pivots (DualSimplex.cs:24)
        std::vector<int> bs(rows - 1);
        for (int i = 0; i < rows - 1; i++) bs[i] = f.n + i;
        for (int k = 0; k < silent; k++) bs[pivots[2 * k + 1]] = pivots[2 * k];
        emit_iterations(updatePivot, title, hist, hist_cap, pivots, silent, rows, cols, f.n, names, bs);
        if (status == LPX_OPTIMAL) {
            // the optimal tableau is printed once more, "z row" (row 0) highlighted (DualSimplex.cs:58-72)
            Highlight hl;
            hl.rows = rows;
            hl.cols = cols;
            hl.v.assign(tsize, 0);
            for (int j = 0; j < cols; j++) hl.v[j] = 1;
            updatePivot(tableau_text(title, T.data(), rows, cols, basis, names, npiv - silent + 1), hl);
        }
    }
    if (status == LPX_S_ITER_LIMIT) throw LpException("Iteration limit exceeded (Dual Simplex).", status);
    std::string head = status == LPX_INFEASIBLE ? "INFEASIBLE (no entering column found)" + NewLine() : "";
    return finalize_report(head, T, rows, cols, basis, x, z, names, status == LPX_INFEASIBLE ? "INFEASIBLE" : "OPTIMAL",
                           false);
}

// ---- LPSolver (LPSolver.cs:16-76) ------------------------------------------------------------
std::string LPSolver::NormalizeAlgorithmKey(const std::string& algorithm) {
    bool blank = true;
    for (char ch : algorithm) blank = blank && std::isspace((unsigned char)ch);
    if (blank) throw LpException("No algorithm selected.");
    std::string lower;
    for (char ch : algorithm) lower += (char)std::tolower((unsigned char)ch);
    std::string key;
    for (size_t i = 0; i < lower.size();) {  // Replace("algorithm", "")
        if (lower.compare(i, 9, "algorithm") == 0) i += 9;
        else key += lower[i++];
    }
    std::string out;  // Regex.Replace(key, @"\s+", " ").Trim(), after the initial Trim()
    bool pending_space = false;
    for (char ch : key) {
        if (std::isspace((unsigned char)ch)) pending_space = !out.empty();
        else {
            if (pending_space) out += ' ';
            pending_space = false;
            out += ch;
        }
    }
    return out;
}

SimplexResult LPSolver::Solve(const LPProblem& problem, const std::string& algorithm, UpdatePivot updatePivot) {
    const std::string key = NormalizeAlgorithmKey(algorithm);
    std::unique_ptr<ILPAlgorithm> algo;
    if (key == "primal simplex" || key == "primal") algo.reset(new PrimalSimplex());
    else if (key == "revised primal simplex" || key == "revised primal") algo.reset(new RevisedPrimalSimplex());
    else if (key == "dual simplex" || key == "dual") algo.reset(new DualSimplex());
    else if (key == "branch and bound simplex" || key == "branch and bound" || key == "bnb") algo.reset(new BranchAndBound());
    else
        throw LpException("Algorithm not supported: '" + algorithm +
                          "'. Try one of: Primal Simplex, Revised Primal Simplex, Dual Simplex, Branch and Bound "
                          "Simplex.");
    SimplexResult result = algo->Solve(problem, updatePivot);
    FinalTableau = result.Tableau;
    return result;
}

SimplexResult LPController::SolvePrimalSimplex(const LPProblem& problem) {
    LPSolver solver;
    return solver.Solve(problem, "Primal Simplex");
}

}  // namespace lpr381
