// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (no tests or fixtures upstream); pinned by the known-answer cases in tests/golden/.
//
// CuttingPlane.Solve / GenerateGomoryCut, following R/Models/CuttingPlane.cs:13-164 statement by
// statement, including its quirk: the source row of the cut is `tableau[i + 1, *]` ("+1 because
// row 0 is objective", CuttingPlane.cs:113) although PrimalSimplex keeps the objective in the LAST
// row — so the cut is read from the row below the fractional variable's (the z-row when that
// variable sits in the last constraint row).
#include <cmath>

#include "orc_dotnet.hpp"
#include "orc_solvers.hpp"

namespace orc {

static const char* rel_name(int rel) { return rel == LE ? "LE" : rel == GE ? "GE" : "EQ"; }

// a_j != 0 ? "{a_j:F3}x{j+1}" : null, joined by " + " (CuttingPlane.cs:28,132)
static std::string nonzero_terms(const std::vector<double>& a) {
    std::string s;
    bool first = true;
    for (size_t j = 0; j < a.size(); j++) {
        if (a[j] == 0) continue;
        if (!first) s += " + ";
        first = false;
        s += fmt_fixed(a[j], 3) + "x" + std::to_string(j + 1);
    }
    return s;
}

Outcome cutting_plane(const Problem& problem, const Sink& sink, CutTrace* trace) {
    const double Eps = 1e-9;
    const std::string& nl = g_newline;
    const int n = problem.nvars();
    Problem model = problem;
    std::string report;
    int iteration = 1;
    const int maxIterations = 50;

    report += "=== Gomory Cutting Plane Algorithm ===" + nl;
    report += "Objective: Maximize ";
    for (int j = 0; j < n; j++) report += (j ? " + " : "") + fmt_fixed(problem.c[j], 3) + "x" + std::to_string(j + 1);
    report += nl + "Subject to:" + nl;
    for (const Row& r : problem.rows) report += nonzero_terms(r.a) + " " + rel_name(r.rel) + " " + fmt_fixed(r.b, 3) + nl;
    report += "x_j >= 0, integer" + nl;

    while (iteration <= maxIterations) {
        report += "\n--- Iteration " + std::to_string(iteration) + " ---" + nl;
        Outcome lp;
        Trace lpt;
        try {
            lp = primal_simplex(model, sink, &lpt);
        } catch (const SolveError& e) {
            report += std::string("Error in PrimalSimplex: ") + e.what() + nl;
            if (trace) trace->end = CUT_LP_ERROR;
            Outcome o;
            o.report = report;
            o.summary = std::string("Error: ") + e.what();
            return o;
        }
        if (trace) {
            trace->lp_pivots.push_back((int)lpt.enter.size());
            trace->lp_status.push_back(lpt.status);
        }
        report += lp.report + nl;
        // PrimalSimplex always fills Tableau/Basis/Solution/VarNames, and Take(NumVars) of a
        // NumVars-long Solution cannot be shorter: CuttingPlane.cs:55-75 never fire.
        std::vector<double> solution(lp.x.begin(), lp.x.begin() + n);
        report += "Current solution: x* = [";
        for (int j = 0; j < n; j++) report += (j ? ", " : "") + fmt_fixed(solution[j], 3);
        report += "], z* = " + fmt_fixed(lp.z, 3) + nl;

        int fracIndex = -1;
        for (int i = 0; i < n; i++) {
            double value = solution[i];
            double frac = value - std::floor(value);
            if (frac > Eps && frac < 1 - Eps) {
                fracIndex = i;
                break;
            }
        }
        if (fracIndex == -1) {
            report += "All variables integer. Optimal integer solution found." + nl;
            Outcome o = lp;
            o.report = report;
            o.summary = "Status: OPTIMAL INTEGER\nz* = " + fmt_fixed(lp.z, 2) + "\nx* = [";
            for (int j = 0; j < n; j++) o.summary += (j ? ", " : "") + fmt_fixed(solution[j], 2);
            o.summary += "]";
            o.x = solution;
            if (trace) trace->end = CUT_INTEGER;
            return o;
        }

        int row = -1;
        for (int i = 0; i < (int)lp.basis.size(); i++)
            if (lp.basis[i] == fracIndex) {
                row = i + 1;
                break;
            }
        if (row == -1) {  // unreachable for a basic fractional value; kept for the text
            report += "Error: Variable x" + std::to_string(fracIndex + 1) + " is not basic." + nl;
            if (trace) trace->end = CUT_NONBASIC;
            Outcome o;
            o.report = report;
            o.summary = "Error: Non-basic fractional variable";
            return o;
        }

        // GenerateGomoryCut (CuttingPlane.cs:142-163)
        Row cut;
        cut.a.assign(n, 0.0);
        cut.rel = LE;
        const double* trow = lp.T.data() + (size_t)row * lp.cols;
        double rhs = trow[lp.cols - 1];
        double f0 = rhs - std::floor(rhs);
        for (int j = 0; j < n; j++) {
            double aij = trow[j];
            double fj = aij - std::floor(aij);
            if (fj > Eps) cut.a[j] = fj;
        }
        cut.b = f0;
        model.rows.push_back(cut);
        if (trace) {
            trace->frac_var.push_back(fracIndex);
            trace->cut_row.push_back(row);
            trace->cut_a.push_back(cut.a);
            trace->cut_b.push_back(cut.b);
        }
        report += "Added Gomory cut: " + nonzero_terms(cut.a) + " <= " + fmt_fixed(cut.b, 3) + nl;
        iteration++;
    }

    report += "Iteration limit reached. Stopping." + nl;
    if (trace) trace->end = CUT_INCOMPLETE;
    Outcome o;
    o.report = report;
    o.summary = "Status: INCOMPLETE";
    return o;
}

}  // namespace orc
