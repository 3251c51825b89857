// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (no tests or fixtures upstream); pinned by the known-answer cases in tests/golden/.
//
// CuttingPlane.Solve / GenerateGomoryCut, following R/Models/CuttingPlane.cs:13-164 statement by
// statement, including its quirk: the source row of the cut is `tableau[i + 1, *]` ("+1 because
// row 0 is objective", CuttingPlane.cs:113) although PrimalSimplex keeps the objective in the LAST
// row — so the cut is read from the row below the fractional variable's (the z-row when that
// variable sits in the last constraint row).
#include <cmath>
#include <cstdlib>

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

// ---- CuttingPlaneRevised.Solve (R/Models/CuttingPlaneRevised.cs:14-111) -------------------------
// Drives RevisedPrimalSimplex; reads the solution back from the Summary TEXT (so it sees x rounded to
// three decimals, :93-110), adds the bound  x_k <= floor(x_k + 1e-12)  for the first fractional x_k,
// at most 50 rounds.  Exceptions of the LP solver are not caught.
Outcome cutting_plane_revised(const Problem& problem, const Sink& sink, CutTrace* trace) {
    const double Eps = 1e-9;
    const std::string& nl = g_newline;
    const int nVars = problem.nvars();
    Problem model = problem;
    std::string report;
    int iter = 1;
    while (true) {
        Outcome lp = revised_primal_simplex(model, sink, nullptr);
        report += "--- Cutting-Plane Iteration " + std::to_string(iter) + " ---" + nl;
        report += lp.report + nl;
        auto finish = [&](const std::string& summary, int end) {
            if (trace) trace->end = end;
            Outcome o;
            o.report = report;
            o.summary = summary;
            return o;
        };
        // Summary.Contains("Status: OPTIMAL", OrdinalIgnoreCase): the solver writes it in this exact case
        if (lp.summary.find("Status: OPTIMAL") == std::string::npos) {
            report += "Stopping: LP not OPTIMAL; cannot continue cutting." + nl;
            return finish("Terminated: LP not OPTIMAL; cutting-plane stopped.", CUT_LP_ERROR);
        }
        // ExtractSolution: first line whose trimmed start is "x* = [", numbers between the brackets
        std::vector<double> x;
        bool parsed = false;
        size_t pos = 0;
        while (pos <= lp.summary.size()) {
            size_t eol = lp.summary.find('\n', pos);
            if (eol == std::string::npos) eol = lp.summary.size();
            std::string line = lp.summary.substr(pos, eol - pos);
            size_t k = line.find_first_not_of(" \t\r");
            if (k != std::string::npos && line.compare(k, 6, "x* = [") == 0) {
                const size_t sb = line.find('['), se = line.find(']');
                if (sb != std::string::npos && se != std::string::npos && se > sb) {
                    std::string body = line.substr(sb + 1, se - sb - 1);
                    size_t q = 0;
                    while (q <= body.size()) {
                        size_t comma = body.find(',', q);
                        if (comma == std::string::npos) comma = body.size();
                        x.push_back(std::strtod(body.substr(q, comma - q).c_str(), nullptr));
                        q = comma + 1;
                    }
                    parsed = true;
                }
                break;
            }
            pos = eol + 1;
        }
        if (!parsed) {
            report += "Stopping: Could not parse primal solution." + nl;
            return finish("Terminated: could not parse solution.", CUT_NONBASIC);
        }
        x.resize(nVars, 0.0);
        int fracIndex = -1;
        for (int i = 0; i < nVars; i++) {
            const double frac = x[i] - std::floor(x[i]);
            if (frac > Eps && frac < 1 - Eps) {
                fracIndex = i;
                break;
            }
        }
        if (fracIndex == -1) {
            report += "All decision variables are integer. Optimal integer solution found." + nl;
            std::string summary = lp.summary;
            for (size_t at = 0; (at = summary.find("Status: OPTIMAL", at)) != std::string::npos; at += 23)
                summary.replace(at, 15, "Status: OPTIMAL INTEGER");
            return finish(summary, CUT_INTEGER);
        }
        const double floorVal = std::floor(x[fracIndex] + 1e-12);
        Row cut;
        cut.a.assign(model.nvars(), 0.0);
        cut.a[fracIndex] = 1.0;
        cut.rel = LE;
        cut.b = floorVal;
        model.rows.push_back(cut);
        if (trace) {
            trace->frac_var.push_back(fracIndex);
            trace->cut_row.push_back(-1);
            trace->cut_a.push_back(cut.a);
            trace->cut_b.push_back(cut.b);
        }
        report += "Added cut: x" + std::to_string(fracIndex + 1) + " \xE2\x89\xA4 " + fmt_roundtrip(floorVal) + " (current x" +
                  std::to_string(fracIndex + 1) + " = " + fmt_custom(x[fracIndex]) + ")" + nl;
        iter++;
        if (iter > 50) {
            report += "Iteration limit reached." + nl;
            return finish("Iteration limit reached (solution may still be fractional).", CUT_INCOMPLETE);
        }
    }
}

}  // namespace orc
