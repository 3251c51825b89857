// Host-side CuttingPlane: the reference's Gomory driver (R/Models/CuttingPlane.cs:13-164) over the
// GPU primal solver.  Every round is one PrimalSimplex::Solve (liblpx.so does the pivots) on the
// model grown by the previous rounds' cuts; what stays here is the cut bookkeeping — a handful of
// floor() calls on one row of the returned tableau — and the report text.
//
// Reference behaviour kept as is:
//   * the cut is read from Tableau[i + 1, *] where i is the basis position of the first fractional
//     x_j ("+1 because row 0 is objective", CuttingPlane.cs:113) although PrimalSimplex stores the
//     objective row last, so the source row is the one BELOW the fractional variable's;
//   * the cut is  sum_{j < NumVars} frac(a_j) x_j <= frac(rhs)  over the decision variables only;
//   * 50 rounds, then "Status: INCOMPLETE".
#include <cmath>

#include <algorithm>
#include <cstdlib>

#include "../../include/lpx.h"
#include "dotnet_text.hpp"
#include "lp_model.hpp"

namespace lpr381 {

using text::fixed;

static const char* rel_name(Rel r) { return r == Rel::LE ? "LE" : r == Rel::GE ? "GE" : "EQ"; }

// a != 0 ? $"{a:F3}x{j+1}" : null, non-null entries joined with " + "
static std::string nonzero_terms(const std::vector<double>& a) {
    std::string s;
    for (size_t j = 0; j < a.size(); j++) {
        if (a[j] == 0) continue;
        if (!s.empty()) s += " + ";
        s += fixed(a[j], 3) + "x" + std::to_string(j + 1);
    }
    return s;
}

static std::string bracketed(const std::vector<double>& x, int decimals) {
    std::string s = "[";
    for (size_t j = 0; j < x.size(); j++) s += (j ? ", " : "") + fixed(x[j], decimals);
    return s + "]";
}

// GenerateGomoryCut (CuttingPlane.cs:142-163)
static Constraint GenerateGomoryCut(const Matrix& tableau, int row, int numVars) {
    const double Eps = 1e-9;
    Constraint cut;
    cut.A.assign(numVars, 0.0);
    cut.Relation = Rel::LE;
    const double rhs = tableau.at(row, tableau.cols - 1);
    for (int j = 0; j < numVars; j++) {
        const double aij = tableau.at(row, j);
        const double fj = aij - std::floor(aij);
        if (fj > Eps) cut.A[j] = fj;
    }
    cut.B = rhs - std::floor(rhs);
    return cut;
}

SimplexResult CuttingPlane::Solve(const LPProblem& problem, UpdatePivot updatePivot) {
    const double Eps = 1e-9;
    const std::string& nl = NewLine();
    const int n = problem.NumVars();
    PrimalSimplex simplex;
    LPProblem model = problem.Clone();
    Cuts.clear();
    std::string report = "=== Gomory Cutting Plane Algorithm ===" + nl + "Objective: Maximize ";
    for (int j = 0; j < n; j++) report += (j ? " + " : "") + fixed(problem.C[j], 3) + "x" + std::to_string(j + 1);
    report += nl + "Subject to:" + nl;
    for (const Constraint& c : problem.Constraints)
        report += nonzero_terms(c.A) + " " + rel_name(c.Relation) + " " + fixed(c.B, 3) + nl;
    report += "x_j >= 0, integer" + nl;

    auto failed = [&](const std::string& summary) {
        SimplexResult r;
        r.Report = report;
        r.Summary = summary;
        return r;
    };

    for (int iteration = 1; iteration <= 50; iteration++) {
        report += "\n--- Iteration " + std::to_string(iteration) + " ---" + nl;
        SimplexResult lp;
        try {
            lp = simplex.Solve(model, updatePivot);
        } catch (const LpException& e) {
            if (e.code <= LPX_E_BAD_ARGS) throw;  // engine / ABI failure (LPX_E_*), not one of the reference's exceptions
            report += std::string("Error in PrimalSimplex: ") + e.what() + nl;
            return failed(std::string("Error: ") + e.what());
        }
        report += lp.Report + nl;
        if (lp.Tableau.is_null() || !lp.HasSolution) {
            report += "Error: Invalid Simplex result." + nl;
            return failed("Error: Invalid Simplex result");
        }
        std::vector<double> solution(lp.Solution.begin(), lp.Solution.begin() + std::min<size_t>(n, lp.Solution.size()));
        if ((int)solution.size() != n) {
            report += "Error: Solution length (" + std::to_string(solution.size()) + ") does not match NumVars (" +
                      std::to_string(n) + ")." + nl;
            return failed("Error: Invalid solution length");
        }
        report += "Current solution: x* = " + bracketed(solution, 3) + ", z* = " + fixed(lp.OptimalValue, 3) + nl;

        int fracIndex = -1;
        for (int i = 0; i < n; i++) {
            const double frac = solution[i] - std::floor(solution[i]);
            if (frac > Eps && frac < 1 - Eps) {
                fracIndex = i;
                break;
            }
        }
        if (fracIndex == -1) {
            report += "All variables integer. Optimal integer solution found." + nl;
            SimplexResult r = lp;
            r.Report = report;
            r.Summary = "Status: OPTIMAL INTEGER\nz* = " + fixed(lp.OptimalValue, 2) + "\nx* = " + bracketed(solution, 2);
            r.Solution = solution;
            return r;
        }
        int row = -1;
        for (size_t i = 0; i < lp.Basis.size(); i++)
            if (lp.Basis[i] == fracIndex) {
                row = (int)i + 1;
                break;
            }
        if (row == -1) {
            report += "Error: Variable x" + std::to_string(fracIndex + 1) + " is not basic." + nl;
            return failed("Error: Non-basic fractional variable");
        }
        Constraint cut = GenerateGomoryCut(lp.Tableau, row, n);
        model.Constraints.push_back(cut);
        Cuts.push_back(cut);
        report += "Added Gomory cut: " + nonzero_terms(cut.A) + " <= " + fixed(cut.B, 3) + nl;
    }
    report += "Iteration limit reached. Stopping." + nl;
    return failed("Status: INCOMPLETE");
}

// ---- CuttingPlaneRevised.Solve (R/Models/CuttingPlaneRevised.cs:14-111) -------------------------
// Over the GPU revised simplex.  As upstream it reads x back from the Summary text (three-decimal
// values), bounds the first fractional variable by its floor and re-solves, at most 50 rounds; the
// LP solver's exceptions pass through.
namespace {

// ExtractSolution (:93-110); returns false when no "x* = [...]" line is found
bool extract_solution(const std::string& summary, int nVars, std::vector<double>& x) {
    size_t pos = 0;
    while (pos <= summary.size()) {
        size_t eol = summary.find('\n', pos);
        if (eol == std::string::npos) eol = summary.size();
        const std::string line = summary.substr(pos, eol - pos);
        const size_t k = line.find_first_not_of(" \t\r");
        if (k != std::string::npos && line.compare(k, 6, "x* = [") == 0) {
            const size_t sb = line.find('['), se = line.find(']');
            if (sb == std::string::npos || se == std::string::npos || se <= sb) return false;
            const std::string body = line.substr(sb + 1, se - sb - 1);
            size_t q = 0;
            while (q <= body.size()) {
                size_t comma = body.find(',', q);
                if (comma == std::string::npos) comma = body.size();
                x.push_back(std::strtod(body.substr(q, comma - q).c_str(), nullptr));
                q = comma + 1;
            }
            x.resize(nVars, 0.0);
            return true;
        }
        pos = eol + 1;
    }
    return false;
}

}  // namespace

SimplexResult CuttingPlaneRevised::Solve(const LPProblem& problem, UpdatePivot updatePivot) {
    const double Eps = 1e-9;
    const std::string& nl = NewLine();
    RevisedPrimalSimplex solver;
    LPProblem model = problem.Clone();
    Cuts.clear();
    std::string report;
    auto finish = [&](const std::string& summary) {
        SimplexResult r;
        r.Report = report;
        r.Summary = summary;
        return r;
    };
    for (int iter = 1;; ) {
        SimplexResult lp = solver.Solve(model, updatePivot);
        report += "--- Cutting-Plane Iteration " + std::to_string(iter) + " ---" + nl;
        report += lp.Report + nl;
        if (lp.Summary.find("Status: OPTIMAL") == std::string::npos) {
            report += "Stopping: LP not OPTIMAL; cannot continue cutting." + nl;
            return finish("Terminated: LP not OPTIMAL; cutting-plane stopped.");
        }
        std::vector<double> x;
        if (!extract_solution(lp.Summary, problem.NumVars(), x)) {
            report += "Stopping: Could not parse primal solution." + nl;
            return finish("Terminated: could not parse solution.");
        }
        int fracIndex = -1;
        for (int i = 0; i < (int)x.size(); i++) {
            const double frac = x[i] - std::floor(x[i]);
            if (frac > Eps && frac < 1 - Eps) {
                fracIndex = i;
                break;
            }
        }
        if (fracIndex == -1) {
            report += "All decision variables are integer. Optimal integer solution found." + nl;
            std::string summary = lp.Summary;
            for (size_t at = 0; (at = summary.find("Status: OPTIMAL", at)) != std::string::npos; at += 23)
                summary.replace(at, 15, "Status: OPTIMAL INTEGER");
            return finish(summary);
        }
        const double floorVal = std::floor(x[fracIndex] + 1e-12);
        Constraint cut;
        cut.A.assign(model.NumVars(), 0.0);
        cut.A[fracIndex] = 1.0;
        cut.Relation = Rel::LE;
        cut.B = floorVal;
        model.Constraints.push_back(cut);
        Cuts.push_back(cut);
        report += "Added cut: x" + std::to_string(fracIndex + 1) + " \xE2\x89\xA4 " + text::round_trip(floorVal) + " (current x" +
                  std::to_string(fracIndex + 1) + " = " + text::custom_hash(x[fracIndex]) + ")" + nl;
        iter++;
        if (iter > 50) {
            report += "Iteration limit reached." + nl;
            return finish("Iteration limit reached (solution may still be fractional).");
        }
    }
}

}  // namespace lpr381
