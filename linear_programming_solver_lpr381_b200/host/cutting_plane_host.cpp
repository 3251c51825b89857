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

}  // namespace lpr381
