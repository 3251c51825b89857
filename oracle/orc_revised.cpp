// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (no tests or fixtures upstream); pinned by the known-answer cases in tests/golden/.
//
// RevisedPrimalSimplex.Solve, following R/Models/RevisedPrimalSimplex.cs:17-145 statement by
// statement: price-out with the basis inverse RECOMPUTED by Gauss-Jordan with partial pivoting
// every iteration (:121, :409-456), sequential dot products (:331-394), ratio margin 1e-12 (:104),
// and the text of BuildIterationBlock / BuildFinalSummary (:190-290).  Quirks kept: Standardize
// negates C when the sense is MAX (:153-154) and the loop then picks the most negative reduced cost;
// the iteration block printed after a pivot mixes the OLD reduced costs / direction / theta with the
// NEW basis, inverse and x_B (so its "ratio test" lines divide the new x_B by the old d); the result
// carries Report and Summary only, and z* is recomputed from the original C.
#include <cmath>

#include "orc_dotnet.hpp"
#include "orc_solvers.hpp"

namespace orc {

namespace {

const double kEps = 1e-9;

struct Mat {
    int r = 0, c = 0;
    std::vector<double> v;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), v((size_t)r_ * c_, 0.0) {}
    double& at(int i, int j) { return v[(size_t)i * c + j]; }
    double at(int i, int j) const { return v[(size_t)i * c + j]; }
};

Mat submatrix(const Mat& A, const std::vector<int>& cols) {
    Mat M(A.r, (int)cols.size());
    for (int i = 0; i < A.r; i++)
        for (int j = 0; j < (int)cols.size(); j++) M.at(i, j) = A.at(i, cols[j]);
    return M;
}

// Invert (:409-456): Gauss-Jordan on [M | I], partial pivoting with the FIRST largest |entry|.
Mat invert(const Mat& M) {
    const int n = M.r;
    if (M.r != M.c) throw SolveError(ERR_SINGULAR, "Matrix not square; cannot invert B.");
    Mat A(n, 2 * n);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) A.at(i, j) = M.at(i, j);
        A.at(i, n + i) = 1.0;
    }
    for (int col = 0; col < n; col++) {
        int pivotRow = col;
        double best = std::fabs(A.at(pivotRow, col));
        for (int r = col + 1; r < n; r++) {
            const double v = std::fabs(A.at(r, col));
            if (v > best) {
                best = v;
                pivotRow = r;
            }
        }
        if (std::fabs(A.at(pivotRow, col)) < kEps) throw SolveError(ERR_SINGULAR, "Singular basis encountered.");
        if (pivotRow != col)
            for (int j = 0; j < 2 * n; j++) std::swap(A.at(col, j), A.at(pivotRow, j));
        const double piv = A.at(col, col);
        for (int j = 0; j < 2 * n; j++) A.at(col, j) /= piv;
        for (int r = 0; r < n; r++) {
            if (r == col) continue;
            const double factor = A.at(r, col);
            for (int j = 0; j < 2 * n; j++) A.at(r, j) -= factor * A.at(col, j);
        }
    }
    Mat inv(n, n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv.at(i, j) = A.at(i, n + j);
    return inv;
}

std::vector<double> mat_vec(const Mat& A, const std::vector<double>& x) {  // Multiply(double[,], double[])
    std::vector<double> y(A.r);
    for (int i = 0; i < A.r; i++) {
        double s = 0;
        for (int j = 0; j < A.c; j++) s += A.at(i, j) * x[j];
        y[i] = s;
    }
    return y;
}

std::vector<double> row_mat(const std::vector<double>& row, const Mat& A) {  // MultiplyRow
    std::vector<double> y(A.c);
    for (int j = 0; j < A.c; j++) {
        double s = 0;
        for (int i = 0; i < A.r; i++) s += row[i] * A.at(i, j);
        y[j] = s;
    }
    return y;
}

double dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}

std::vector<double> subvector(const std::vector<double>& v, const std::vector<int>& idx) {
    std::vector<double> r(idx.size());
    for (size_t i = 0; i < idx.size(); i++) r[i] = v[idx[i]];
    return r;
}

// BuildIterationBlock (:190-247)
std::string iteration_block(int iter, const std::vector<int>& Bidx, const std::vector<int>& Nidx,
                            const std::vector<std::string>& names, const Mat& Binv, const std::vector<double>& xB,
                            double z, const std::vector<double>* rN, int entering, const std::vector<double>* d,
                            const double* bestTheta) {
    const std::string& nl = g_newline;
    auto join_names = [&](const std::vector<int>& idx) {
        std::string s;
        for (size_t k = 0; k < idx.size(); k++) s += (k ? ", " : "") + names[idx[k]];
        return s;
    };
    auto join_vals = [&](const std::vector<double>& v) {
        std::string s;
        for (size_t k = 0; k < v.size(); k++) s += (k ? ", " : "") + fmt_custom(v[k]);
        return s;
    };
    std::string sb = "=== Revised Simplex Iteration " + std::to_string(iter) + " ===" + nl;
    sb += "Basis: " + join_names(Bidx) + nl;
    sb += "Nonbasic: " + join_names(Nidx) + nl;
    sb += "\nProduct-form: current B^{-1}" + nl;
    for (int i = 0; i < Binv.r; i++) {  // MatrixToString (:249-261)
        for (int j = 0; j < Binv.c; j++) sb += pad_left(fmt_custom(Binv.at(i, j)), 12);
        sb += nl;
    }
    sb += "x_B = [" + join_vals(xB) + "]" + nl;
    sb += "z = " + fmt_custom(z) + nl;
    if (rN) {
        sb += "\nReduced costs (r_N = c_N - c_B^T B^{-1} N):" + nl;
        for (size_t j = 0; j < rN->size(); j++)
            sb += "  r(" + std::to_string(Nidx[j]) + ":" + names[Nidx[j]] + ") = " + fmt_custom((*rN)[j]) + nl;
    }
    if (entering >= 0) sb += "\nEntering variable: " + names[entering] + nl;
    if (d) {
        sb += "Direction d = B^{-1} * a_entering:" + nl;
        sb += "  d = [" + join_vals(*d) + "]" + nl;
        sb += "\nRatio test (theta):" + nl;
        for (size_t i = 0; i < d->size(); i++) {
            if ((*d)[i] > kEps)
                sb += "  row " + std::to_string(i + 1) + ": " + fmt_custom(xB[i]) + " / " + fmt_custom((*d)[i]) + " = " +
                      fmt_custom(xB[i] / (*d)[i]) + nl;
            else
                sb += "  row " + std::to_string(i + 1) + ": d_i <= 0 (skip)" + nl;
        }
        if (bestTheta) sb += "Chosen theta* = " + fmt_custom(*bestTheta) + nl;
    }
    sb += nl;
    return sb;
}

// BuildFinalSummary (:264-297)
Outcome final_summary(const Problem& original, const std::vector<int>& Bidx, int nVars, const std::vector<double>& xB,
                      const char* status) {
    const std::string& nl = g_newline;
    std::vector<double> x(nVars, 0.0);
    for (size_t i = 0; i < Bidx.size(); i++)
        if (Bidx[i] < nVars) x[Bidx[i]] = xB[i];
    std::string sb = "\nStatus: " + std::string(status) + nl;
    for (int j = 0; j < nVars; j++) sb += "  x" + std::to_string(j + 1) + " = " + fmt_custom(math_round(x[j], 3)) + nl;
    std::string summary = "Status: " + std::string(status) + nl + "x* = [";
    for (int j = 0; j < nVars; j++) summary += (j ? ", " : "") + fmt_roundtrip(math_round(x[j], 3));
    summary += "]" + nl;
    double zOriginal = 0;
    for (int j = 0; j < nVars; j++) zOriginal += original.c[j] * x[j];
    sb += "  z* = " + fmt_custom(math_round(zOriginal, 3)) + nl;
    summary += "z* = " + fmt_custom(math_round(zOriginal, 3)) + nl;
    Outcome o;
    o.report = sb;
    o.summary = summary;
    o.x = x;  // kept for the numeric comparison in tests (the reference returns Report / Summary only)
    o.z = zOriginal;
    return o;
}

}  // namespace

Outcome revised_primal_simplex(const Problem& original, const Sink& sink, RevTrace* trace, int max_iterations) {
    for (const Row& r : original.rows)
        if (!(r.rel == LE && r.b >= -1e-9))
            throw SolveError(ERR_REV_UNSUPPORTED,
                             "Revised Primal Simplex currently supports only <= constraints with RHS >= 0. Use Dual "
                             "Simplex for models with >= or =.");
    // Standardize (:148-186): with all rows <= and b >= -1e-9 only the objective flip remains
    Problem model = original;
    if (model.sense == MAX)
        for (double& v : model.c) v = -v;

    const int m = (int)model.rows.size(), n = model.nvars(), Ntot = n + m;
    Mat A(m, Ntot);
    std::vector<double> c(Ntot, 0.0), b(m);
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) A.at(i, j) = model.rows[i].a[j];
        A.at(i, n + i) = 1.0;
        b[i] = model.rows[i].b;
    }
    for (int j = 0; j < n; j++) c[j] = model.c[j];
    std::vector<int> Bidx(m), Nidx(n);
    for (int i = 0; i < m; i++) Bidx[i] = n + i;
    for (int j = 0; j < n; j++) Nidx[j] = j;
    std::vector<std::string> names(Ntot);
    for (int j = 0; j < n; j++) names[j] = "x" + std::to_string(j + 1);
    for (int j = 0; j < m; j++) names[n + j] = "c" + std::to_string(j + 1);

    Mat Binv = invert(submatrix(A, Bidx));
    std::vector<double> xB = mat_vec(Binv, b);
    std::vector<double> cB = subvector(c, Bidx);
    double z = dot(cB, xB);
    if (sink) sink(iteration_block(0, Bidx, Nidx, names, Binv, xB, z, nullptr, -1, nullptr, nullptr), Mask{});
    auto snapshot = [&]() {
        if (!trace) return;
        trace->basis = Bidx;
        trace->xB = xB;
        trace->Binv = Binv.v;
        trace->z = z;
    };
    snapshot();

    for (int iter = 1; iter <= max_iterations; iter++) {
        Mat Nmat = submatrix(A, Nidx);
        std::vector<double> cN = subvector(c, Nidx);
        std::vector<double> piT = row_mat(cB, Binv);
        std::vector<double> piN = row_mat(piT, Nmat);
        std::vector<double> rN(cN.size());
        for (size_t j = 0; j < cN.size(); j++) rN[j] = cN[j] - piN[j];

        int enteringPos = -1;
        double minRC = -kEps;
        for (size_t j = 0; j < rN.size(); j++)
            if (rN[j] < minRC) {
                minRC = rN[j];
                enteringPos = (int)j;
            }
        if (enteringPos == -1) {
            if (trace) trace->status = 0;
            return final_summary(original, Bidx, n, xB, "OPTIMAL");
        }
        const int entering = Nidx[enteringPos];
        std::vector<double> aE(m);
        for (int i = 0; i < m; i++) aE[i] = A.at(i, entering);
        std::vector<double> d = mat_vec(Binv, aE);

        int leaveRow = -1;
        double bestTheta = HUGE_VAL;
        for (int i = 0; i < m; i++)
            if (d[i] > kEps) {
                const double theta = xB[i] / d[i];
                if (theta < bestTheta - 1e-12) {
                    bestTheta = theta;
                    leaveRow = i;
                }
            }
        if (leaveRow == -1) {
            if (trace) trace->status = 1;
            return final_summary(original, Bidx, n, xB, "UNBOUNDED");
        }
        const int leaving = Bidx[leaveRow];
        Bidx[leaveRow] = entering;
        Nidx.erase(Nidx.begin() + enteringPos);
        Nidx.push_back(leaving);

        Binv = invert(submatrix(A, Bidx));
        cB = subvector(c, Bidx);
        xB = mat_vec(Binv, b);
        z = dot(cB, xB);
        if (trace) {
            trace->enter.push_back(entering);
            trace->leave.push_back(leaveRow);
            trace->theta.push_back(bestTheta);
        }
        snapshot();
        if (sink) {
            Mask mk;
            mk.rows = m;
            mk.cols = 4;
            mk.bits.assign((size_t)m * 4, 0);
            for (int j = 0; j < 4; j++) mk.bits[(size_t)leaveRow * 4 + j] = 1;
            sink(iteration_block(iter, Bidx, Nidx, names, Binv, xB, z, &rN, entering, &d, &bestTheta), mk);
        }
    }
    throw SolveError(ERR_ITER_LIMIT, "Iteration limit exceeded in Revised Primal Simplex.");
}

}  // namespace orc
