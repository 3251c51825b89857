// Host-side RevisedPrimalSimplex: the reference's interface and text
// (R/Models/RevisedPrimalSimplex.cs:17-145, 190-297) over lpx_revised_solve.  The engine returns one
// record per BuildIterationBlock call — B^-1, x_B, z, the reduced costs, the direction, theta and the
// basic / nonbasic lists — and this file prints them; the only arithmetic left here is what the
// reference's own printing code does (the x_B / d quotients of the "ratio test" lines and z* from the
// original objective).
#include <cmath>

#include "../../include/lpx.h"
#include "dotnet_text.hpp"
#include "host_util.hpp"
#include "lp_model.hpp"

namespace lpr381 {

using text::custom_hash;
using text::pad_left;

namespace {

const double kEps = 1e-9;

struct RevRecord {  // one history record, see include/lpx.h
    const double *Binv, *xB, *rN, *d, *Bidx, *Nidx;
    double z, theta;
    int entering;
};
RevRecord view(const double* h, int m, int n) {
    RevRecord r;
    r.Binv = h;
    h += (size_t)m * m;
    r.xB = h;
    h += m;
    r.z = *h++;
    r.rN = h;
    h += n;
    r.d = h;
    h += m;
    r.theta = *h++;
    r.Bidx = h;
    h += m;
    r.Nidx = h;
    h += n;
    r.entering = (int)*h;
    return r;
}

// BuildIterationBlock (:190-247)
std::string iteration_block(int iter, const RevRecord& r, int m, int n, const std::vector<std::string>& names) {
    const std::string& nl = NewLine();
    auto join_names = [&](const double* idx, int k) {
        std::string s;
        for (int q = 0; q < k; q++) s += (q ? ", " : "") + names[(int)idx[q]];
        return s;
    };
    auto join_vals = [&](const double* v, int k) {
        std::string s;
        for (int q = 0; q < k; q++) s += (q ? ", " : "") + custom_hash(v[q]);
        return s;
    };
    std::string sb = "=== Revised Simplex Iteration " + std::to_string(iter) + " ===" + nl;
    sb += "Basis: " + join_names(r.Bidx, m) + nl;
    sb += "Nonbasic: " + join_names(r.Nidx, n) + nl;
    sb += "\nProduct-form: current B^{-1}" + nl;
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < m; j++) sb += pad_left(custom_hash(r.Binv[(size_t)i * m + j]), 12);
        sb += nl;
    }
    sb += "x_B = [" + join_vals(r.xB, m) + "]" + nl;
    sb += "z = " + custom_hash(r.z) + nl;
    if (iter > 0) {
        sb += "\nReduced costs (r_N = c_N - c_B^T B^{-1} N):" + nl;
        for (int j = 0; j < n; j++)
            sb += "  r(" + std::to_string((int)r.Nidx[j]) + ":" + names[(int)r.Nidx[j]] + ") = " + custom_hash(r.rN[j]) + nl;
        sb += "\nEntering variable: " + names[r.entering] + nl;
        sb += "Direction d = B^{-1} * a_entering:" + nl;
        sb += "  d = [" + join_vals(r.d, m) + "]" + nl;
        sb += "\nRatio test (theta):" + nl;
        for (int i = 0; i < m; i++) {
            if (r.d[i] > kEps)
                sb += "  row " + std::to_string(i + 1) + ": " + custom_hash(r.xB[i]) + " / " + custom_hash(r.d[i]) + " = " +
                      custom_hash(r.xB[i] / r.d[i]) + nl;
            else
                sb += "  row " + std::to_string(i + 1) + ": d_i <= 0 (skip)" + nl;
        }
        sb += "Chosen theta* = " + custom_hash(r.theta) + nl;
    }
    sb += nl;
    return sb;
}

}  // namespace

SimplexResult RevisedPrimalSimplex::Solve(const LPProblem& original, UpdatePivot updatePivot) {
    const std::string& nl = NewLine();
    Flat f = flatten(original);
    lpx_options opt;
    lpx_default_options(&opt);
    int status = 0, iters = 0;
    std::vector<int> basis(f.m), nonbasic(f.n);
    std::vector<double> xB(f.m), x(f.n), hist;
    int hist_cap = 0;
    if (updatePivot) {
        // the engine returns every iteration's record in one call; a first pass sizes the buffer
        throw_on(lpx_revised_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status,
                                   &iters, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0));
        if (status >= 0 || status == LPX_S_ITER_LIMIT || status == LPX_S_SINGULAR) {
            hist_cap = iters + 1;
            hist.resize(lpx_revised_history_stride(f.m, f.n) * (size_t)hist_cap);
        }
    }
    throw_on(lpx_revised_solve(f.m, f.n, f.sense, f.A.data(), f.rel.data(), f.b.data(), f.c.data(), &opt, &status, &iters,
                               nullptr, nullptr, 0, basis.data(), nonbasic.data(), xB.data(), nullptr, x.data(),
                               hist_cap ? hist.data() : nullptr, hist_cap));
    if (status == LPX_S_REV_UNSUPPORTED) throw LpException(lpx_status_message(status), status);

    std::vector<std::string> names;
    for (int j = 0; j < f.n; j++) names.push_back("x" + std::to_string(j + 1));
    for (int j = 0; j < f.m; j++) names.push_back("c" + std::to_string(j + 1));
    if (updatePivot) {
        const size_t hs = lpx_revised_history_stride(f.m, f.n);
        // a singular basis is met inside Invert, before the block of that iteration is printed
        const int blocks = std::min(hist_cap, status == LPX_S_SINGULAR ? iters : iters + 1);
        for (int k = 0; k < blocks; k++) {
            const RevRecord r = view(hist.data() + hs * k, f.m, f.n);
            Highlight hl;
            if (k > 0) {  // new bool[m, 4] with the leaving row set (:129-130)
                hl.rows = f.m;
                hl.cols = 4;
                hl.v.assign((size_t)f.m * 4, 0);
                int leave_row = 0;
                for (int i = 0; i < f.m; i++)
                    if ((int)r.Bidx[i] == r.entering) leave_row = i;
                for (int j = 0; j < 4; j++) hl.v[(size_t)leave_row * 4 + j] = 1;
            }
            updatePivot(iteration_block(k, r, f.m, f.n, names), hl);
        }
    }
    if (status == LPX_S_SINGULAR) throw LpException(lpx_status_message(status), status);
    if (status == LPX_S_ITER_LIMIT) throw LpException("Iteration limit exceeded in Revised Primal Simplex.", status);

    // BuildFinalSummary (:264-297): Report and Summary only; z* from the ORIGINAL objective
    const char* st = status == LPX_UNBOUNDED ? "UNBOUNDED" : "OPTIMAL";
    std::string sb = "\nStatus: " + std::string(st) + nl;
    for (int j = 0; j < f.n; j++) sb += "  x" + std::to_string(j + 1) + " = " + custom_hash(text::math_round(x[j], 3)) + nl;
    std::string summary = "Status: " + std::string(st) + nl + "x* = [";
    for (int j = 0; j < f.n; j++) summary += (j ? ", " : "") + text::round_trip(text::math_round(x[j], 3));
    summary += "]" + nl;
    double zOriginal = 0;
    for (int j = 0; j < f.n; j++) zOriginal += original.C[j] * x[j];
    sb += "  z* = " + custom_hash(text::math_round(zOriginal, 3)) + nl;
    summary += "z* = " + custom_hash(text::math_round(zOriginal, 3)) + nl;
    SimplexResult r;
    r.Report = sb;
    r.Summary = summary;
    return r;
}

}  // namespace lpr381
