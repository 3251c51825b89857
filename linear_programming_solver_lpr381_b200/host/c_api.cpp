// C wrappers over the host layer for ctypes (tests) — the host layer itself is C++.
#include <cstring>
#include <string>

#include "dotnet_text.hpp"
#include "lp_model.hpp"

using namespace lpr381;

struct lpr_text {
    int code = 0, chunks = 0, highlighted = 0;
    std::vector<Highlight> masks;    // one per callback chunk (rows == 0: null)
    std::vector<size_t> chunk_len;
    std::string error, log, report, summary;
    SimplexResult result;
    std::vector<Constraint> cuts;
};

static thread_local std::string t_buf;

extern "C" {

void lpr_set_newline(const char* nl) { NewLine() = nl; }

// algorithm: any LPSolver key; "cutting plane" constructs CuttingPlane directly (Form1.cs:249-254);
// "knapsack" constructs BranchAndBoundKnapsack directly (it has no
// LPSolver key upstream); "controller" goes through LPController.SolvePrimalSimplex (no callback).
lpr_text* lpr_solve_text(const char* input, const char* algorithm) {
    lpr_text* t = new lpr_text();
    UpdatePivot cb = [t](const std::string& s, const Highlight& h) {
        t->log += s;
        t->chunks++;
        t->masks.push_back(h);
        t->chunk_len.push_back(s.size());
        if (h.rows) t->highlighted++;
    };
    try {
        LPProblem p = LPParser::ParseFromText(input);
        SimplexResult r;
        const std::string algo = algorithm;
        if (algo == "knapsack") r = BranchAndBoundKnapsack().Solve(p, cb);
        else if (algo == "controller") r = LPController::SolvePrimalSimplex(p);
        else if (algo == "revised cutting plane") {
            CuttingPlaneRevised cp;
            r = cp.Solve(p, cb);
            t->cuts = cp.Cuts;
        } else if (algo == "cutting plane") {
            CuttingPlane cp;
            r = cp.Solve(p, cb);
            t->cuts = cp.Cuts;
        }
        else r = LPSolver().Solve(p, algo, cb);
        t->report = r.Report;
        t->summary = r.Summary;
        t->result = r;
    } catch (const LpException& e) {
        t->code = e.code ? e.code : -1;
        t->error = e.what();
    }
    return t;
}
int lpr_text_code(const lpr_text* t) { return t->code; }
int lpr_text_chunks(const lpr_text* t) { return t->chunks; }
int lpr_text_highlighted(const lpr_text* t) { return t->highlighted; }
long lpr_text_chunk_len(const lpr_text* t, int k) { return k >= 0 && k < t->chunks ? (long)t->chunk_len[k] : -1; }
// the bool[,] handed to updatePivot with chunk k (what Form1.AppendPivotRow paints, Form1.cs:326-368)
int lpr_text_mask(const lpr_text* t, int k, int* rows, int* cols, unsigned char* bits, long cap) {
    if (k < 0 || k >= t->chunks) return -1;
    const Highlight& h = t->masks[k];
    *rows = h.rows;
    *cols = h.cols;
    const long need = (long)h.rows * h.cols;
    if (bits && cap >= need)
        for (long q = 0; q < need; q++) bits[q] = h.v[q];
    return h.rows ? 1 : 0;
}
const char* lpr_text_error(const lpr_text* t) { return t->error.c_str(); }
const char* lpr_text_log(const lpr_text* t) { return t->log.c_str(); }
const char* lpr_text_report(const lpr_text* t) { return t->report.c_str(); }
const char* lpr_text_summary(const lpr_text* t) { return t->summary.c_str(); }
// numeric part of the SimplexResult; returns 1 when Tableau != null
int lpr_text_result_dims(const lpr_text* t, int* rows, int* cols, int* nx, int* nbasis) {
    *rows = t->result.Tableau.rows;
    *cols = t->result.Tableau.cols;
    *nx = t->result.HasSolution ? (int)t->result.Solution.size() : 0;
    *nbasis = (int)t->result.Basis.size();
    return t->result.Tableau.is_null() ? 0 : 1;
}
const double* lpr_text_tableau(const lpr_text* t) { return t->result.Tableau.v.data(); }
const double* lpr_text_solution(const lpr_text* t) { return t->result.Solution.data(); }
const int* lpr_text_basis(const lpr_text* t) { return t->result.Basis.data(); }
double lpr_text_z(const lpr_text* t) { return t->result.OptimalValue; }
int lpr_text_cut_count(const lpr_text* t) { return (int)t->cuts.size(); }
int lpr_text_cut(const lpr_text* t, int k, double* a, double* b) {
    if (k < 0 || k >= (int)t->cuts.size()) return -1;
    *b = t->cuts[k].B;
    for (size_t j = 0; j < t->cuts[k].A.size(); j++) a[j] = t->cuts[k].A[j];
    return (int)t->cuts[k].A.size();
}
void lpr_text_free(lpr_text* t) { delete t; }

// two-phase like the oracle's: A == NULL -> sizes only
int lpr_parse_text(const char* text, int* sense, int* m, int* n, double* A, int* rel, double* b, double* c,
                   char* err, int err_cap) {
    try {
        LPProblem p = LPParser::ParseFromText(text);
        *sense = (int)p.ObjectiveSense;
        *m = (int)p.Constraints.size();
        *n = p.NumVars();
        if (A) {
            for (int j = 0; j < *n; j++) c[j] = p.C[j];
            for (int i = 0; i < *m; i++) {
                for (int j = 0; j < *n; j++)
                    A[(size_t)i * *n + j] = j < (int)p.Constraints[i].A.size() ? p.Constraints[i].A[j] : 0.0;
                rel[i] = (int)p.Constraints[i].Relation;
                b[i] = p.Constraints[i].B;
            }
        }
        return 0;
    } catch (const LpException& e) {
        if (err && err_cap > 0) {
            std::strncpy(err, e.what(), err_cap - 1);
            err[err_cap - 1] = 0;
        }
        return -6;
    }
}

const char* lpr_normalize_key(const char* algorithm) {
    try {
        t_buf = LPSolver::NormalizeAlgorithmKey(algorithm);
    } catch (const LpException& e) {
        t_buf = std::string("!") + e.what();
    }
    return t_buf.c_str();
}

const char* lpr_fmt_custom(double v, int decimals) { t_buf = text::custom_hash(v, decimals); return t_buf.c_str(); }
const char* lpr_fmt_fixed(double v, int decimals) { t_buf = text::fixed(v, decimals); return t_buf.c_str(); }
const char* lpr_fmt_roundtrip(double v) { t_buf = text::round_trip(v); return t_buf.c_str(); }
double lpr_math_round(double v, int digits) { return text::math_round(v, digits); }

}  // extern "C"
