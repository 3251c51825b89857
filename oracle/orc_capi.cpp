// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.h).  C wrappers for ctypes.
#include "oracle.h"
#include "orc_dotnet.hpp"
#include "orc_solvers.hpp"

#include <atomic>
#include <cstring>
#include <thread>

using namespace orc;

static thread_local std::string t_err;
static thread_local std::string t_fmt;

static Problem make_problem(int m, int n, int sense, const double* A, const int* rel, const double* b,
                            const double* c) {
    Problem p;
    p.sense = sense;
    p.c.assign(c, c + n);
    p.rows.resize(m);
    for (int i = 0; i < m; i++) {
        p.rows[i].a.assign(A + (size_t)i * n, A + (size_t)(i + 1) * n);
        p.rows[i].rel = rel ? rel[i] : LE;
        p.rows[i].b = b[i];
    }
    return p;
}

extern "C" {

void orc_set_newline(const char* nl) { g_newline = nl; }
const char* orc_last_error(void) { return t_err.c_str(); }

int orc_parse_text(const char* text, int* sense, int* m, int* n, double* A, int* rel, double* b, double* c) {
    try {
        Problem p = parse_text(text);
        *sense = p.sense;
        *m = (int)p.rows.size();
        *n = p.nvars();
        if (A) {
            for (int j = 0; j < *n; j++) c[j] = p.c[j];
            for (int i = 0; i < *m; i++) {
                // rows may be ragged (the reference only fails later, in BuildTableau)
                for (int j = 0; j < *n; j++) A[(size_t)i * *n + j] = j < (int)p.rows[i].a.size() ? p.rows[i].a[j] : 0.0;
                if ((int)p.rows[i].a.size() < *n) {
                    t_err = "Index was outside the bounds of the array.";
                    return ERR_BAD_ARGS;
                }
                rel[i] = p.rows[i].rel;
                b[i] = p.rows[i].b;
            }
        }
        return 0;
    } catch (const SolveError& e) {
        t_err = e.what();
        return e.code;
    }
}

void orc_tableau_dims(int m, int n, const int* rel, int* rows, int* cols) {
    int mm = 0;
    for (int i = 0; i < m; i++) mm += (rel && rel[i] == EQ) ? 2 : 1;
    *rows = mm + 1;
    *cols = n + mm + 1;
}

static void export_outcome(const Outcome& o, const Trace& t, int* n_pivots, int* pivots, int pivots_cap, int* basis,
                           double* x, double* z, double* tableau, double* history, int history_cap) {
    *n_pivots = (int)t.enter.size();
    for (int k = 0; k < *n_pivots && k < pivots_cap; k++) {
        pivots[2 * k] = t.enter[k];
        pivots[2 * k + 1] = t.leave[k];
    }
    if (o.has_tableau) {
        if (basis) std::memcpy(basis, o.basis.data(), o.basis.size() * sizeof(int));
        if (tableau) std::memcpy(tableau, o.T.data(), o.T.size() * sizeof(double));
    }
    if (o.has_x && x) std::memcpy(x, o.x.data(), o.x.size() * sizeof(double));
    if (z) *z = o.z;
    if (history)
        for (int k = 0; k < (int)t.history.size() && k < history_cap; k++)
            std::memcpy(history + (size_t)k * t.history[k].size(), t.history[k].data(),
                        t.history[k].size() * sizeof(double));
}

int orc_primal_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                     int max_iterations, int* status, int* n_pivots, int* pivots, int pivots_cap, int* basis,
                     double* x, double* z, double* tableau, double* history, int history_cap) {
    Problem p = make_problem(m, n, sense, A, rel, b, c);
    Trace t;
    t.keep_history = history != nullptr;
    PrimalOptions o;
    o.max_iterations = max_iterations;
    o.format_every_iteration = false;
    try {
        Outcome out = primal_simplex(p, Sink(), &t, o);
        *status = t.status;
        export_outcome(out, t, n_pivots, pivots, pivots_cap, basis, x, z, tableau, history, history_cap);
        return 0;
    } catch (const SolveError& e) {
        t_err = e.what();
        *status = e.code;
        *n_pivots = (int)t.enter.size();
        for (int k = 0; k < *n_pivots && k < pivots_cap; k++) {
            pivots[2 * k] = t.enter[k];
            pivots[2 * k + 1] = t.leave[k];
        }
        return e.code;
    }
}

// DualSimplex returns no tableau/solution upstream; for arithmetic parity the oracle re-derives
// them from a second run that keeps the working arrays (same code path, full=true is not
// available there), so we expose the last history entry instead.
int orc_dual_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                   int* status, int* n_pivots, int* silent, int* pivots, int pivots_cap, int* basis, double* x,
                   double* z, double* tableau, double* history, int history_cap) {
    Problem p = make_problem(m, n, sense, A, rel, b, c);
    Trace t;
    t.keep_history = true;
    try {
        Outcome out = dual_simplex(p, Sink(), &t, false);
        *status = t.status;
        *silent = t.silent_pivots;
        export_outcome(out, t, n_pivots, pivots, pivots_cap, nullptr, nullptr, nullptr, nullptr, history,
                       history_cap);
        if (!t.history.empty()) {
            const std::vector<double>& T = t.history.back();
            int rows, cols;
            orc_tableau_dims(m, n, rel, &rows, &cols);
            if (tableau) std::memcpy(tableau, T.data(), T.size() * sizeof(double));
            // rebuild basis from the pivot list
            std::vector<int> bs(rows - 1);
            for (int i = 0; i < rows - 1; i++) bs[i] = n + i;
            for (size_t k = 0; k < t.enter.size(); k++) bs[t.leave[k]] = t.enter[k];
            if (basis) std::memcpy(basis, bs.data(), bs.size() * sizeof(int));
            if (x) {
                for (int j = 0; j < n; j++) x[j] = 0;
                for (int i = 0; i < rows - 1; i++)
                    if (bs[i] < n) x[bs[i]] = T[(size_t)i * cols + cols - 1];
            }
            if (z) *z = T[(size_t)(rows - 1) * cols + cols - 1];
        }
        return 0;
    } catch (const SolveError& e) {
        t_err = e.what();
        *status = e.code;
        *n_pivots = (int)t.enter.size();
        *silent = t.silent_pivots;
        return e.code;
    }
}

int orc_primal_core(double* T, int m, int width, int* basis, int max_pivots, int* n_pivots, int* pivots,
                    int pivots_cap) {
    return primal_core(T, m, width, basis, max_pivots, n_pivots, pivots, pivots_cap);
}

long orc_primal_batch(int count, int m, int n, const double* A, const double* b, const double* c,
                      int max_iterations, int threads, int with_format, int* status, int* n_pivots, int* basis,
                      double* x, double* z, double* tableau) {
    if (threads < 1) threads = 1;
    std::atomic<int> next(0);
    std::atomic<long> total(0);
    const int rows = m + 1, cols = n + m + 1;
    auto worker = [&]() {
        std::vector<double> T((size_t)rows * cols);
        std::vector<int> bs(m);
        std::vector<int> ones(m, LE);
        while (true) {
            int k = next.fetch_add(1);
            if (k >= count) break;
            const double* Ak = A + (size_t)k * m * n;
            const double* bk = b + (size_t)k * m;
            const double* ck = c + (size_t)k * n;
            int st, np = 0;
            if (with_format) {
                Problem p = make_problem(m, n, MAX, Ak, ones.data(), bk, ck);
                Trace t;
                PrimalOptions o;
                o.max_iterations = max_iterations;
                o.format_every_iteration = true;
                try {
                    Outcome out = primal_simplex(p, Sink(), &t, o);
                    st = t.status;
                    T = out.T;
                    bs = out.basis;
                } catch (const SolveError& e) {
                    st = e.code;
                }
                np = (int)t.enter.size();
            } else {
                std::fill(T.begin(), T.end(), 0.0);
                for (int i = 0; i < m; i++) {
                    for (int j = 0; j < n; j++) T[(size_t)i * cols + j] = Ak[(size_t)i * n + j];
                    T[(size_t)i * cols + n + i] = 1.0;
                    T[(size_t)i * cols + n + m] = bk[i];
                    bs[i] = n + i;
                }
                for (int j = 0; j < n; j++) T[(size_t)m * cols + j] = -ck[j];
                st = primal_core(T.data(), m, cols, bs.data(), max_iterations, &np, nullptr, 0);
            }
            total += np;
            if (status) status[k] = st;
            if (n_pivots) n_pivots[k] = np;
            if (basis) std::memcpy(basis + (size_t)k * m, bs.data(), m * sizeof(int));
            if (x) {
                double* xk = x + (size_t)k * n;
                for (int j = 0; j < n; j++) xk[j] = 0;
                for (int i = 0; i < m; i++)
                    if (bs[i] < n) xk[bs[i]] = T[(size_t)i * cols + cols - 1];
            }
            if (z) z[k] = T[(size_t)m * cols + cols - 1];
            if (tableau) std::memcpy(tableau + (size_t)k * rows * cols, T.data(), T.size() * sizeof(double));
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return total.load();
}

int orc_bnb_simplex(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                    int* found, double* best_z, double* best_x, int* n_nodes, long* total_pivots, int node_cap,
                    int* node_outcome, int* node_algo, int* node_pivots, double* node_z, int* node_branch_var,
                    int* node_depth) {
    Problem p = make_problem(m, n, sense, A, rel, b, c);
    BnbTrace t;
    try {
        branch_and_bound(p, Sink(), &t, false);
    } catch (const SolveError& e) {
        t_err = e.what();
        return e.code;
    }
    *found = t.found ? 1 : 0;
    *best_z = t.best_z;
    if (t.found && best_x)
        for (int j = 0; j < n; j++) best_x[j] = t.best_x[j];
    *n_nodes = (int)t.nodes.size();
    if (total_pivots) *total_pivots = t.total_pivots;
    for (int k = 0; k < *n_nodes && k < node_cap; k++) {
        const BnbNode& nd = t.nodes[k];
        if (node_outcome) node_outcome[k] = nd.outcome;
        if (node_algo) node_algo[k] = nd.algo;
        if (node_pivots) node_pivots[k] = nd.n_pivots;
        if (node_z) node_z[k] = nd.z;
        if (node_branch_var) node_branch_var[k] = nd.branch_var;
        if (node_depth) node_depth[k] = nd.depth;
    }
    return 0;
}

int orc_bnb_pooled(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c, int batch,
                   long max_nodes, int* found, double* best_z, double* best_x, long* n_nodes, long* total_pivots, long* rounds,
                   long* skipped, long node_cap, int* node_id, int* node_outcome, int* node_pivots, double* node_z) {
    Problem p = make_problem(m, n, sense, A, rel, b, c);
    PooledTrace t;
    try {
        bnb_pooled(p, batch, max_nodes > 0 ? max_nodes : (1L << 30), &t);
    } catch (const SolveError& e) {
        t_err = e.what();
        return e.code;
    }
    *found = t.found ? 1 : 0;
    *best_z = t.best_z;
    if (t.found && best_x)
        for (int j = 0; j < n; j++) best_x[j] = t.best_x[j];
    *n_nodes = (long)t.node_id.size();
    if (total_pivots) *total_pivots = t.pivots;
    if (rounds) *rounds = t.rounds;
    if (skipped) *skipped = t.skipped;
    for (long k = 0; k < *n_nodes && k < node_cap; k++) {
        if (node_id) node_id[k] = t.node_id[k];
        if (node_outcome) node_outcome[k] = t.node_outcome[k];
        if (node_pivots) node_pivots[k] = t.node_pivots[k];
        if (node_z) node_z[k] = t.node_z[k];
    }
    return 0;
}

int orc_knapsack(int n, const double* profit, const double* weight, double capacity, int* found, double* best,
                 int* best_x, long* n_evals, long* n_pops, long eval_cap, int* ev_parent, int* ev_child,
                 int* ev_var, double* ev_bound, double* ev_weight, int* ev_frac, int* ev_decision) {
    Problem p;
    p.sense = MAX;
    p.c.assign(profit, profit + n);
    p.rows.resize(1);
    p.rows[0].a.assign(weight, weight + n);
    p.rows[0].rel = LE;
    p.rows[0].b = capacity;
    KnapTrace t;
    try {
        knapsack_bnb(p, Sink(), &t, false);
    } catch (const SolveError& e) {
        t_err = e.what();
        return e.code;
    }
    *found = t.found ? 1 : 0;
    *best = t.best;
    if (best_x)
        for (int j = 0; j < n; j++) best_x[j] = t.best_x[j];
    *n_evals = (long)t.evals.size();
    if (n_pops) *n_pops = t.pops;
    for (long k = 0; k < *n_evals && k < eval_cap; k++) {
        const KnapEval& e = t.evals[k];
        if (ev_parent) ev_parent[k] = e.parent_pop;
        if (ev_child) ev_child[k] = e.child;
        if (ev_var) ev_var[k] = e.var;
        if (ev_bound) ev_bound[k] = e.bound;
        if (ev_weight) ev_weight[k] = e.weight;
        if (ev_frac) ev_frac[k] = e.frac_sorted;
        if (ev_decision) ev_decision[k] = e.decision;
    }
    return 0;
}

// RevisedPrimalSimplex numerics: pivots (entering column, leaving row, theta) and the final basis state.
int orc_revised_solve(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                      int max_iterations, int* status, int* n_iters, int* enter, int* leave, double* theta, int cap,
                      int* basis, double* xB, double* Binv, double* x, double* z_original) {
    try {
        Problem p = make_problem(m, n, sense, A, rel, b, c);
        RevTrace t;
        Outcome o;
        int st = 0;
        try {
            o = revised_primal_simplex(p, Sink(), &t, max_iterations);
            st = t.status;
        } catch (const SolveError& e) {
            if (e.code != ERR_ITER_LIMIT) {
                *status = e.code;
                *n_iters = 0;
                return 0;
            }
            st = ERR_ITER_LIMIT;
        }
        *status = st;
        *n_iters = (int)t.enter.size();
        for (int k = 0; k < *n_iters && k < cap; k++) {
            if (enter) enter[k] = t.enter[k];
            if (leave) leave[k] = t.leave[k];
            if (theta) theta[k] = t.theta[k];
        }
        for (int i = 0; i < m; i++) {
            if (basis) basis[i] = t.basis[i];
            if (xB) xB[i] = t.xB[i];
        }
        if (Binv)
            for (size_t k = 0; k < t.Binv.size(); k++) Binv[k] = t.Binv[k];
        if (st >= 0) {
            if (x)
                for (int j = 0; j < n; j++) x[j] = o.x[j];
            if (z_original) *z_original = o.z;
        }
        return 0;
    } catch (const std::exception& e) {
        t_err = e.what();
        return -1;
    }
}

struct orc_text {
    int code = 0;
    int chunks = 0;
    std::vector<Mask> masks;          // one per callback chunk (rows == 0: null)
    std::vector<size_t> chunk_len;    // characters of each chunk
    std::string error, log, report, summary;
    Outcome result;   // numeric part of the SimplexResult (Tableau / Solution / Basis may be null)
    CutTrace cuts;
};

orc_text* orc_solve_text(const char* input, const char* algorithm) {
    orc_text* t = new orc_text();
    Sink sink = [t](const std::string& s, const Mask& mk) {
        t->log += s;
        t->chunks++;
        t->masks.push_back(mk);
        t->chunk_len.push_back(s.size());
    };
    try {
        Problem p = parse_text(input);
        Outcome o;
        std::string algo = algorithm;
        if (algo == "knapsack") o = knapsack_bnb(p, sink, nullptr, true);
        else if (algo == "cutting plane") o = cutting_plane(p, sink, &t->cuts);   // Form1.cs:249-254
        else if (algo == "revised cutting plane") o = cutting_plane_revised(p, sink, &t->cuts);   // Form1.cs:256-261
        else o = lp_solver_solve(p, algo, sink, nullptr, true);
        t->report = o.report;
        t->summary = o.summary;
        t->result = o;
    } catch (const SolveError& e) {
        t->code = e.code;
        t->error = e.what();
    }
    return t;
}
int orc_text_code(const orc_text* t) { return t->code; }
const char* orc_text_error(const orc_text* t) { return t->error.c_str(); }
const char* orc_text_log(const orc_text* t) { return t->log.c_str(); }
const char* orc_text_report(const orc_text* t) { return t->report.c_str(); }
const char* orc_text_summary(const orc_text* t) { return t->summary.c_str(); }
int orc_text_masks(const orc_text* t) { return t->chunks; }
long orc_text_chunk_len(const orc_text* t, int k) { return k >= 0 && k < t->chunks ? (long)t->chunk_len[k] : -1; }
int orc_text_mask(const orc_text* t, int k, int* rows, int* cols, unsigned char* bits, long cap) {
    if (k < 0 || k >= t->chunks) return -1;
    const Mask& mk = t->masks[k];
    *rows = mk.rows;
    *cols = mk.cols;
    const long need = (long)mk.rows * mk.cols;
    if (bits && cap >= need)
        for (long q = 0; q < need; q++) bits[q] = mk.bits[q];
    return mk.rows ? 1 : 0;
}
int orc_text_result_dims(const orc_text* t, int* rows, int* cols, int* nx, int* nbasis) {
    *rows = t->result.has_tableau ? t->result.rows : 0;
    *cols = t->result.has_tableau ? t->result.cols : 0;
    *nx = t->result.has_x ? (int)t->result.x.size() : 0;
    *nbasis = (int)t->result.basis.size();
    return t->result.has_tableau ? 1 : 0;
}
const double* orc_text_tableau(const orc_text* t) { return t->result.T.data(); }
const double* orc_text_solution(const orc_text* t) { return t->result.x.data(); }
const int* orc_text_basis(const orc_text* t) { return t->result.basis.data(); }
double orc_text_z(const orc_text* t) { return t->result.z; }
int orc_text_cut_count(const orc_text* t) { return (int)t->cuts.cut_b.size(); }
int orc_text_cut(const orc_text* t, int k, int* frac_var, int* row, double* a, double* b) {
    if (k < 0 || k >= (int)t->cuts.cut_b.size()) return -1;
    *frac_var = t->cuts.frac_var[k];
    *row = t->cuts.cut_row[k];
    *b = t->cuts.cut_b[k];
    for (size_t j = 0; j < t->cuts.cut_a[k].size(); j++) a[j] = t->cuts.cut_a[k][j];
    return (int)t->cuts.cut_a[k].size();
}
int orc_text_cut_end(const orc_text* t) { return t->cuts.end; }
void orc_text_free(orc_text* t) { delete t; }

const char* orc_fmt_custom(double v, int decimals) { t_fmt = fmt_custom(v, decimals); return t_fmt.c_str(); }
const char* orc_fmt_fixed(double v, int decimals) { t_fmt = fmt_fixed(v, decimals); return t_fmt.c_str(); }
const char* orc_fmt_roundtrip(double v) { t_fmt = fmt_roundtrip(v); return t_fmt.c_str(); }
double orc_math_round(double v, int digits) { return math_round(v, digits); }

}  // extern "C"
