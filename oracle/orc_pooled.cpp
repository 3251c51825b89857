// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).
//
// "Pooled tree" Branch & Bound (Mode B) — NOT what the reference computes.  The reference's
// BranchAndBound never explores a '>=' child (its Dual Simplex result is rejected), so
// its tree is a single floor path.  This is the tree its own doc comment describes
// (R/Models/Branch&Bound.cs:9-19: branch on the fractional part closest to 0.5, lowest subscript on
// ties, ceil branch first) with both children honoured, written down here as the checker of the
// engine's lpx_bnb_pooled.  There is no upstream code to pin it to; it is pinned by (a) agreeing
// bit for bit with the engine, node by node, and (b) reaching the optimum an independent MILP solver
// finds (tests/test_gpu_pooled.py).
//
// The search is defined so that it does not depend on how many GPUs evaluate it:
//   root      Primal Simplex on the all-'<=' problem with the reference's rules
//             (R/Models/PrimalSimplex.cs:205-257); its final tableau is kept.
//   child     parent's final tableau + one row (x_k <= floor or x_k >= ceil, written in the parent's
//             non-basic variables) + one slack column, then Dual Simplex pivots with the reference's
//             rules (R/Models/DualSimplex.cs:45-113: most negative RHS below -1e-9 leaves, minimum
//             z_j / -a ratio with the 1e-12 margin scan enters, limit 10000) — a warm start.
//   rounds    each round takes the (at most) `batch` open nodes with the largest bound (the parent's z;
//             lower node id on ties) that can still beat the incumbent, evaluates them, and commits
//             them in that order: infeasible | z <= best + 1e-6 pruned | integral -> incumbent (z, x
//             rounded) | branched: ceil child then floor child join the pool.
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <queue>

#include "orc_solvers.hpp"

namespace orc {

namespace {

const double kBB = 1e-6;  // BranchAndBound.EPS

struct PNode {
    int id = 0, parent = -1, var = -1, side = 0, bound_val = 0, depth = 0;
    double bound = 0;  // the parent's z: an upper bound on this node's
    // after evaluation
    int status = 0, pivots = 0;
    double z = 0;
    std::vector<double> x;
    std::vector<double> T;  // final tableau, rows x cols
    std::vector<int> basis;
    int rows = 0, cols = 0;
    int open_children = 0;
};

struct ByBound {
    bool operator()(const PNode* a, const PNode* b) const {  // priority_queue: "less" = lower priority
        if (a->bound != b->bound) return a->bound < b->bound;
        return a->id > b->id;
    }
};

// the child tableau of `par` for the row x_var <= val (side 0) or x_var >= val (side 1)
void extend_tableau(const PNode& par, int var, int side, double val, PNode& ch) {
    const int rp = par.rows, cp = par.cols, mp = rp - 1, rhs_p = cp - 1;
    int r = -1;
    for (int i = 0; i < mp; i++)
        if (par.basis[i] == var) r = i;
    ch.rows = rp + 1;
    ch.cols = cp + 1;
    ch.T.assign((size_t)ch.rows * ch.cols, 0.0);
    auto at = [&](int i, int j) -> double& { return ch.T[(size_t)i * ch.cols + j]; };
    auto pt = [&](int i, int j) { return par.T[(size_t)i * cp + j]; };
    for (int i = 0; i < mp; i++) {
        for (int j = 0; j < rhs_p; j++) at(i, j) = pt(i, j);
        at(i, rhs_p) = 0.0;
        at(i, cp) = pt(i, rhs_p);
    }
    for (int j = 0; j < rhs_p; j++) {
        const double unit = j == var ? 1.0 : 0.0;
        at(mp, j) = side == 0 ? unit - pt(r, j) : pt(r, j) - unit;
    }
    at(mp, rhs_p) = 1.0;
    at(mp, cp) = side == 0 ? val - pt(r, rhs_p) : pt(r, rhs_p) - val;
    for (int j = 0; j < rhs_p; j++) at(mp + 1, j) = pt(mp, j);
    at(mp + 1, rhs_p) = 0.0;
    at(mp + 1, cp) = pt(mp, rhs_p);
    ch.basis = par.basis;
    ch.basis.push_back(rhs_p);
}

// Dual Simplex pivots on a dual-feasible tableau (DualSimplex.cs:36-113); status 0 optimal, 2 infeasible, -3 limit
int dual_core(std::vector<double>& T, int rows, int cols, std::vector<int>& basis, int* n_pivots) {
    const int m = rows - 1, ns = cols - 1;
    int iter = 1;
    *n_pivots = 0;
    while (true) {
        if (iter > 10000) return ERR_ITER_LIMIT;
        int leave = -1;
        double most_neg = -1e-9;
        for (int i = 0; i < m; i++) {
            const double rhs = T[(size_t)i * cols + ns];
            if (rhs < most_neg) {
                most_neg = rhs;
                leave = i;
            }
        }
        if (leave == -1) return 0;
        int enter = -1;
        double best_ratio = std::numeric_limits<double>::infinity();
        for (int j = 0; j < ns; j++) {
            const double a = T[(size_t)leave * cols + j];
            if (a < -1e-9) {
                const double ratio = T[(size_t)m * cols + j] / (-a);
                if (ratio < best_ratio - 1e-12) {
                    best_ratio = ratio;
                    enter = j;
                }
            }
        }
        if (enter == -1) return 2;
        pivot(T.data(), rows, cols, leave, enter);
        basis[leave] = enter;
        (*n_pivots)++;
        iter++;
    }
}

void read_solution(PNode& nd, int n) {
    const int m = nd.rows - 1, rhs = nd.cols - 1;
    nd.x.assign(n, 0.0);
    for (int i = 0; i < m; i++)
        if (nd.basis[i] < n) nd.x[nd.basis[i]] = nd.T[(size_t)i * nd.cols + rhs];
    nd.z = nd.T[(size_t)m * nd.cols + rhs];
}

}  // namespace

int bnb_pooled(const Problem& original, int batch, long max_nodes, PooledTrace* t) {
    Problem model = original;
    if (model.sense == MIN)
        for (double& v : model.c) v = -v;
    const int n = model.nvars(), m = (int)model.rows.size();
    for (const Row& r : model.rows)
        if (r.rel != LE || r.b < -1e-9) throw SolveError(ERR_GE_ROW, "pooled B&B: all rows must be '<=' with b >= 0");
    std::vector<std::unique_ptr<PNode>> nodes;
    auto root = std::make_unique<PNode>();
    root->rows = m + 1;
    root->cols = n + m + 1;
    root->T.assign((size_t)root->rows * root->cols, 0.0);
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) root->T[(size_t)i * root->cols + j] = model.rows[i].a[j];
        root->T[(size_t)i * root->cols + n + i] = 1.0;
        root->T[(size_t)i * root->cols + n + m] = model.rows[i].b;
    }
    for (int j = 0; j < n; j++) root->T[(size_t)m * root->cols + j] = -model.c[j];
    root->basis.resize(m);
    for (int i = 0; i < m; i++) root->basis[i] = n + i;
    root->status = primal_core(root->T.data(), m, root->cols, root->basis.data(), 10000, &root->pivots, nullptr, 0);
    root->bound = std::numeric_limits<double>::infinity();
    nodes.push_back(std::move(root));

    t->found = false;
    t->best_z = -std::numeric_limits<double>::infinity();
    t->best_x.assign(n, 0.0);
    t->rounds = 0;
    std::priority_queue<PNode*, std::vector<PNode*>, ByBound> pool;

    auto release = [&](int id) {  // a node's tableau is needed until both children are evaluated or discarded
        if (id < 0) return;
        PNode& p = *nodes[id];
        if (--p.open_children == 0) {
            std::vector<double>().swap(p.T);
        }
    };
    auto commit = [&](PNode& nd) {
        t->node_id.push_back(nd.id);
        t->node_status.push_back(nd.status);
        t->node_pivots.push_back(nd.pivots);
        t->pivots += nd.pivots;
        int outcome;
        if (nd.status != 0) {
            outcome = BNB_INFEASIBLE;
            t->node_z.push_back(0.0);
        } else {
            read_solution(nd, n);
            t->node_z.push_back(nd.z);
            bool integral = true;
            for (double v : nd.x)
                if (std::fabs(v - std::nearbyint(v)) > kBB) integral = false;
            if (nd.z <= t->best_z + kBB) outcome = BNB_PRUNED;
            else if (integral) {
                outcome = BNB_INCUMBENT;
                t->found = true;
                t->best_z = nd.z;
                for (int j = 0; j < n; j++) t->best_x[j] = std::nearbyint(nd.x[j]);
            } else {
                int k = -1;
                double min_dist = std::numeric_limits<double>::max();
                for (int i = 0; i < n; i++) {
                    const double frac = nd.x[i] - std::floor(nd.x[i]);
                    if (frac > kBB && (1 - frac) > kBB) {
                        const double dist = std::fabs(frac - 0.5);
                        if (dist < min_dist) {
                            min_dist = dist;
                            k = i;
                        }
                    }
                }
                if (k < 0) outcome = BNB_NOFRAC;
                else {
                    outcome = BNB_BRANCHED;
                    for (int side = 1; side >= 0; side--) {  // ceil child first
                        auto ch = std::make_unique<PNode>();
                        ch->id = (int)nodes.size();
                        ch->parent = nd.id;
                        ch->var = k;
                        ch->side = side;
                        ch->bound_val = (int)(side ? std::ceil(nd.x[k]) : std::floor(nd.x[k]));
                        ch->depth = nd.depth + 1;
                        ch->bound = nd.z;
                        pool.push(ch.get());
                        nodes.push_back(std::move(ch));
                    }
                    nd.open_children = 2;
                }
            }
        }
        t->node_outcome.push_back(outcome);
        if (outcome != BNB_BRANCHED) std::vector<double>().swap(nd.T);
    };

    commit(*nodes[0]);
    while (!pool.empty()) {
        std::vector<PNode*> sel;
        while (!pool.empty() && (int)sel.size() < batch) {
            PNode* nd = pool.top();
            pool.pop();
            if (nd->bound <= t->best_z + kBB) {  // cannot beat the incumbent any more: no LP
                t->skipped++;
                release(nd->parent);
                continue;
            }
            sel.push_back(nd);
        }
        if (sel.empty()) break;
        t->rounds++;
        for (PNode* nd : sel) {
            extend_tableau(*nodes[nd->parent], nd->var, nd->side, (double)nd->bound_val, *nd);
            nd->status = dual_core(nd->T, nd->rows, nd->cols, nd->basis, &nd->pivots);
        }
        for (PNode* nd : sel) {
            commit(*nd);
            release(nd->parent);
        }
        if ((long)t->node_id.size() > max_nodes) return 1;
    }
    return 0;
}

}  // namespace orc
