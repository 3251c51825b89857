"""Regenerates tests/golden/kat.json from tests/pyref.py (the pure-Python restatement of the C#
sources).  The reference ships no fixtures of its own, so these known-answer cases
are the pins: hand-checkable LPs/IPs whose expected values come from a second, independent
implementation, stored as hex floats so every bit is compared.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pyref  # noqa: E402

LE, GE, EQ = 0, 1, 2

LP_CASES = {
    # name: (sense, c, [(a, rel, b), ...], max_iterations)
    "wyndor": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)], 10000),
    "unbounded": (0, [1, 1], [([1, -1], LE, 1)], 10000),
    "degenerate_tie": (0, [2, 3], [([1, 1], LE, 4), ([1, 3], LE, 6), ([0, 1], LE, 2)], 10000),
    "near_tie_margin": (0, [1, 0], [([1, 0], LE, 1.0000000005), ([1, 0], LE, 1.0), ([0, 1], LE, 3)], 10000),
    "near_tie_margin_rev": (0, [1, 0], [([1, 0], LE, 1.0), ([1, 0], LE, 1.0000000005), ([0, 1], LE, 3)], 10000),
    "eq_expansion": (0, [1, 2], [([1, 1], EQ, 4), ([1, 0], LE, 3)], 10000),
    "min_trivial": (1, [2, 3], [([1, 1], LE, 4), ([1, 3], LE, 6)], 10000),
    "min_negative_costs": (1, [-2, -3], [([1, 1], LE, 4), ([1, 3], LE, 6)], 10000),
    "ge_row": (0, [1, 1], [([1, 1], LE, 4), ([1, 0], GE, 1)], 10000),
    "neg_rhs": (0, [1, 1], [([1, 1], LE, -4), ([1, 0], GE, 1)], 10000),
    "neg_rhs_tolerated": (0, [1, 1], [([1, 1], LE, -1e-10), ([1, 0], LE, 1)], 10000),
    "iter_limit_hit": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)], 2),
    "iter_limit_ok": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)], 3),
    "zero_cost_negzero": (0, [0, 1], [([1, 1], LE, 2), ([0, 0], LE, 0)], 10000),
    "klee_minty3": (0, [100, 10, 1], [([1, 0, 0], LE, 1), ([20, 1, 0], LE, 100), ([200, 20, 1], LE, 10000)], 10000),
}

DUAL_CASES = {
    "dual_ge": (1, [2, 3], [([1, 1], GE, 4), ([1, 3], GE, 6)]),
    "dual_mixed": (0, [3, 2], [([1, 1], LE, 4), ([1, 0], GE, 1), ([0, 1], GE, 1)]),
    "dual_eq": (0, [1, 2], [([1, 1], EQ, 4), ([1, 0], LE, 3)]),
    "dual_infeasible": (0, [1, 1], [([1, 1], LE, 2), ([1, 1], GE, 5)]),
    "dual_ge_child_becomes_le": (0, [5, 4], [([6, 4], LE, 24), ([1, 2], LE, 6), ([0, 1], GE, 2)]),
}

IP_CASES = {
    "ip_floor_path": (0, [5, 4], [([6, 4], LE, 24), ([1, 2], LE, 6)]),
    "ip_integral_root": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)]),
    "ip_with_ge_root": (0, [1, 1], [([1, 1], LE, 4), ([1, 0], GE, 1)]),
    "ip_three_vars": (0, [7, 3, 4], [([3, 2, 5], LE, 17), ([4, 1, 2], LE, 11), ([1, 3, 1], LE, 9)]),
}

KNAP_CASES = {
    "knap_classic": ([60, 100, 120], [10, 20, 30], 50),
    "knap_ties": ([10, 10, 10, 10], [5, 5, 5, 5], 12),
    "knap_all_fit": ([3, 4, 5], [1, 2, 3], 10),
    "knap_zero_weight": ([5, 4, 3, 7], [0, 2, 3, 4], 5),
    "knap_fractional_data": ([10.5, 7.25, 3.125, 8.75, 6.5], [3.5, 2.25, 1.5, 4.75, 2.5], 7.3),
    "knap_nothing_fits": ([5, 6], [10, 12], 4),
}

CUT_CASES = {
    # CuttingPlane: the reference's cut never removes the current vertex (it is read one row too low
    # and bounds only decision-variable columns from above), so a fractional root runs all 50 rounds
    "cut_classic_incomplete": (0, [5, 8], [([1, 1], LE, 6), ([5, 9], LE, 45)]),
    "cut_integral_root": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)]),
    "cut_from_z_row": (0, [1, 1], [([2, 2], LE, 3)]),
    "cut_three_vars": (0, [7, 3, 4], [([3, 2, 5], LE, 17), ([4, 1, 2], LE, 11), ([1, 3, 1], LE, 9)]),
    "cut_min_sense": (1, [-3, -2], [([2, 1], LE, 5), ([1, 3], LE, 7)]),
    "cut_eq_row": (0, [2, 1], [([4, 2], EQ, 7), ([1, 0], LE, 3)]),
    "cut_ge_error": (0, [1, 1], [([1, 1], LE, 4), ([1, 0], GE, 1)]),
}

REV_CASES = {
    # RevisedPrimalSimplex: B^-1 recomputed by Gauss-Jordan each iteration, ratio margin 1e-12
    "rev_wyndor": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)], 10000),
    "rev_min_negative_costs": (1, [-2, -3], [([1, 1], LE, 4), ([1, 3], LE, 6)], 10000),
    "rev_unbounded": (0, [1, 1], [([1, -1], LE, 1)], 10000),
    "rev_degenerate_tie": (0, [2, 3], [([1, 1], LE, 4), ([1, 3], LE, 6), ([0, 1], LE, 2)], 10000),
    "rev_three_vars": (0, [7, 3, 4], [([3, 2, 5], LE, 17), ([4, 1, 2], LE, 11), ([1, 3, 1], LE, 9)], 10000),
    "rev_klee_minty3": (0, [100, 10, 1], [([1, 0, 0], LE, 1), ([20, 1, 0], LE, 100), ([200, 20, 1], LE, 10000)], 10000),
    "rev_needs_row_swaps": (0, [2, 1, 3], [([0, 2, 1], LE, 10), ([3, 0, 1], LE, 12), ([1, 4, 0], LE, 8), ([1, 1, 1], LE, 7)], 10000),
    "rev_ge_row": (0, [1, 1], [([1, 1], LE, 4), ([1, 0], GE, 1)], 10000),
    "rev_neg_rhs": (0, [1, 1], [([1, 1], LE, -4)], 10000),
    "rev_iter_limit": (0, [3, 5], [([1, 0], LE, 4), ([0, 2], LE, 12), ([3, 2], LE, 18)], 1),
}


def hx(v):
    return float(v).hex()


def main():
    out = {"lp": {}, "dual": {}, "ip": {}, "knap": {}, "cut": {}, "rev": {}}
    for name, (sense, c, rows, mi) in LP_CASES.items():
        A = [[float(v) for v in r[0]] for r in rows]
        rel = [r[1] for r in rows]
        b = [float(r[2]) for r in rows]
        case = dict(sense=sense, c=[hx(v) for v in c], A=[[hx(v) for v in r] for r in A], rel=rel, b=[hx(v) for v in b],
                    max_iterations=mi)
        try:
            r = pyref.primal(A, b, [float(v) for v in c], rel, sense, mi, history=True)
            case.update(status=r["status"], pivots=[list(p) for p in r["pivots"]], basis=r["basis"],
                        x=[hx(v) for v in r["x"]], z=hx(r["z"]), tableau=[[hx(v) for v in row] for row in r["tableau"]],
                        n_history=len(r["history"]))
        except pyref.SolveError as e:
            case.update(status=e.code)
        out["lp"][name] = case
    for name, (sense, c, rows) in DUAL_CASES.items():
        A = [[float(v) for v in r[0]] for r in rows]
        rel = [r[1] for r in rows]
        b = [float(r[2]) for r in rows]
        r = pyref.dual(A, b, [float(v) for v in c], rel, sense)
        out["dual"][name] = dict(sense=sense, c=[hx(v) for v in c], A=[[hx(v) for v in r_] for r_ in A], rel=rel,
                                 b=[hx(v) for v in b], status=r["status"], silent=r["silent"],
                                 pivots=[list(p) for p in r["pivots"]], basis=r["basis"], x=[hx(v) for v in r["x"]],
                                 z=hx(r["z"]), tableau=[[hx(v) for v in row] for row in r["tableau"]])
    for name, (sense, c, rows) in IP_CASES.items():
        A = [[float(v) for v in r[0]] for r in rows]
        rel = [r[1] for r in rows]
        b = [float(r[2]) for r in rows]
        found, best, bx, nodes = pyref.bnb(A, b, [float(v) for v in c], rel, sense)
        out["ip"][name] = dict(sense=sense, c=[hx(v) for v in c], A=[[hx(v) for v in r_] for r_ in A], rel=rel,
                               b=[hx(v) for v in b], found=found, best_z=hx(best) if found else None,
                               best_x=[hx(v) for v in bx] if found else None, node_outcomes=nodes)
    for name, (p, w, cap) in KNAP_CASES.items():
        found, best, bx, evals, pops, order = pyref.knapsack([float(v) for v in p], [float(v) for v in w], float(cap))
        out["knap"][name] = dict(p=[hx(v) for v in p], w=[hx(v) for v in w], cap=hx(cap), found=found,
                                 best=hx(best) if found else None, best_x=bx, pops=pops, rank_order=order,
                                 evals=[[hx(e[0]), hx(e[1]), e[2], e[3]] for e in evals])
    for name, (sense, c, rows) in CUT_CASES.items():
        A = [[float(v) for v in r[0]] for r in rows]
        rel = [r[1] for r in rows]
        b = [float(r[2]) for r in rows]
        r = pyref.cutting_plane(A, b, [float(v) for v in c], rel, sense)
        case = dict(sense=sense, c=[hx(v) for v in c], A=[[hx(v) for v in r_] for r_ in A], rel=rel,
                    b=[hx(v) for v in b], end=r["end"], rounds=[list(t) for t in r["rounds"]],
                    cuts=[dict(frac_var=fv, row=row, a=[hx(v) for v in a], b=hx(f0)) for fv, row, a, f0 in r["cuts"]])
        if r["end"] == 0:
            case.update(x=[hx(v) for v in r["x"]], z=hx(r["z"]), basis=r["basis"],
                        tableau=[[hx(v) for v in row] for row in r["tableau"]])
        out["cut"][name] = case
    for name, (sense, c, rows, mi) in REV_CASES.items():
        A = [[float(v) for v in r[0]] for r in rows]
        rel = [r[1] for r in rows]
        b = [float(r[2]) for r in rows]
        case = dict(sense=sense, c=[hx(v) for v in c], A=[[hx(v) for v in r_] for r_ in A], rel=rel,
                    b=[hx(v) for v in b], max_iterations=mi)
        try:
            r = pyref.revised(A, b, [float(v) for v in c], rel, sense, mi)
            case.update(status=r["status"], pivots=[[e, l, hx(t)] for e, l, t in r["pivots"]], basis=r["basis"],
                        xB=[hx(v) for v in r["xB"]], Binv=[[hx(v) for v in row] for row in r["Binv"]],
                        x=[hx(v) for v in r["x"]], z=hx(r["z"]))
        except pyref.SolveError as e:
            case.update(status=e.code)
        out["rev"][name] = case
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote kat.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
