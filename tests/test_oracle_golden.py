"""CPU suite: the C++ oracle against the golden known-answer cases (tests/golden/kat.json, made
by the independent pure-Python restatement tests/pyref.py) and against hand-derived values."""
import numpy as np
import pytest

from conftest import assert_bits_equal, case_arrays, unhex

from linear_programming_solver_lpr381_b200 import workloads


def test_wyndor_hand_values(orc):
    # pivots (1,1),(0,2); basis [2,1,0]; x=(2,6); z=36; final rows as fractions
    p = orc.parse_text(workloads.WYNDOR_TEXT)
    r = orc.primal_solve(p["A"], p["b"], p["c"], p["rel"], p["sense"])
    assert r["status"] == 0 and r["pivots"].tolist() == [[1, 1], [0, 2]]
    assert r["basis"].tolist() == [2, 1, 0]
    assert r["x"].tolist() == [2.0, 6.0] and r["z"] == 36.0
    third = 1.0 / 3.0
    T = r["tableau"]
    assert np.allclose(T, [[0, 0, 1, third, -third, 2], [0, 1, 0, 0.5, 0, 6], [1, 0, 0, -third, third, 2],
                           [0, 0, 0, 1.5, 1, 36]], atol=1e-15)


@pytest.mark.parametrize("name", ["wyndor", "unbounded", "degenerate_tie", "near_tie_margin", "near_tie_margin_rev",
                                  "eq_expansion", "min_trivial", "min_negative_costs", "ge_row", "neg_rhs",
                                  "neg_rhs_tolerated", "iter_limit_hit", "iter_limit_ok", "zero_cost_negzero",
                                  "klee_minty3"])
def test_primal_kat(orc, kat, name):
    case = kat["lp"][name]
    A, b, c, rel = case_arrays(case)
    r = orc.primal_solve(A, b, c, rel, case["sense"], max_iterations=case["max_iterations"], history=True)
    assert r["status"] == case["status"]
    if case["status"] < 0:
        assert r["rc"] == case["status"]
        return
    assert r["pivots"].tolist() == case["pivots"]
    assert r["basis"].tolist() == case["basis"]
    assert_bits_equal(r["x"], unhex(case["x"]), "x")
    assert_bits_equal([r["z"]], [unhex(case["z"])], "z")
    assert_bits_equal(r["tableau"], unhex(case["tableau"]), "tableau")
    assert len(r["history"]) == case["n_history"]


def test_margin_rule_is_not_argmin(orc, kat):
    # F6: rows 0 and 1 have ratios 1.0000000005 and 1.0; the earlier row wins although it is larger
    case = kat["lp"]["near_tie_margin"]
    assert case["pivots"][0] == [0, 0]
    assert kat["lp"]["near_tie_margin_rev"]["pivots"][0] == [0, 0]


@pytest.mark.parametrize("name", ["dual_ge", "dual_mixed", "dual_eq", "dual_infeasible", "dual_ge_child_becomes_le"])
def test_dual_kat(orc, kat, name):
    case = kat["dual"][name]
    A, b, c, rel = case_arrays(case)
    r = orc.dual_solve(A, b, c, rel, case["sense"])
    assert r["status"] == case["status"] and r["silent"] == case["silent"]
    assert r["pivots"].tolist() == case["pivots"]
    assert r["basis"].tolist() == case["basis"]
    assert_bits_equal(r["x"], unhex(case["x"]), "x")
    assert_bits_equal(r["tableau"], unhex(case["tableau"]), "tableau")


@pytest.mark.parametrize("name", ["ip_floor_path", "ip_integral_root", "ip_with_ge_root", "ip_three_vars"])
def test_bnb_kat(orc, kat, name):
    case = kat["ip"][name]
    A, b, c, rel = case_arrays(case)
    r = orc.bnb_simplex(A, b, c, rel, case["sense"])
    assert r["found"] == case["found"]
    assert r["outcome"].tolist() == case["node_outcomes"]
    if case["found"]:
        assert_bits_equal([r["best_z"]], [unhex(case["best_z"])], "best_z")
        assert_bits_equal(r["best_x"], unhex(case["best_x"]), "best_x")


def test_bnb_floor_path_quirk(kat):
    # '>=' children are rejected, so the reference reports 19, not the true optimum 20
    case = kat["ip"]["ip_floor_path"]
    assert unhex(case["best_z"]) == 19.0 and case["node_outcomes"] == [5, 5, 1, 5, 1, 4]


@pytest.mark.parametrize("name", ["knap_classic", "knap_ties", "knap_all_fit", "knap_zero_weight",
                                  "knap_fractional_data", "knap_nothing_fits"])
def test_knapsack_kat(orc, kat, name):
    case = kat["knap"][name]
    r = orc.knapsack(unhex(case["p"]), unhex(case["w"]), unhex(case["cap"]), eval_cap=4096)
    assert r["found"] == case["found"] and r["n_pops"] == case["pops"]
    assert r["best_x"].tolist() == case["best_x"]
    if case["found"]:
        assert_bits_equal([r["best"]], [unhex(case["best"])], "best")
    assert r["n_evals"] == len(case["evals"])
    assert_bits_equal(r["bound"], [unhex(e[0]) for e in case["evals"]], "bounds")
    assert_bits_equal(r["weight"], [unhex(e[1]) for e in case["evals"]], "weights")
    assert r["frac"].tolist() == [e[2] for e in case["evals"]]
    assert r["decision"].tolist() == [e[3] for e in case["evals"]]


def test_oracle_matches_pyref_on_random_small(orc):
    import pyref
    rng = np.random.default_rng(5)
    for t in range(40):
        m, n = int(rng.integers(2, 7)), int(rng.integers(2, 8))
        A = rng.integers(-3, 10, size=(m, n)).astype(float)
        b = rng.integers(0, 30, size=m).astype(float)
        c = rng.integers(-4, 10, size=n).astype(float)
        rel = rng.choice([0, 0, 0, 2], size=m).astype(np.int32)
        sense = int(rng.integers(0, 2))
        want = pyref.primal(A.tolist(), b.tolist(), c.tolist(), rel.tolist(), sense, 200)
        got = orc.primal_solve(A, b, c, rel, sense, max_iterations=200)
        assert got["status"] == want["status"]
        assert got["pivots"].tolist() == [list(p) for p in want["pivots"]]
        assert_bits_equal(got["tableau"], want["tableau"], f"tableau case {t}")
    for t in range(20):
        m, n = int(rng.integers(2, 6)), int(rng.integers(2, 6))
        A = rng.integers(0, 9, size=(m, n)).astype(float)
        b = rng.integers(1, 25, size=m).astype(float)
        c = rng.integers(1, 9, size=n).astype(float)
        rel = rng.choice([0, 1, 2], size=m).astype(np.int32)
        want = pyref.dual(A.tolist(), b.tolist(), c.tolist(), rel.tolist(), 0)
        got = orc.dual_solve(A, b, c, rel, 0)
        assert got["status"] == want["status"] and got["silent"] == want["silent"]
        assert got["pivots"].tolist() == [list(p) for p in want["pivots"]]
        assert_bits_equal(got["tableau"], want["tableau"], f"dual tableau case {t}")


def test_oracle_knapsack_matches_pyref_random(orc):
    import pyref
    rng = np.random.default_rng(9)
    for t in range(15):
        n = int(rng.integers(3, 14))
        w = np.round(rng.random(n) * 20 + 0.5, 2)
        p = np.round(rng.random(n) * 30 + 0.5, 2)
        cap = float(np.round(w.sum() * 0.45, 2))
        found, best, bx, evals, pops, order = pyref.knapsack(p.tolist(), w.tolist(), cap)
        r = orc.knapsack(p, w, cap, eval_cap=1 << 16)
        assert r["found"] == found and r["n_pops"] == pops and r["best_x"].tolist() == bx
        assert_bits_equal(r["bound"], [e[0] for e in evals], "bounds")
        assert r["decision"].tolist() == [e[3] for e in evals]


def test_oracle_bnb_matches_pyref_random(orc):
    import pyref
    rng = np.random.default_rng(21)
    for t in range(12):
        m, n = int(rng.integers(2, 5)), int(rng.integers(2, 6))
        A = rng.integers(1, 12, size=(m, n)).astype(float)
        b = rng.integers(10, 60, size=m).astype(float)
        c = rng.integers(1, 15, size=n).astype(float)
        found, best, bx, nodes = pyref.bnb(A.tolist(), b.tolist(), c.tolist())
        r = orc.bnb_simplex(A, b, c)
        assert r["found"] == found and r["outcome"].tolist() == nodes
        if found:
            assert_bits_equal([r["best_z"]], [best], "best")
            assert_bits_equal(r["best_x"], bx, "best_x")


CUT_NAMES = ["cut_classic_incomplete", "cut_integral_root", "cut_from_z_row", "cut_three_vars", "cut_min_sense",
             "cut_eq_row", "cut_ge_error"]


@pytest.mark.parametrize("name", CUT_NAMES)
def test_cutting_plane_kat(orc, kat, name):
    # CuttingPlane.cs:13-164: rounds, the row each cut is read from, and every cut coefficient
    case = kat["cut"][name]
    A, b, c, rel = case_arrays(case)
    r = orc.solve_text(workloads.lp_to_text(A, b, c, rel, case["sense"]), "cutting plane")
    assert r["code"] == 0 and r["cut_end"] == case["end"]
    assert len(r["cuts"]) == len(case["cuts"])
    for got, want in zip(r["cuts"], case["cuts"]):
        assert (got["frac_var"], got["row"]) == (want["frac_var"], want["row"])
        assert_bits_equal(got["a"], unhex(want["a"]), "cut a")
        assert_bits_equal([got["b"]], [unhex(want["b"])], "cut b")
    if case["end"] == 0:
        assert_bits_equal(r["tableau"], unhex(case["tableau"]), "tableau")
        assert_bits_equal(r["x"], unhex(case["x"]), "x")
        assert r["summary"].startswith("Status: OPTIMAL INTEGER\nz* = ")
    elif case["end"] == 1:
        assert r["summary"] == "Status: INCOMPLETE" and r["tableau"] is None
        assert r["report"].count("Added Gomory cut: ") == 50
    else:
        assert r["summary"].startswith("Error: Constraint contains '>=' sign.")


def test_cutting_plane_reads_the_row_below(kat):
    # the quirk: with one constraint the fractional variable sits in row 0, so the cut comes from
    # tableau row 1 = the objective row, whose decision-variable entries are 0 -> the empty cut
    case = kat["cut"]["cut_from_z_row"]
    assert all(k["row"] == 1 and unhex(k["a"]) == [0.0, 0.0] and unhex(k["b"]) == 0.5 for k in case["cuts"])


REV_NAMES = ["rev_wyndor", "rev_min_negative_costs", "rev_unbounded", "rev_degenerate_tie", "rev_three_vars",
             "rev_klee_minty3", "rev_needs_row_swaps", "rev_ge_row", "rev_neg_rhs", "rev_iter_limit"]


@pytest.mark.parametrize("name", REV_NAMES)
def test_revised_kat(orc, kat, name):
    # RevisedPrimalSimplex.cs:17-145: pivots (entering column, leaving row, theta), final Bidx, x_B, B^-1
    case = kat["rev"][name]
    A, b, c, rel = case_arrays(case)
    r = orc.revised_solve(A, b, c, rel, case["sense"], max_iterations=case["max_iterations"])
    assert r["status"] == case["status"]
    if case["status"] in (-10, -11):
        return
    assert [[int(e), int(l)] for e, l in zip(r["enter"], r["leave"])] == [p[:2] for p in case["pivots"]]
    assert_bits_equal(r["theta"], [unhex(p[2]) for p in case["pivots"]], "theta")
    assert r["basis"].tolist() == case["basis"]
    assert_bits_equal(r["xB"], unhex(case["xB"]), "xB")
    assert_bits_equal(r["Binv"], unhex(case["Binv"]), "Binv")
    if case["status"] >= 0:
        assert_bits_equal(r["x"], unhex(case["x"]), "x")
        assert_bits_equal([r["z"]], [unhex(case["z"])], "z")


def test_revised_matches_pyref_on_random_small(orc):
    import pyref
    rng = np.random.default_rng(31)
    for t in range(40):
        m, n = int(rng.integers(1, 9)), int(rng.integers(1, 10))
        A = rng.integers(-2, 9, size=(m, n)).astype(float)
        b = rng.integers(0, 30, size=m).astype(float)
        c = rng.integers(-3, 9, size=n).astype(float)
        sense = int(rng.integers(0, 2))
        try:
            want = pyref.revised(A.tolist(), b.tolist(), c.tolist(), None, sense, 200)
        except pyref.SolveError as e:
            assert orc.revised_solve(A, b, c, None, sense, max_iterations=200)["status"] == e.code
            continue
        got = orc.revised_solve(A, b, c, None, sense, max_iterations=200)
        assert got["status"] == want["status"], t
        assert [[int(e), int(l)] for e, l in zip(got["enter"], got["leave"])] == [[e, l] for e, l, _ in want["pivots"]], t
        assert_bits_equal(got["xB"], want["xB"], f"xB {t}")
        assert_bits_equal(got["Binv"], want["Binv"], f"Binv {t}")
