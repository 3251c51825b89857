"""GPU parity: Revised Primal Simplex (lpx_revised_solve, R/Models/RevisedPrimalSimplex.cs:17-145)
against the golden cases and the oracle — pivots, theta, the final basis lists, x_B and every double
of the recomputed basis inverse — and the host layer's iteration text against the oracle's."""
import numpy as np
import pytest

import host_ffi as H
from conftest import assert_bits_equal, case_arrays, unhex

from linear_programming_solver_lpr381_b200 import workloads

pytestmark = pytest.mark.gpu

REV_NAMES = ["rev_wyndor", "rev_min_negative_costs", "rev_unbounded", "rev_degenerate_tie", "rev_three_vars",
             "rev_klee_minty3", "rev_needs_row_swaps", "rev_ge_row", "rev_neg_rhs", "rev_iter_limit"]


@pytest.mark.parametrize("name", REV_NAMES)
def test_revised_kat(lpx, kat, name):
    case = kat["rev"][name]
    A, b, c, rel = case_arrays(case)
    r = lpx.revised_solve(A, b, c, rel, case["sense"], max_iterations=case["max_iterations"])
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


def test_revised_random_vs_oracle(lpx, orc):
    rng = np.random.default_rng(8)
    seen = set()
    for t in range(50):
        m, n = int(rng.integers(1, 30)), int(rng.integers(1, 45))
        A = rng.integers(-2, 10, size=(m, n)).astype(float)
        b = rng.integers(0, 40, size=m).astype(float)
        c = rng.integers(-3, 10, size=n).astype(float)
        if t % 5 == 0:
            A = np.round(rng.random((m, n)) * 9 - 1, 3)
            b = np.round(rng.random(m) * 30, 3)
        sense = int(rng.integers(0, 2))
        want = orc.revised_solve(A, b, c, None, sense, max_iterations=150)
        got = lpx.revised_solve(A, b, c, None, sense, max_iterations=150)
        what = f"case {t} {m}x{n} sense={sense}"
        assert got["status"] == want["status"], what
        seen.add(want["status"])
        if want["status"] == -11:
            continue
        assert got["enter"].tolist() == want["enter"].tolist() and got["leave"].tolist() == want["leave"].tolist(), what
        assert_bits_equal(got["theta"], want["theta"], what + " theta")
        assert got["basis"].tolist() == want["basis"].tolist(), what
        assert_bits_equal(got["xB"], want["xB"], what + " xB")
        assert_bits_equal(got["Binv"], want["Binv"], what + " Binv")
        if want["status"] >= 0:
            assert_bits_equal(got["x"], want["x"], what + " x")
    assert {0, 1} <= seen


def test_revised_mid_size(lpx, orc):
    A, b, c = workloads.lp_integer(60, 90, 3)
    want = orc.revised_solve(A, b, c)
    got = lpx.revised_solve(A, b, c)
    assert got["status"] == want["status"] == 0 and got["n_iters"] == want["n_iters"] > 10
    assert got["enter"].tolist() == want["enter"].tolist() and got["leave"].tolist() == want["leave"].tolist()
    assert_bits_equal(got["Binv"], want["Binv"], "Binv")
    assert_bits_equal(got["xB"], want["xB"], "xB")


def same_text(orc, text, algorithm):
    want = orc.solve_text(text, algorithm)
    got = H.solve_text(text, algorithm)
    assert (got["code"] != 0) == (want["code"] != 0), (got["error"], want["error"])
    assert got["error"] == want["error"]
    assert got["log"] == want["log"]
    assert got["report"] == want["report"]
    assert got["summary"] == want["summary"]
    assert got["chunks"] == want["chunks"]
    return got


@pytest.mark.parametrize("name", REV_NAMES)
@pytest.mark.parametrize("algorithm", ["Revised Primal Simplex", "revised primal"])
def test_revised_text(lpx, orc, kat, name, algorithm):
    case = kat["rev"][name]
    if case["max_iterations"] != 10000:
        pytest.skip("the text entry points use the reference's MaxIterations")
    A, b, c, rel = case_arrays(case)
    same_text(orc, workloads.lp_to_text(A, b, c, rel, case["sense"]), algorithm)


def test_revised_random_text(lpx, orc):
    rng = np.random.default_rng(55)
    for t in range(6):
        m, n = int(rng.integers(2, 9)), int(rng.integers(2, 9))
        A = np.round(rng.random((m, n)) * 9, 2)
        b = np.round(rng.random(m) * 40 + 1, 2)
        c = np.round(rng.random(n) * 9, 2)
        same_text(orc, workloads.lp_to_text(A, b, c, np.zeros(m, dtype=np.int32), t % 2), "Revised Primal Simplex")


CPR_TEXTS = [
    "Max: 5x1 + 8x2\n1x1 + 1x2 <= 6\n5x1 + 9x2 <= 45\n",          # two bounds, then integral
    "Max: 3x1 + 5x2\n1x1 + 0x2 <= 4\n0x1 + 2x2 <= 12\n3x1 + 2x2 <= 18\n",   # integral at the root
    "Max: 1x1 + 1x2\n1x1 - 1x2 <= 1\n",                             # unbounded LP: stops
    "Max: 7x1 + 3x2 + 4x3\n3x1 + 2x2 + 5x3 <= 17\n4x1 + 1x2 + 2x3 <= 11\n1x1 + 3x2 + 1x3 <= 9\n",
    "Min: -5x1 - 4x2\n6x1 + 4x2 <= 24\n1x1 + 2x2 <= 6\n",
    "Max: 1x1 + 1x2\n1x1 + 1x2 <= 4\n1x1 + 0x2 >= 1\n",             # '>=' row: the solver's exception passes through
]


@pytest.mark.parametrize("text", CPR_TEXTS)
def test_cutting_plane_revised_text(lpx, orc, text):
    # CuttingPlaneRevised.cs:14-111 over the GPU revised simplex: same text, same bounds added
    got = same_text(orc, text, "revised cutting plane")
    want = orc.solve_text(text, "revised cutting plane")
    assert len(got["cuts"]) == len(want["cuts"])
    for g, w in zip(got["cuts"], want["cuts"]):
        assert_bits_equal(g["a"], w["a"], "cut a")
        assert_bits_equal([g["b"]], [w["b"]], "cut b")
