"""GPU: the host layer (C++ mirror of the controllers over liblpx.so) must print exactly what the
oracle's restatement of the C# prints: the updatePivot stream, Report and Summary."""
import os
import subprocess

import numpy as np
import pytest

import host_ffi as H
from conftest import assert_bits_equal, case_arrays, unhex

from linear_programming_solver_lpr381_b200 import workloads

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same_text(orc, text, algorithm):
    want = orc.solve_text(text, algorithm)
    got = H.solve_text(text, algorithm)
    assert (got["code"] != 0) == (want["code"] != 0), (got["error"], want["error"])
    assert got["error"] == want["error"]
    assert got["log"] == want["log"]
    assert got["report"] == want["report"]
    assert got["summary"] == want["summary"]
    assert got["chunks"] == want["chunks"]
    # the bool[,] highlight of every callback chunk, cell by cell (PrimalSimplex.cs:117-119,
    # DualSimplex.cs:65-69,104-106; consumer: Form1.AppendPivotRow, Form1.cs:326-368)
    assert got["chunk_len"] == want["chunk_len"]
    for k, (g, w) in enumerate(zip(got["masks"], want["masks"])):
        assert (g is None) == (w is None), f"chunk {k}: null mask on one side only"
        if g is not None:
            assert g.shape == w.shape and np.array_equal(g, w), f"chunk {k}: highlight cells differ"
    return got


def kat_text(case):
    A, b, c, rel = case_arrays(case)
    return workloads.lp_to_text(A, b, c, rel, case["sense"])


def test_wyndor_primal_text(lpx, orc):
    got = same_text(orc, workloads.WYNDOR_TEXT, "Primal Simplex")
    assert "TABLEAU Iteration 2" in got["log"] and got["summary"].startswith("Status: OPTIMAL\nz* = 36\nx* = [2, 6]")
    assert got["highlighted"] == 2


@pytest.mark.parametrize("name", ["unbounded", "degenerate_tie", "near_tie_margin", "eq_expansion", "min_trivial",
                                  "min_negative_costs", "ge_row", "neg_rhs", "zero_cost_negzero", "klee_minty3"])
@pytest.mark.parametrize("algorithm", ["Primal Simplex", "primal simplex algorithm", "Dual Simplex"])
def test_kat_text(lpx, orc, kat, name, algorithm):
    same_text(orc, kat_text(kat["lp"][name]), algorithm)


@pytest.mark.parametrize("name", ["dual_ge", "dual_mixed", "dual_eq", "dual_infeasible", "dual_ge_child_becomes_le"])
def test_dual_text(lpx, orc, kat, name):
    same_text(orc, kat_text(kat["dual"][name]), "Dual Simplex")


def test_unsupported_and_empty_algorithm(lpx, orc):
    for algo in ("Revised Dual Simplex", "Cutting Plane", "  "):  # no such LPSolver keys (LPSolver.cs:24-33)
        want = orc.solve_text(workloads.WYNDOR_TEXT, algo)
        got = H.solve_text(workloads.WYNDOR_TEXT, algo)
        assert got["error"] == want["error"] != ""


def test_random_lp_text(lpx, orc):
    rng = np.random.default_rng(8)
    for t in range(12):
        m, n = int(rng.integers(2, 8)), int(rng.integers(2, 9))
        A = np.round(rng.normal(size=(m, n)) * 4, 2)
        b = np.round(rng.random(m) * 30, 2)
        c = np.round(rng.normal(size=n) * 5, 2)
        rel = rng.choice([0, 0, 2], size=m)
        text = workloads.lp_to_text(A, b, c, rel, int(rng.integers(0, 2)))
        same_text(orc, text, "Primal Simplex")
        same_text(orc, text, "Dual Simplex")


@pytest.mark.parametrize("name", ["ip_floor_path", "ip_integral_root", "ip_with_ge_root", "ip_three_vars"])
def test_bnb_text(lpx, orc, kat, name):
    got = same_text(orc, kat_text(kat["ip"][name]), "Branch and Bound")
    if name == "ip_floor_path":
        assert "Subproblem 2.4: x1 <= 3 is integer feasible. Updated BestObjective = 19.000" in got["log"]


def test_bnb_random_text(lpx, orc):
    rng = np.random.default_rng(12)
    for t in range(6):
        m, n = int(rng.integers(2, 5)), int(rng.integers(2, 6))
        A = rng.integers(1, 12, size=(m, n)).astype(float)
        b = rng.integers(10, 60, size=m).astype(float)
        c = rng.integers(1, 15, size=n).astype(float)
        same_text(orc, workloads.lp_to_text(A, b, c), "bnb")


def knap_text(p, w, cap):
    return workloads.lp_to_text(np.array([w]), np.array([cap]), np.array(p))


@pytest.mark.parametrize("name", ["knap_classic", "knap_ties", "knap_all_fit", "knap_zero_weight",
                                  "knap_fractional_data", "knap_nothing_fits"])
def test_knapsack_text(lpx, orc, kat, name):
    from conftest import unhex
    case = kat["knap"][name]
    same_text(orc, knap_text(unhex(case["p"]), unhex(case["w"]), unhex(case["cap"])), "knapsack")


def test_knapsack_medium_text(lpx, orc):
    p, w, cap = workloads.knapsack_c5(n=40, seed=2, kind="fractional")
    same_text(orc, knap_text(p, w, cap), "knapsack")


def test_knapsack_requires_one_le_row(lpx, orc):
    for text in ("Max: 1x1 + 2x2\n1x1 + 1x2 <= 3\n1x1 + 0x2 <= 1\n", "Max: 1x1 + 2x2\n1x1 + 1x2 >= 3\n"):
        want, got = orc.solve_text(text, "knapsack"), H.solve_text(text, "knapsack")
        assert got["error"] == want["error"] != ""


@pytest.mark.parametrize("name", ["cut_classic_incomplete", "cut_integral_root", "cut_from_z_row", "cut_three_vars",
                                  "cut_min_sense", "cut_eq_row", "cut_ge_error"])
def test_cutting_plane_text_and_cuts(lpx, orc, kat, name):
    # CuttingPlane over the GPU primal solver: same text as the oracle, the golden cuts bit for bit
    case = kat["cut"][name]
    got = same_text(orc, kat_text(case), "cutting plane")
    assert len(got["cuts"]) == len(case["cuts"])
    for g, w in zip(got["cuts"], case["cuts"]):
        assert_bits_equal(g["a"], unhex(w["a"]), "cut a")
        assert_bits_equal([g["b"]], [unhex(w["b"])], "cut b")
    if case["end"] == 0:
        assert_bits_equal(got["tableau"], unhex(case["tableau"]), "tableau")
        assert_bits_equal(got["x"], unhex(case["x"]), "x")
        assert got["basis"].tolist() == case["basis"]
    else:
        assert got["tableau"] is None


def test_cutting_plane_random_text(lpx, orc):
    rng = np.random.default_rng(77)
    for trial in range(4):
        m, n = int(rng.integers(2, 6)), int(rng.integers(2, 6))
        A = rng.integers(0, 9, size=(m, n)).astype(np.float64)
        b = rng.integers(5, 40, size=m).astype(np.float64) + (0.5 if trial % 2 else 0.0)
        c = rng.integers(1, 9, size=n).astype(np.float64)
        text = workloads.lp_to_text(A, b, c, np.zeros(m, dtype=np.int32), 0)
        got = same_text(orc, text, "cutting plane")
        want = orc.solve_text(text, "cutting plane")
        assert len(got["cuts"]) == len(want["cuts"])
        for g, w in zip(got["cuts"], want["cuts"]):
            assert_bits_equal(g["a"], w["a"], "cut a")
            assert_bits_equal([g["b"]], [w["b"]], "cut b")


def test_controller_entry_point(lpx, orc):
    got = H.solve_text(workloads.WYNDOR_TEXT, "controller")
    want = orc.solve_text(workloads.WYNDOR_TEXT, "Primal Simplex")
    assert got["report"] == want["report"] and got["summary"] == want["summary"] and got["chunks"] == 0


def test_cli_example_input(lpx, orc, tmp_path):
    exe = os.path.join(ROOT, "linear_programming_solver_lpr381_b200", "lpr381")
    example = os.path.join(ROOT, "tests", "golden", "example_input.txt")
    out = subprocess.run([exe, "Primal Simplex", example], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    want = orc.solve_text(open(example).read(), "Primal Simplex")
    assert out.stdout == want["log"] + "\n\nFinal Report:\n" + want["report"] + "\n\nSummary:\n" + want["summary"]
    assert out.stdout == open(os.path.join(ROOT, "tests", "golden", "example_output.txt")).read()
    bad = subprocess.run([exe, "Primal Simplex"], input="Max: 1x1\n1x1 >= 2\n", capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "Constraint contains '>=' sign." in bad.stderr


@pytest.mark.parametrize("algorithm,oracle_key", [("Primal Simplex", "Primal Simplex"), ("Dual Simplex", "Dual Simplex"),
                                                  ("Branch and Bound", "Branch and Bound"),
                                                  ("Branch and Bound Knapsack", "Branch and Bound"),
                                                  ("Cutting Plane", "cutting plane"),
                                                  ("BranchAndBoundKnapsack", "knapsack")])
@pytest.mark.parametrize("crlf", [False, True])
def test_cli_export_layout(lpx, orc, algorithm, oracle_key, crlf):
    """`lpr381 --export`: BtnExport_Click's file (R/Form1.cs:310-314) = WriteLine("Linear Program:"),
    WriteLine(input), WriteLine(), WriteLine("Iterations:"), WriteLine(box) where box is the callback
    stream + "\n\nFinal Report:\n" + Report + "\n\nSummary:\n" + Summary (R/Form1.cs:277-278; those
    "\n" are literals, the WriteLine terminators are Environment.NewLine)."""
    import orc_ffi
    exe = os.path.join(ROOT, "linear_programming_solver_lpr381_b200", "lpr381")
    text = "Max: 60x1 + 100x2 + 120x3\n10x1 + 20x2 + 30x3 <= 50\n" if "napsack" in algorithm else workloads.WYNDOR_TEXT
    nl = "\r\n" if crlf else "\n"
    orc_ffi.lib().orc_set_newline(nl.encode())
    try:
        want = orc.solve_text(text, oracle_key)
    finally:
        orc_ffi.lib().orc_set_newline(b"\n")
    assert want["code"] == 0
    box = want["log"] + "\n\nFinal Report:\n" + want["report"] + "\n\nSummary:\n" + want["summary"]
    expect = "Linear Program:" + nl + text + nl + nl + "Iterations:" + nl + box + nl
    args = [exe, "--export"] + (["--crlf"] if crlf else []) + [algorithm]
    out = subprocess.run(args, input=text.encode(), capture_output=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.decode() == expect


def test_ragged_rows_fail_like_build_tableau(lpx, orc):
    """A constraint with fewer coefficients than the objective parses, then BuildTableau indexes past
    the row (IndexOutOfRangeException, PrimalSimplex.cs:190)."""
    text = "Max: 3x1 + 5x2 + 1x3\n1x1 + 2x2 <= 4\n1x1 + 1x2 + 1x3 <= 5\n"
    for algo in ("Primal Simplex", "Dual Simplex", "Branch and Bound", "knapsack"):
        got = H.solve_text(text, algo)
        assert got["code"] != 0 and got["error"] != "", algo
    assert H.solve_text(text, "Primal Simplex")["error"] == "Index was outside the bounds of the array."


def test_empty_and_blank_inputs(lpx, orc):
    for text in ("", "\n\n", "Max: 3x1\n"):
        want, got = orc.solve_text(text, "Primal Simplex"), H.solve_text(text, "Primal Simplex")
        assert got["error"] == want["error"] == "Input must contain an objective and at least one constraint."


def test_windows_newlines(lpx, orc):
    """Environment.NewLine is "\r\n" where the reference runs; literal "\n" inside format strings stay."""
    import orc_ffi
    texts = [(workloads.WYNDOR_TEXT, a) for a in ("Primal Simplex", "Dual Simplex", "Branch and Bound",
                                                   "Revised Primal Simplex", "cutting plane", "revised cutting plane")]
    texts.append(("Max: 60x1 + 100x2 + 120x3\n10x1 + 20x2 + 30x3 <= 50\n", "knapsack"))
    orc_ffi.lib().orc_set_newline(b"\r\n")
    H.lib().lpr_set_newline(b"\r\n")
    try:
        for text, algo in texts:
            got = same_text(orc, text, algo)
            assert "\r\n" in got["report"] + got["log"], algo
    finally:
        orc_ffi.lib().orc_set_newline(b"\n")
        H.lib().lpr_set_newline(b"\n")
