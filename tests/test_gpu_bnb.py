"""GPU parity: Branch & Bound (simplex) and Branch & Bound Knapsack through the C ABI."""
import numpy as np
import pytest

from conftest import assert_bits_equal, case_arrays, unhex

from linear_programming_solver_lpr381_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ip_floor_path", "ip_integral_root", "ip_with_ge_root", "ip_three_vars"])
def test_bnb_kat(lpx, kat, name):
    case = kat["ip"][name]
    A, b, c, rel = case_arrays(case)
    r = lpx.bnb_simplex(A, b, c, rel, case["sense"], trace=True)
    assert r["found"] == case["found"]
    assert [nd["outcome"] for nd in r["nodes"]] == case["node_outcomes"]
    if case["found"]:
        assert_bits_equal([r["best_z"]], [unhex(case["best_z"])], "best_z")
        assert_bits_equal(r["best_x"], unhex(case["best_x"]), "best_x")


@pytest.mark.parametrize("name", ["ip_floor_path", "ip_integral_root", "ip_with_ge_root", "ip_three_vars"])
def test_bnb_kat_without_callback(lpx, kat, name):
    """The same known-answer cases as ONE tree without a callback: the pipelined driver and the condensed-tableau
    node kernel, down to an integral root and a root that Dual Simplex has to take."""
    case = kat["ip"][name]
    A, b, c, rel = case_arrays(case)
    traced = lpx.bnb_simplex(A, b, c, rel, case["sense"], trace=True)
    r = lpx.bnb_simplex(A, b, c, rel, case["sense"])
    assert r["found"] == case["found"]
    assert r["n_nodes"] == len(case["node_outcomes"]) and r["lp_pivots"] == traced["lp_pivots"]
    if case["found"]:
        assert_bits_equal([r["best_z"]], [unhex(case["best_z"])], "best_z")
        assert_bits_equal(r["best_x"], unhex(case["best_x"]), "best_x")


def test_bnb_single_trees_without_callback(lpx, orc):
    """One tree per call, no callback (1 .. 5 instances per call as well): unbounded, infeasible-by-rounding and
    ordinary instances against the oracle."""
    rng = np.random.default_rng(32)
    for t in range(20):
        m, n = int(rng.integers(2, 7)), int(rng.integers(2, 9))
        A = rng.integers(-2, 12, size=(m, n)).astype(float)
        b = rng.integers(5, 80, size=m).astype(float)
        c = rng.integers(-3, 15, size=n).astype(float)
        want = orc.bnb_simplex(A, b, c, node_cap=1 << 16)
        got = lpx.bnb_simplex(A, b, c)
        assert got["found"] == want["found"] and got["n_nodes"] == want["n_nodes"], t
        assert got["lp_pivots"] == want["total_pivots"], t
        if want["found"]:
            assert_bits_equal([got["best_z"]], [want["best_z"]], f"best_z {t}")
            assert_bits_equal(got["best_x"], want["best_x"], f"best_x {t}")
    for count in (2, 3, 5):
        A = rng.integers(1, 12, size=(count, 5, 7)).astype(float)
        b = rng.integers(10, 80, size=(count, 5)).astype(float)
        c = rng.integers(1, 15, size=(count, 7)).astype(float)
        got = lpx.bnb_simplex_batched(A, b, c)
        _check_batch(got, [orc.bnb_simplex(A[k], b[k], c[k], node_cap=1 << 16) for k in range(count)], f"count {count}")


def compare_bnb(got, want, what):
    assert got["found"] == want["found"], what
    assert got["n_nodes"] == want["n_nodes"], what
    assert [nd["outcome"] for nd in got["nodes"]] == want["outcome"].tolist(), what
    assert [nd["algo"] for nd in got["nodes"]] == want["algo"].tolist(), what
    assert [nd["n_pivots"] for nd in got["nodes"]] == want["pivots"].tolist(), what
    assert [nd["depth"] for nd in got["nodes"]] == want["depth"].tolist(), what
    assert got["lp_pivots"] == want["total_pivots"], what
    zs = [nd["z"] if nd["algo"] == 0 and nd["lp_status"] >= 0 else 0.0 for nd in got["nodes"]]
    assert_bits_equal(zs, want["z"], what + " node z")
    assert [nd["branch_var"] for nd in got["nodes"]] == want["branch_var"].tolist(), what
    if want["found"]:
        assert_bits_equal([got["best_z"]], [want["best_z"]], what + " best_z")
        assert_bits_equal(got["best_x"], want["best_x"], what + " best_x")


def test_bnb_random_small(lpx, orc):
    rng = np.random.default_rng(31)
    for t in range(25):
        m, n = int(rng.integers(2, 7)), int(rng.integers(2, 9))
        A = rng.integers(1, 12, size=(m, n)).astype(float)
        b = rng.integers(10, 80, size=m).astype(float)
        c = rng.integers(1, 15, size=n).astype(float)
        want = orc.bnb_simplex(A, b, c)
        got = lpx.bnb_simplex(A, b, c, trace=True)
        compare_bnb(got, want, f"case {t}")


def test_bnb_c4_instance(lpx, orc):
    """BASELINE config 4: 60 x 120 general IP, every node of the reference's tree."""
    A, b, c = workloads.ip_c4(seed=11)
    want = orc.bnb_simplex(A, b, c)
    got = lpx.bnb_simplex(A, b, c, trace=True)
    compare_bnb(got, want, "C4 seed 11")
    # node tableaux outgrow shared memory on the way down: both kernel variants were exercised
    assert max(nd["rows"] * nd["cols"] for nd in got["nodes"]) * 8 > 227 * 1024


def test_bnb_batched_instances(lpx, orc):
    count = 6
    As, bs, cs = zip(*[workloads.ip_c4(m=20, n=30, seed=40 + k) for k in range(count)])
    got = lpx.bnb_simplex_batched(np.stack(As), np.stack(bs), np.stack(cs))
    for k in range(count):
        want = orc.bnb_simplex(As[k], bs[k], cs[k])
        assert bool(got["found"][k]) == want["found"]
        assert got["n_nodes"][k] == want["n_nodes"]
        assert got["lp_pivots"][k] == want["total_pivots"]
        if want["found"]:
            assert_bits_equal([got["best_z"][k]], [want["best_z"]], f"best_z {k}")
            assert_bits_equal(got["best_x"][k], want["best_x"], f"best_x {k}")


def test_bnb_batched_c4_shape(lpx, orc):
    """BASELINE config 4 as bench.py runs it: >= 32 instances of 60 x 120, so the host commit is
    threaded and each round mixes the shared-memory and the cluster launch groups
    (R/Models/Branch&Bound.cs:128-258).  Every instance against the oracle."""
    count = 36
    As, bs, cs = zip(*[workloads.ip_c4(seed=300 + k) for k in range(count)])
    got = lpx.bnb_simplex_batched(np.stack(As), np.stack(bs), np.stack(cs))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:  # the oracle call releases the GIL
        wants = list(ex.map(lambda k: orc.bnb_simplex(As[k], bs[k], cs[k]), range(count)))
    for k in range(count):
        want = wants[k]
        assert bool(got["found"][k]) == want["found"], k
        assert got["n_nodes"][k] == want["n_nodes"], k
        assert got["lp_pivots"][k] == want["total_pivots"], k
        if want["found"]:
            assert_bits_equal([got["best_z"][k]], [want["best_z"]], f"best_z {k}")
            assert_bits_equal(got["best_x"][k], want["best_x"], f"best_x {k}")


def _check_batch(got, wants, label):
    for k, want in enumerate(wants):
        assert bool(got["found"][k]) == want["found"], (label, k)
        assert got["n_nodes"][k] == want["n_nodes"], (label, k)
        assert got["lp_pivots"][k] == want["total_pivots"], (label, k)
        if want["found"]:
            assert_bits_equal([got["best_z"][k]], [want["best_z"]], f"{label} best_z {k}")
            assert_bits_equal(got["best_x"][k], want["best_x"], f"{label} best_x {k}")


def test_bnb_batched_relations_and_sense(lpx, orc):
    """Batches without a callback (the pipelined driver, condensed-tableau node kernel) whose root is NOT an
    all-'<=' Max problem: '>=' and '=' rows send the root LP through Dual Simplex — silent primal pivots, then
    the dual loop, equality rows expanded to two (DualSimplex.cs:117-158) — and Branch & Bound rejects its
    result (Branch&Bound.cs:60-70); a Min objective negates c (PrimalSimplex.cs:62-63)."""
    rng = np.random.Generator(np.random.PCG64(77))
    count, m, n = 40, 6, 9
    for rel, sense in (([0, 1, 0, 0, 0, 0], 0), ([0, 0, 2, 0, 0, 0], 0), ([1, 1, 2, 0, 0, 1], 1), ([0] * 6, 1)):
        A = rng.integers(1, 12, size=(count, m, n)).astype(float)
        b = rng.integers(3 * n, 12 * n, size=(count, m)).astype(float)
        c = rng.integers(1, 15, size=(count, n)).astype(float)
        got = lpx.bnb_simplex_batched(A, b, c, rel, sense)
        wants = [orc.bnb_simplex(A[k], b[k], c[k], rel, sense, node_cap=1 << 16) for k in range(count)]
        _check_batch(got, wants, f"rel {rel} sense {sense}")


def test_bnb_batched_full_tableau_kernels(lpx, orc, monkeypatch):
    """LPX_BNB_FULL_TABLEAU=1 keeps the pipelined driver on the full-tableau node kernels (what it uses when a
    node's condensed tableau does not fit one SM): same trees, bit for bit."""
    count = 32
    As, bs, cs = zip(*[workloads.ip_c4(m=30, n=50, seed=500 + k) for k in range(count)])
    wants = [orc.bnb_simplex(As[k], bs[k], cs[k], node_cap=1 << 16) for k in range(count)]
    _check_batch(lpx.bnb_simplex_batched(np.stack(As), np.stack(bs), np.stack(cs)), wants, "condensed")
    monkeypatch.setenv("LPX_BNB_FULL_TABLEAU", "1")
    _check_batch(lpx.bnb_simplex_batched(np.stack(As), np.stack(bs), np.stack(cs)), wants, "full tableau")


def test_bnb_node_history(lpx, orc, kat):
    case = kat["ip"]["ip_floor_path"]
    A, b, c, rel = case_arrays(case)
    r = lpx.bnb_simplex(A, b, c, rel, case["sense"], want_history=True)
    root = r["nodes"][0]
    want = orc.primal_solve(A, b, c, rel, case["sense"], history=True)
    assert_bits_equal(root["history"], want["history"], "root history")
    # the ceil child runs Dual Simplex on "x2 >= 2" (flipped twice into "x2 <= 2", DualSimplex.cs:141-153)
    ceil = r["nodes"][2]
    assert ceil["algo"] == 1 and ceil["outcome"] == 1
    A2 = np.vstack([A, [0.0, 1.0]])
    w2 = orc.dual_solve(A2, np.append(b, 2.0), c, np.append(rel, 1).astype(np.int32), case["sense"], history=True)
    assert ceil["n_pivots"] == w2["n_pivots"] and ceil["silent"] == w2["silent"]
    assert_bits_equal(ceil["history"], w2["history"], "dual child history")


def test_bnb_cycling_root_logs_every_pivot(lpx, orc):
    """A root LP that cycles to the iteration limit (Beale): B&B reports "Error: Infeasible"
    (R/Models/Branch&Bound.cs:59-63) and the node record carries all 10 000 pivot pairs and, on request,
    all 10 001 tableaux — consumers index pivots[0 .. n_pivots) (include/lpx.h)."""
    from test_gpu_primal import BEALE_A, BEALE_B, BEALE_C
    want = orc.primal_solve(BEALE_A, BEALE_B, BEALE_C, max_iterations=10000)
    for hist in (False, True):
        r = lpx.bnb_simplex(BEALE_A, BEALE_B, BEALE_C, trace=True, want_history=hist)
        assert not r["found"] and r["n_nodes"] == 1
        root = r["nodes"][0]
        assert root["lp_status"] == -3 and root["outcome"] == 0 and root["n_pivots"] == 10000
        assert root["pivots"].tolist() == want["pivots"].tolist()
        if hist:
            assert root["history"].shape[0] == 10001
    import host_ffi as H
    text = workloads.lp_to_text(BEALE_A, BEALE_B, BEALE_C)
    got, wt = H.solve_text(text, "Branch and Bound"), orc.solve_text(text, "Branch and Bound")
    assert got["log"] == wt["log"] and got["report"] == wt["report"] and got["summary"] == wt["summary"]


# ---- knapsack -----------------------------------------------------------------------------------

def knap_events(got):
    ev = []
    for p in got["pops"]:
        if p["left"] is not None:
            for side in ("left", "right"):
                e = p[side]
                ev.append((e["bound"], e["weight"], e["frac_rank"], e["decision"], p["pop_index"], e["child"], e["var"]))
    return ev


def compare_knap(got, want, what):
    assert got["found"] == want["found"], what
    assert got["n_pops"] == want["n_pops"], what
    assert got["n_evals"] == want["n_evals"], what
    assert got["best_x"].tolist() == want["best_x"].tolist(), what
    if want["found"]:
        assert_bits_equal([got["best"]], [want["best"]], what + " best")
    if "pops" in got:
        ev = knap_events(got)
        assert len(ev) == want["n_evals"] - 1, what
        assert_bits_equal([e[0] for e in ev], want["bound"][1:], what + " bounds")
        assert_bits_equal([e[1] for e in ev], want["weight"][1:], what + " weights")
        assert [e[2] for e in ev] == want["frac"][1:].tolist(), what
        assert [e[3] for e in ev] == want["decision"][1:].tolist(), what
        assert [e[4] for e in ev] == want["parent"][1:].tolist(), what
        assert [e[6] for e in ev] == want["var"][1:].tolist(), what


@pytest.mark.parametrize("name", ["knap_classic", "knap_ties", "knap_all_fit", "knap_zero_weight",
                                  "knap_fractional_data", "knap_nothing_fits"])
def test_knapsack_kat(lpx, orc, kat, name):
    case = kat["knap"][name]
    p, w, cap = unhex(case["p"]), unhex(case["w"]), unhex(case["cap"])
    got = lpx.bnb_knapsack(p, w, cap, trace=True)
    assert got["rank_order"].tolist() == case["rank_order"]
    compare_knap(got, orc.knapsack(p, w, cap, eval_cap=4096), name)


@pytest.mark.parametrize("spec", [(1, 1), (4, 2), (8, 3), (16, 5)])
def test_knapsack_random_any_speculation(lpx, orc, spec):
    rng = np.random.default_rng(55)
    for t in range(10):
        n = int(rng.integers(3, 40))
        w = np.round(rng.random(n) * 20 + 0.5, 2)
        p = np.round(rng.random(n) * 30 + 0.5, 2)
        cap = float(np.round(w.sum() * 0.45, 2))
        want = orc.knapsack(p, w, cap, eval_cap=1 << 18)
        got = lpx.bnb_knapsack(p, w, cap, trace=True, spec_nodes=spec[0], spec_depth=spec[1])
        compare_knap(got, want, f"case {t} spec={spec}")


@pytest.mark.parametrize("kind", ["uncorrelated", "weak", "fractional"])
@pytest.mark.parametrize("sequential", [False, True])
def test_knapsack_c5(lpx, orc, kind, sequential):
    """BASELINE config 5: 2000 items.  Integer data take the warp-parallel exact-sum path unless
    `sequential` forces the ordered-summation path; both must reproduce the reference."""
    p, w, cap = workloads.knapsack_c5(kind=kind)
    want = orc.knapsack(p, w, cap)
    got = lpx.bnb_knapsack(p, w, cap, sequential=sequential)
    compare_knap(got, want, kind)


@pytest.mark.parametrize("warps", [1, 2])
@pytest.mark.parametrize("kind", ["uncorrelated", "fractional"])
def test_knapsack_one_and_two_warps_per_instance(lpx, orc, kind, warps):
    """The search kernel's two builds — one warp per instance (large batches) and a main + helper pair (the
    helper evaluates the right child while the main warp sinks the heap and evaluates the left one) — on a
    single traced instance and on a batch, against the oracle."""
    p, w, cap = workloads.knapsack_c5(n=500, seed=21, kind=kind)
    compare_knap(lpx.bnb_knapsack(p, w, cap, trace=True, warps=warps), orc.knapsack(p, w, cap), f"{kind} warps={warps}")
    ps, ws, caps = zip(*[workloads.knapsack_c5(n=300, seed=200 + k, kind=kind) for k in range(9)])
    got = lpx.bnb_knapsack_batched(np.stack(ps), np.stack(ws), np.array(caps), warps=warps)
    for k in range(9):
        want = orc.knapsack(ps[k], ws[k], caps[k])
        assert got["n_evals"][k] == want["n_evals"] and got["n_pops"][k] == want["n_pops"], (kind, warps, k)
        assert_bits_equal([got["best"][k]], [want["best"]], f"best {k}")
        assert got["best_x"][k].tolist() == want["best_x"].tolist()


def test_knapsack_integer_edge_values(lpx, orc):
    """Zero and repeated weights, zero profits, capacity hit exactly: the exact-sum path."""
    rng = np.random.default_rng(91)
    for t in range(12):
        n = int(rng.integers(4, 70))
        w = rng.integers(0, 9, size=n).astype(float)
        p = rng.integers(0, 12, size=n).astype(float)
        cap = float(rng.integers(1, max(2, int(w.sum()))))
        want = orc.knapsack(p, w, cap, eval_cap=1 << 20)
        for seq in (False, True):
            got = lpx.bnb_knapsack(p, w, cap, trace=True, sequential=seq)
            compare_knap(got, want, f"case {t} sequential={seq}")


def test_knapsack_c5_trace(lpx, orc):
    p, w, cap = workloads.knapsack_c5(n=300, seed=4, kind="fractional")
    want = orc.knapsack(p, w, cap)
    got = lpx.bnb_knapsack(p, w, cap, trace=True)
    compare_knap(got, want, "n=300 fractional trace")


def test_knapsack_batched(lpx, orc):
    ps, ws, caps = zip(*[workloads.knapsack_c5(n=400, seed=100 + k) for k in range(5)])
    got = lpx.bnb_knapsack_batched(np.stack(ps), np.stack(ws), np.array(caps))
    for k in range(5):
        want = orc.knapsack(ps[k], ws[k], caps[k])
        assert got["n_evals"][k] == want["n_evals"] and got["n_pops"][k] == want["n_pops"]
        assert_bits_equal([got["best"][k]], [want["best"]], f"best {k}")
        assert got["best_x"][k].tolist() == want["best_x"].tolist()
