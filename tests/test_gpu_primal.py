"""GPU parity: Primal / Dual Simplex through the C ABI against the oracle and the golden cases.
Bit-exact: status, pivot sequence, basis, every double of x, z and the tableau (incl. -0.0)."""
import numpy as np
import pytest

from conftest import assert_bits_equal, case_arrays, unhex

from linear_programming_solver_lpr381_b200 import _ffi as F
from linear_programming_solver_lpr381_b200 import workloads

pytestmark = pytest.mark.gpu

KERNELS = [F.KERNEL_AUTO, F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_STREAM, F.KERNEL_CTA_CLUSTER]
LP_NAMES = ["wyndor", "unbounded", "degenerate_tie", "near_tie_margin", "near_tie_margin_rev", "eq_expansion",
            "min_trivial", "min_negative_costs", "ge_row", "neg_rhs", "neg_rhs_tolerated", "iter_limit_hit",
            "iter_limit_ok", "zero_cost_negzero", "klee_minty3"]


def compare_primal(got, want, what):
    assert got["status"] == want["status"], what
    if want["status"] < 0 and want["status"] != F.S_ITER_LIMIT:
        return
    assert got["n_pivots"] == want["n_pivots"], what
    assert got["pivots"].tolist() == want["pivots"].tolist(), what
    if want["status"] == F.S_ITER_LIMIT:
        return
    assert got["basis"].tolist() == want["basis"].tolist(), what
    assert_bits_equal(got["x"], want["x"], what + " x")
    assert_bits_equal([got["z"]], [want["z"]], what + " z")
    assert_bits_equal(got["tableau"], want["tableau"], what + " tableau")


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", LP_NAMES)
def test_primal_kat(lpx, kat, name, kernel):
    case = kat["lp"][name]
    A, b, c, rel = case_arrays(case)
    r = lpx.primal_solve(A, b, c, rel, case["sense"], max_iterations=case["max_iterations"], kernel=kernel)
    assert r["status"] == case["status"]
    if case["status"] < 0:
        return
    assert r["pivots"].tolist() == case["pivots"]
    assert r["basis"].tolist() == case["basis"]
    assert_bits_equal(r["x"], unhex(case["x"]), "x")
    assert_bits_equal([r["z"]], [unhex(case["z"])], "z")
    assert_bits_equal(r["tableau"], unhex(case["tableau"]), "tableau")


@pytest.mark.parametrize("kernel", [F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_CTA_CLUSTER])
def test_history_matches_oracle(lpx, orc, kernel):
    p = orc.parse_text(workloads.WYNDOR_TEXT)
    want = orc.primal_solve(p["A"], p["b"], p["c"], p["rel"], p["sense"], history=True)
    got = lpx.primal_solve(p["A"], p["b"], p["c"], p["rel"], p["sense"], history=8, kernel=kernel)
    assert_bits_equal(got["history"], want["history"], "history")
    A, b, c = workloads.lp_integer(12, 20, 4)
    want = orc.primal_solve(A, b, c, history=True)
    got = lpx.primal_solve(A, b, c, history=want["n_pivots"] + 1, kernel=kernel)
    assert_bits_equal(got["history"], want["history"], "history 12x20")


@pytest.mark.parametrize("name", ["dual_ge", "dual_mixed", "dual_eq", "dual_infeasible", "dual_ge_child_becomes_le"])
@pytest.mark.parametrize("kernel", [F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_CTA_CLUSTER])
def test_dual_kat(lpx, kat, name, kernel):
    case = kat["dual"][name]
    A, b, c, rel = case_arrays(case)
    r = lpx.dual_solve(A, b, c, rel, case["sense"], kernel=kernel)
    assert r["status"] == case["status"] and r["silent"] == case["silent"]
    assert r["pivots"].tolist() == case["pivots"]
    assert r["basis"].tolist() == case["basis"]
    assert_bits_equal(r["x"], unhex(case["x"]), "x")
    assert_bits_equal(r["tableau"], unhex(case["tableau"]), "tableau")


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_small_vs_oracle(lpx, orc, kernel):
    rng = np.random.default_rng(101)
    for t in range(60):
        m, n = int(rng.integers(1, 24)), int(rng.integers(1, 40))
        A = rng.integers(-3, 10, size=(m, n)).astype(float)
        b = rng.integers(0, 40, size=m).astype(float)
        c = rng.integers(-4, 10, size=n).astype(float)
        rel = rng.choice([0, 0, 0, 0, 2], size=m).astype(np.int32)
        sense = int(rng.integers(0, 2))
        want = orc.primal_solve(A, b, c, rel, sense, max_iterations=300)
        got = lpx.primal_solve(A, b, c, rel, sense, max_iterations=300, kernel=kernel)
        compare_primal(got, want, f"case {t} m={m} n={n} kernel={kernel}")


def test_random_dual_vs_oracle(lpx, orc):
    rng = np.random.default_rng(77)
    for t in range(60):
        m, n = int(rng.integers(1, 16)), int(rng.integers(1, 24))
        A = rng.integers(0, 9, size=(m, n)).astype(float)
        b = rng.integers(1, 25, size=m).astype(float)
        c = rng.integers(1, 9, size=n).astype(float)
        rel = rng.choice([0, 1, 2], size=m).astype(np.int32)
        sense = int(rng.integers(0, 2))
        want = orc.dual_solve(A, b, c, rel, sense)
        got = lpx.dual_solve(A, b, c, rel, sense)
        assert got["status"] == want["status"] and got["silent"] == want["silent"], t
        assert got["pivots"].tolist() == want["pivots"].tolist(), t
        assert_bits_equal(got["tableau"], want["tableau"], f"dual tableau {t}")
        assert_bits_equal(got["x"], want["x"], f"dual x {t}")


@pytest.mark.parametrize("shape", [(100, 130), (150, 150), (180, 220), (60, 300)])
def test_cluster_kernel_mid_size_primal_and_dual(lpx, orc, shape):
    """Tableaux too large for one SM's shared memory: AUTO takes the 2- or 4-CTA cluster kernel
    (rows split over the cluster, exchanges through distributed shared memory); same bits as the
    global-memory kernel and the oracle, primal and dual, with and without the iteration history."""
    m, n = shape
    A, b, c = workloads.lp_integer(m, n, 5 + m)
    want = orc.primal_solve(A, b, c, history=False)
    for kernel in (F.KERNEL_CTA_CLUSTER, F.KERNEL_AUTO):
        got = lpx.primal_solve(A, b, c, kernel=kernel)
        compare_primal(got, want, f"{m}x{n} kernel={kernel}")
    rng = np.random.default_rng(m + n)
    rel = rng.choice([0, 0, 1, 2], size=m).astype(np.int32)
    want = orc.dual_solve(A, b, c, rel, 1)
    got = lpx.dual_solve(A, b, c, rel, 1, kernel=F.KERNEL_CTA_CLUSTER)
    assert got["status"] == want["status"] and got["silent"] == want["silent"]
    assert got["pivots"].tolist() == want["pivots"].tolist()
    assert got["basis"].tolist() == want["basis"].tolist()
    assert_bits_equal(got["tableau"], want["tableau"], "dual tableau")
    assert_bits_equal(got["x"], want["x"], "dual x")
    if m <= 100:
        w = orc.primal_solve(A[:, :40], b, c[:40], history=True)
        g = lpx.primal_solve(A[:, :40], b, c[:40], history=w["n_pivots"] + 1, kernel=F.KERNEL_CTA_CLUSTER)
        assert_bits_equal(g["history"], w["history"], "history")


def test_tie_free_decimal_mid_size(lpx, orc):
    for kernel in (F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_STREAM, F.KERNEL_CTA_CLUSTER):
        A, b, c = workloads.lp_decimal(40, 70, 12)
        want = orc.primal_solve(A, b, c)
        got = lpx.primal_solve(A, b, c, kernel=kernel)
        compare_primal(got, want, f"decimal 40x70 kernel={kernel}")


@pytest.mark.parametrize("kind", ["integer", "decimal"])
@pytest.mark.parametrize("kernel", [F.KERNEL_AUTO, F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL])
def test_batched_c2_shape(lpx, orc, kind, kernel):
    A, b, c = workloads.batch_c2(count=192, seed=5, kind=kind)
    want = orc.primal_batch(A, b, c, threads=8, want_tableau=True)
    got = lpx.primal_solve_batched(A, b, c, kernel=kernel)
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["n_pivots"], want["n_pivots"])
    assert got["total_pivots"] == want["total_pivots"]
    assert np.array_equal(got["basis"], want["basis"])
    assert_bits_equal(got["x"], want["x"], "x")
    assert_bits_equal(got["z"], want["z"], "z")
    assert_bits_equal(got["tableau"], want["tableau"], "tableau")


def test_batched_full_c2_against_oracle(lpx, orc):
    """BASELINE config 2 at full size: 4096 LPs of 64 x 128, every output bit compared."""
    A, b, c = workloads.batch_c2(count=4096, seed=1)
    want = orc.primal_batch(A, b, c, threads=8, want_tableau=True)
    got = lpx.primal_solve_batched(A, b, c)
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["n_pivots"], want["n_pivots"])
    assert np.array_equal(got["basis"], want["basis"])
    assert_bits_equal(got["z"], want["z"], "z")
    assert_bits_equal(got["x"], want["x"], "x")
    assert_bits_equal(got["tableau"], want["tableau"], "tableau")
    # size-independent properties: basic columns are exact unit vectors, z = c_B . x_B on integers
    T = got["tableau"]
    k = 17
    for i, col in enumerate(got["basis"][k]):
        unit = np.zeros(65)
        unit[i] = 1.0
        assert np.array_equal(np.abs(T[k][:, col]), unit)


def test_batched_edge_cases(lpx, orc):
    # empty batch
    r = lpx.primal_solve_batched(np.zeros((0, 3, 4)), np.zeros((0, 3)), np.zeros((0, 4)))
    assert r["total_pivots"] == 0 and r["status"].shape == (0,)
    # batch of one, 1 x 1
    r = lpx.primal_solve_batched(np.array([[[2.0]]]), np.array([[6.0]]), np.array([[1.0]]))
    assert r["status"][0] == 0 and r["x"][0, 0] == 3.0 and r["z"][0] == 3.0
    # a batch that mixes optimal, unbounded and negative-RHS instances, with an EQ row pattern
    rng = np.random.default_rng(3)
    A = rng.integers(-2, 8, size=(40, 6, 9)).astype(float)
    b = rng.integers(-1, 30, size=(40, 6)).astype(float)
    c = rng.integers(-3, 9, size=(40, 9)).astype(float)
    rel = np.array([0, 2, 0, 0, 2, 0], dtype=np.int32)
    got = lpx.primal_solve_batched(A, b, c, rel=rel, max_iterations=50)
    for k in range(40):
        want = orc.primal_solve(A[k], b[k], c[k], rel, 0, max_iterations=50)
        assert got["status"][k] == want["status"], k
        if want["status"] >= 0:
            assert got["n_pivots"][k] == want["n_pivots"]
            assert_bits_equal(got["tableau"][k], want["tableau"], f"tableau {k}")
            assert_bits_equal(got["x"][k], want["x"], f"x {k}")


def test_batched_device_pointers(lpx, orc):
    torch = pytest.importorskip("torch")
    A, b, c = workloads.batch_c2(count=128, seed=9)
    dev = torch.device("cuda:0")
    dA, db, dc = (torch.from_numpy(v).to(dev) for v in (A, b, c))
    count, m, n = A.shape
    st = torch.zeros(count, dtype=torch.int32, device=dev)
    npv = torch.zeros(count, dtype=torch.int32, device=dev)
    basis = torch.zeros((count, m), dtype=torch.int32, device=dev)
    x = torch.zeros((count, n), dtype=torch.float64, device=dev)
    z = torch.zeros(count, dtype=torch.float64, device=dev)
    T = torch.zeros((count, m + 1, n + m + 1), dtype=torch.float64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    lpx.primal_solve_batched_dev(count, m, n, 0, dA.data_ptr(), None, db.data_ptr(), dc.data_ptr(), st.data_ptr(),
                                 npv.data_ptr(), basis.data_ptr(), x.data_ptr(), z.data_ptr(), T.data_ptr(),
                                 tot.data_ptr(), stream)
    torch.cuda.synchronize()
    want = orc.primal_batch(A, b, c, threads=8, want_tableau=True)
    assert np.array_equal(st.cpu().numpy(), want["status"])
    assert int(tot.item()) == want["total_pivots"]
    assert_bits_equal(T.cpu().numpy(), want["tableau"], "tableau")
    assert_bits_equal(x.cpu().numpy(), want["x"], "x")


def test_wide_and_tall_shapes(lpx, orc):
    for (m, n, seed) in [(3, 300, 1), (150, 5, 2), (100, 100, 3), (130, 260, 4)]:
        A, b, c = workloads.lp_integer(m, n, seed)
        want = orc.primal_solve(A, b, c)
        for kernel in (F.KERNEL_AUTO, F.KERNEL_CTA_GLOBAL, F.KERNEL_STREAM):
            got = lpx.primal_solve(A, b, c, kernel=kernel)
            compare_primal(got, want, f"{m}x{n} kernel={kernel}")


@pytest.mark.parametrize("reg_variant", [0, 1, 2, 3, 4])
def test_register_resident_kernel(lpx, orc, reg_variant):
    """LPX_KERNEL_CTA_REG on every shape class it serves, against the oracle, bit for bit: the full-tableau
    builds (1, 2, 3), the condensed build (4: non-basic columns only, the full tableau put back together on the
    way out) and the default choice between them (0)."""
    rng = np.random.default_rng(17)
    shapes = [(64, 128), (64, 100), (60, 120), (10, 20), (1, 1), (33, 7), (64, 1), (5, 187), (40, 128), (64, 127)]
    for (m, n) in shapes:
        if reg_variant == 4 and n > 128:  # the condensed build holds n <= 128 columns and says so
            with pytest.raises(F.LpxError):
                lpx.primal_solve_batched(*workloads.batch_c2(count=2, m=m, n=n, seed=1), kernel=F.KERNEL_CTA_REG, reg_variant=4)
            continue
        for kind, sense in (("integer", 0), ("decimal", 0), ("integer", 1)):
            A, b, c = workloads.batch_c2(count=24, m=m, n=n, seed=3 + m + n, kind=kind)
            if sense == 1:
                c = -c
            want = orc.primal_batch(A, b, c if sense == 0 else -c, threads=8, want_tableau=True)
            got = lpx.primal_solve_batched(A, b, c, sense=sense, kernel=F.KERNEL_CTA_REG, reg_variant=reg_variant)
            what = f"{m}x{n} {kind} sense={sense}"
            assert np.array_equal(got["status"], want["status"]), what
            assert np.array_equal(got["n_pivots"], want["n_pivots"]), what
            assert np.array_equal(got["basis"], want["basis"]), what
            assert_bits_equal(got["tableau"], want["tableau"], what + " tableau")
            assert_bits_equal(got["x"], want["x"], what + " x")
            assert_bits_equal(got["z"], want["z"], what + " z")
    # unbounded, negative RHS, iteration limit and degenerate instances in one batch
    A = rng.integers(-3, 8, size=(60, 12, 15)).astype(float)
    b = rng.integers(-1, 30, size=(60, 12)).astype(float)
    c = rng.integers(-3, 9, size=(60, 15)).astype(float)
    A[50:, :, 0] = -1.0  # column 0 enters first (largest cost) and has no positive entry: unbounded
    b[50:] = np.abs(b[50:])
    c[50:, 0] = 9.0
    got = lpx.primal_solve_batched(A, b, c, max_iterations=9, kernel=F.KERNEL_CTA_REG, reg_variant=reg_variant)
    seen = set()
    for k in range(60):
        want = orc.primal_solve(A[k], b[k], c[k], None, 0, max_iterations=9)
        assert got["status"][k] == want["status"], k
        seen.add(want["status"])
        if want["status"] >= 0:
            assert got["n_pivots"][k] == want["n_pivots"]
            assert_bits_equal(got["tableau"][k], want["tableau"], f"tableau {k}")
            assert_bits_equal(got["x"][k], want["x"], f"x {k}")
    assert {0, 1, -2, -3} <= seen
    # shapes it does not serve are refused, not silently rerouted
    with pytest.raises(F.LpxError):
        lpx.primal_solve_batched(np.ones((2, 70, 10)), np.ones((2, 70)), np.ones((2, 10)), kernel=F.KERNEL_CTA_REG)


BEALE_A = np.array([[0.25, -8.0, -1.0, 9.0], [0.5, -12.0, -0.5, 3.0], [0.0, 0.0, 1.0, 0.0]])
BEALE_B = np.array([0.0, 0.0, 1.0])
BEALE_C = np.array([0.75, -20.0, 0.5, -6.0])


@pytest.mark.parametrize("reg_variant", [1, 2, 4])
def test_register_kernel_near_ties_and_signed_zeros(lpx, orc, reg_variant):
    """The certified shortcut of the leaving-row scan must hand ties, near-ties inside the 1e-9
    margin, zero ratios and -0.0 / +0.0 pairs to the exact replay (PrimalSimplex.cs:222-243)."""
    rng = np.random.default_rng(5)
    count, m, n = 96, 16, 6
    A = rng.integers(1, 6, size=(count, m, n)).astype(np.float64)
    b = np.empty((count, m))
    c = rng.integers(1, 9, size=(count, n)).astype(np.float64)
    for k in range(count):
        base = float(rng.integers(2, 9))
        # ratios of column j are b_i / A_ij: make whole groups of rows tie, or differ by less /
        # slightly more than the margin, in shuffled order
        eps = rng.choice([0.0, 2e-10, 6e-10, 9.9e-10, 1.1e-9, 3e-9], size=m)
        b[k] = (base + eps) * A[k, :, 0]
        if k % 3 == 0:
            b[k, rng.integers(0, m, size=3)] = 0.0          # degenerate rows: zero ratios, signed zeros later
        if k % 4 == 0:
            A[k, rng.integers(0, m), :] *= -1.0             # rows that are never eligible in some columns
        rng.shuffle(b[k])
    b = np.abs(b)
    want = orc.primal_batch(A, b, c, threads=8, want_tableau=True)
    got = lpx.primal_solve_batched(A, b, c, kernel=F.KERNEL_CTA_REG, reg_variant=reg_variant)
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["n_pivots"], want["n_pivots"])
    assert np.array_equal(got["basis"], want["basis"])
    assert_bits_equal(got["tableau"], want["tableau"], "tableau")
    assert_bits_equal(got["x"], want["x"], "x")
    # the cases are only worth something if the margin rule really differs from a plain argmin on them
    plain = 0
    for k in range(count):
        T0 = np.zeros((m + 1, n + m + 1))
        T0[:m, :n], T0[:m, n:n + m], T0[:m, -1], T0[m, :n] = A[k], np.eye(m), b[k], -c[k]
        e = int(np.argmin(T0[m, :-1]))
        col = T0[:m, e]
        ratios = np.where(col > 1e-9, T0[:m, -1] / np.where(col > 1e-9, col, 1.0), np.inf)
        first = orc.primal_solve(A[k], b[k], c[k], None, 0)["pivots"]
        if len(first) and int(np.argmin(ratios)) != int(first[0][1]):
            plain += 1
    assert plain > 0


@pytest.mark.parametrize("kernel", KERNELS)
def test_cycling_lp_hits_the_iteration_limit(lpx, orc, kernel):
    """Beale's example cycles under Dantzig + lowest-index ties: the reference throws "Iteration
    limit exceeded." after exactly 10000 pivots (PrimalSimplex.cs:95-96); so must the engine, with
    the same 10000 (entering, leaving) pairs."""
    want = orc.primal_solve(BEALE_A, BEALE_B, BEALE_C, max_iterations=10000)
    assert want["status"] == F.S_ITER_LIMIT and want["n_pivots"] == 10000
    got = lpx.primal_solve(BEALE_A, BEALE_B, BEALE_C, max_iterations=10000, kernel=kernel)
    assert got["status"] == F.S_ITER_LIMIT and got["n_pivots"] == 10000
    assert got["pivots"].tolist() == want["pivots"].tolist()
    assert F.status_message(got["status"]) == "Iteration limit exceeded."


def test_cycling_lp_batched_register_kernel(lpx, orc):
    A = np.repeat(BEALE_A[None], 5, axis=0)
    b = np.repeat(BEALE_B[None], 5, axis=0)
    c = np.repeat(BEALE_C[None], 5, axis=0)
    c[3] = [1.0, 1.0, 1.0, 1.0]  # one well-behaved instance among the cycling ones
    for kernel in (F.KERNEL_CTA_REG, F.KERNEL_CTA_SMEM):
        got = lpx.primal_solve_batched(A, b, c, max_iterations=10000, kernel=kernel)
        for k in range(5):
            want = orc.primal_solve(A[k], b[k], c[k], max_iterations=10000)
            assert got["status"][k] == want["status"] and got["n_pivots"][k] == want["n_pivots"], (kernel, k)


def test_extreme_aspect_ratios(lpx, orc):
    """Tall-thin and short-wide tableaux through the automatic kernel choice."""
    for (m, n, seed) in [(2500, 4, 1), (4, 12000, 2), (700, 900, 3)]:
        A, b, c = workloads.lp_integer(m, n, seed)
        want = orc.primal_solve(A, b, c)
        got = lpx.primal_solve(A, b, c)
        compare_primal(got, want, f"{m}x{n}")


def test_forced_register_kernel_refuses_non_le_rows(lpx):
    """LPX_KERNEL_CTA_REG serves all-'<=' batches only: a '>=' or '=' row is an error of the call, never a
    silent solve of a different problem; an all-LE rel array is the same batch as rel = None."""
    A, b, c = workloads.batch_c2(count=4, m=8, n=12, seed=5)
    for bad in (1, 2):
        rel = np.zeros(8, dtype=np.int32)
        rel[3] = bad
        with pytest.raises(F.LpxError) as ei:
            lpx.primal_solve_batched(A, b, c, rel=rel, kernel=F.KERNEL_CTA_REG)
        assert ei.value.code == F.E_CAPACITY
    g0 = lpx.primal_solve_batched(A, b, c, kernel=F.KERNEL_CTA_REG)
    g1 = lpx.primal_solve_batched(A, b, c, rel=np.zeros(8, dtype=np.int32), kernel=F.KERNEL_CTA_REG)
    assert_bits_equal(g0["tableau"], g1["tableau"], "tableau")


@pytest.mark.parametrize("kernel", [F.KERNEL_AUTO, F.KERNEL_CTA_REG, F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL])
def test_rejected_problems_return_zeroed_outputs(lpx, orc, kernel):
    """Problems the reference rejects before building a tableau (negative RHS, PrimalSimplex.cs:73-76)
    get their status and zero-filled result slots, not leftovers of an earlier solve in reused buffers."""
    A, b, c = workloads.batch_c2(count=6, m=8, n=12, seed=9)
    first = lpx.primal_solve_batched(A, b, c, kernel=kernel)  # fills the cached device buffers
    assert (first["status"] == 0).all()
    b2 = b.copy()
    b2[2, 5] = -3.0
    b2[4, 0] = -1.0
    out = {k: np.full_like(v, 77) for k, v in first.items() if isinstance(v, np.ndarray)}
    got = lpx.primal_solve_batched(A, b2, c, kernel=kernel, out=out)
    for k in range(6):
        if k in (2, 4):
            assert got["status"][k] == F.S_NEG_RHS and got["n_pivots"][k] == 0
            assert not got["tableau"][k].any() and not got["x"][k].any() and got["z"][k] == 0 and not got["basis"][k].any()
        else:
            want = orc.primal_solve(A[k], b2[k], c[k])
            assert_bits_equal(got["tableau"][k], want["tableau"], f"tableau {k}")
