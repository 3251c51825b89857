"""GPU parity for Mode B, the pooled-tree Branch & Bound (lpx_bnb_pooled) — NOT the reference's tree.

Checked two ways: node for node against the oracle's statement of the same search (ids, outcomes, dual
pivots, every z bit, the incumbent), and — for what the tree is FOR — against an independent MILP solver
(scipy's HiGHS) for the optimal value."""
import numpy as np
import pytest

from conftest import assert_bits_equal

from linear_programming_solver_lpr381_b200 import workloads

pytestmark = pytest.mark.gpu


def compare_pooled(got, want, what):
    assert want["rc"] == 0, what
    assert got["found"] == want["found"], what
    assert got["n_nodes"] == want["n_nodes"], what
    assert got["rounds"] == want["rounds"], what
    assert got["node_id"].tolist() == want["node_id"].tolist(), what
    assert got["outcome"].tolist() == want["outcome"].tolist(), what
    assert got["pivots"].tolist() == want["pivots"].tolist(), what
    assert got["total_pivots"] == want["total_pivots"], what
    assert_bits_equal(got["z"], want["z"], what + " node z")
    if want["found"]:
        assert_bits_equal([got["best_z"]], [want["best_z"]], what + " best_z")
        assert_bits_equal(got["best_x"], want["best_x"], what + " best_x")


@pytest.mark.parametrize("batch", [1, 3, 16])
def test_pooled_random_small_vs_oracle_and_milp(lpx, orc, batch):
    from scipy.optimize import Bounds, LinearConstraint, milp
    rng = np.random.default_rng(3)
    for t in range(14):
        m, n = int(rng.integers(3, 9)), int(rng.integers(3, 10))
        A = rng.integers(1, 12, size=(m, n)).astype(float)
        b = rng.integers(3 * n, 12 * n, size=m).astype(float)
        c = rng.integers(1, 15, size=n).astype(float)
        want = orc.bnb_pooled(A, b, c, batch=batch)
        got = lpx.bnb_pooled(A, b, c, batch=batch)
        compare_pooled(got, want, f"case {t} batch {batch}")
        ref = milp(-c, constraints=LinearConstraint(A, ub=b), integrality=np.ones(n), bounds=Bounds(0, np.inf))
        assert got["found"] and abs(got["best_z"] - (-ref.fun)) < 1e-6, (t, got["best_z"], -ref.fun)


@pytest.mark.parametrize("batch", [1, 64, 512])
def test_pooled_c4_instance(lpx, orc, batch):
    """BASELINE config 4, one 60 x 120 IP: the whole pooled tree (thousands of warm-started nodes, node tableaux
    that outgrow one SM's shared memory on the way down) against the oracle."""
    A, b, c = workloads.ip_c4(seed=11)
    want = orc.bnb_pooled(A, b, c, batch=batch)
    got = lpx.bnb_pooled(A, b, c, batch=batch)
    compare_pooled(got, want, f"C4 seed 11 batch {batch}")
    # the reference-exact tree (a floor path) stops at a worse incumbent: that is the point of Mode B
    ref_tree = lpx.bnb_simplex(A, b, c)
    assert got["best_z"] >= ref_tree["best_z"] - 1e-9


def test_pooled_min_sense_and_rejections(lpx, orc):
    from linear_programming_solver_lpr381_b200 import _ffi as F
    A = np.array([[2.0, 3.0, 1.0], [4.0, 1.0, 2.0], [3.0, 4.0, 2.0]])
    b = np.array([5.5, 11.0, 8.0])
    c = np.array([-5.0, -4.0, -3.0])
    compare_pooled(lpx.bnb_pooled(A, b, c, sense=1, batch=4), orc.bnb_pooled(A, b, c, sense=1, batch=4), "min sense")
    for rel, bb in ((np.array([0, 1, 0], dtype=np.int32), b), (None, np.array([5.0, -1.0, 8.0]))):
        with pytest.raises(F.LpxError) as ei:
            lpx.bnb_pooled(A, bb, c, rel=rel)
        assert ei.value.code == F.E_BAD_ARGS
