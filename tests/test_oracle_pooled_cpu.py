"""CPU: the Mode B oracle (oracle/orc_pooled.cpp — the pooled tree, NOT the reference's tree) pinned against an
independent MILP solver (scipy / HiGHS) for the optimal value, and against itself across batch sizes (the
optimum must not depend on how many open nodes a round takes)."""
import numpy as np
import pytest

from linear_programming_solver_lpr381_b200 import workloads


def test_pooled_oracle_reaches_the_milp_optimum(orc):
    from scipy.optimize import Bounds, LinearConstraint, milp
    rng = np.random.default_rng(17)
    for t in range(20):
        m, n = int(rng.integers(3, 10)), int(rng.integers(3, 11))
        A = rng.integers(1, 12, size=(m, n)).astype(float)
        b = rng.integers(3 * n, 12 * n, size=m).astype(float)
        c = rng.integers(1, 15, size=n).astype(float)
        ref = milp(-c, constraints=LinearConstraint(A, ub=b), integrality=np.ones(n), bounds=Bounds(0, np.inf))
        for batch in (1, 7, 64):
            r = orc.bnb_pooled(A, b, c, batch=batch)
            assert r["rc"] == 0 and r["found"], (t, batch)
            assert abs(r["best_z"] - (-ref.fun)) < 1e-6, (t, batch, r["best_z"], -ref.fun)
            x = r["best_x"]
            assert np.all(x == np.round(x)) and np.all(x >= 0) and np.all(A @ x <= b + 1e-9)
            assert abs(c @ x - r["best_z"]) < 1e-6


def test_pooled_oracle_beats_or_ties_the_reference_tree(orc):
    """The reference's tree is one floor path: Mode B explores both children, so its incumbent can
    only be at least as good."""
    for seed in (11, 12, 13):
        A, b, c = workloads.ip_c4(m=12, n=20, seed=seed)
        pooled, exact = orc.bnb_pooled(A, b, c, batch=16), orc.bnb_simplex(A, b, c)
        assert pooled["found"]
        if exact["found"]:
            assert pooled["best_z"] >= exact["best_z"] - 1e-9


def test_pooled_oracle_rejects_what_it_does_not_cover(orc):
    A = np.array([[1.0, 2.0], [3.0, 1.0]])
    assert orc.bnb_pooled(A, np.array([4.0, 6.0]), np.array([1.0, 1.0]), rel=np.array([0, 1], dtype=np.int32))["rc"] != 0
    assert orc.bnb_pooled(A, np.array([4.0, -6.0]), np.array([1.0, 1.0]))["rc"] != 0
