"""One-off randomized parity sweep on a GPU box (heavier than the test-suite): every kernel family
against the oracle, bit for bit, on shapes and data chosen to stress ties, near-ties inside the
margins, zero pivots rows/columns, degenerate and unbounded instances.

    python tests/fuzz_parity.py [seed] [rounds]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc_ffi as orc  # noqa: E402

from linear_programming_solver_lpr381_b200 import _ffi as F  # noqa: E402
from linear_programming_solver_lpr381_b200 import api  # noqa: E402



def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def gen(rng, m, n, kind):
    if kind == "int":
        A = rng.integers(-3, 10, size=(m, n)).astype(float)
        b = rng.integers(0, 40, size=m).astype(float)
        c = rng.integers(-4, 10, size=n).astype(float)
    elif kind == "tie":
        A = rng.integers(0, 4, size=(m, n)).astype(float)
        b = rng.integers(0, 6, size=m).astype(float) * rng.choice([1.0, 1.0 + 4e-10, 1.0 + 1.2e-9], size=m)
        c = rng.integers(0, 5, size=n).astype(float)
    else:
        A = np.round(rng.random((m, n)) * 10 - 1, 3)
        b = np.round(rng.random(m) * 50, 3)
        c = np.round(rng.random(n) * 10 - 2, 3)
    return A, b, c


def sweep(seed=1, rounds=3, cnt=192, singles=60, revised=30, every=4, threads=16):
    """`rounds` rounds of the sweep; returns the number of mismatches.  The defaults are the heavy
    stand-alone run; tests/test_gpu_fuzz.py runs a bounded fixed-seed slice of the same code."""
    F.check(F.lib().lpx_init(0))
    rng = np.random.default_rng(seed)
    bad = 0
    t0 = time.time()
    for rd in range(rounds):
        # 1. batched register kernel (full-tableau builds and the condensed build) on full and ragged shapes
        for kind in ("int", "tie", "dec"):
            for (m, n) in ((64, 128), (64, 100), (37, 90), (5, 150), (64, 1)):
                A = np.stack([gen(rng, m, n, kind)[0] for _ in range(cnt)])
                b = np.stack([np.abs(gen(rng, m, n, kind)[1]) for _ in range(cnt)])
                c = np.stack([gen(rng, m, n, kind)[2] for _ in range(cnt)])
                want = orc.primal_batch(A, b, c, threads=threads, want_tableau=True, max_iterations=400)
                for rv in (1, 2, 4):
                    if rv == 4 and n > 128:
                        continue  # the condensed build holds n <= 128 columns
                    got = api.primal_solve_batched(A, b, c, max_iterations=400, kernel=F.KERNEL_CTA_REG, reg_variant=rv)
                    ok = (np.array_equal(got["status"], want["status"]) and np.array_equal(got["n_pivots"], want["n_pivots"])
                          and np.array_equal(got["basis"][want["status"] >= 0], want["basis"][want["status"] >= 0])
                          and same(got["tableau"][want["status"] >= 0], want["tableau"][want["status"] >= 0]))
                    if not ok:
                        bad += 1
                        print("MISMATCH reg", kind, m, n, rv, flush=True)
        # 2. single solves through every per-tableau kernel, primal and dual
        for t in range(singles):
            m, n = int(rng.integers(1, 70)), int(rng.integers(1, 90))
            kind = ("int", "tie", "dec")[t % 3]
            A, b, c = gen(rng, m, n, kind)
            rel = rng.choice([0, 0, 0, 2], size=m).astype(np.int32)
            sense = int(rng.integers(0, 2))
            want = orc.primal_solve(A, np.abs(b), c, rel, sense, max_iterations=300)
            for kernel in (F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_CTA_CLUSTER, F.KERNEL_STREAM):
                got = api.primal_solve(A, np.abs(b), c, rel, sense, max_iterations=300, kernel=kernel)
                ok = got["status"] == want["status"]
                if ok and (want["status"] >= 0 or want["status"] == -3):
                    ok = got["pivots"].tolist() == want["pivots"].tolist()
                if ok and want["status"] >= 0:
                    ok = same(got["tableau"], want["tableau"]) and same(got["x"], want["x"])
                if not ok:
                    bad += 1
                    print("MISMATCH primal", kind, m, n, kernel, want["status"], got["status"], flush=True)
            rel = rng.choice([0, 1, 2], size=m).astype(np.int32)
            want = orc.dual_solve(A, b, c, rel, sense)
            for kernel in (F.KERNEL_CTA_SMEM, F.KERNEL_CTA_GLOBAL, F.KERNEL_CTA_CLUSTER):
                got = api.dual_solve(A, b, c, rel, sense, kernel=kernel)
                ok = got["status"] == want["status"]
                if ok and want["status"] >= 0:  # the oracle exports nothing after the iteration-limit exception
                    ok = (got["silent"] == want["silent"] and got["pivots"].tolist() == want["pivots"].tolist()
                          and same(got["tableau"], want["tableau"]))
                if not ok:
                    bad += 1
                    gp, wp = got["pivots"].tolist(), want["pivots"].tolist()
                    first = next((k for k in range(min(len(gp), len(wp))) if gp[k] != wp[k]), -1)
                    print("MISMATCH dual", kind, m, n, kernel, "status", got["status"], want["status"], "silent",
                          got["silent"], want["silent"], "npiv", got["n_pivots"], want["n_pivots"], "first diff", first,
                          gp[first] if first >= 0 else None, wp[first] if first >= 0 else None,
                          "tableau same" if same(got["tableau"], want["tableau"]) else "tableau differs", flush=True)
        # 3. mid-size tableaux that need a cluster
        for (m, n) in ((120, 140), (170, 230)):
            A, b, c = gen(rng, m, n, "int")
            A = np.abs(A) + 1
            want = orc.primal_solve(A, np.abs(b) + n, c)
            got = api.primal_solve(A, np.abs(b) + n, c)
            if not (got["status"] == want["status"] and got["pivots"].tolist() == want["pivots"].tolist()
                    and same(got["tableau"], want["tableau"])):
                bad += 1
                print("MISMATCH cluster", m, n, flush=True)
        # 4. revised simplex
        for t in range(revised):
            m, n = int(rng.integers(1, 40)), int(rng.integers(1, 50))
            A, b, c = gen(rng, m, n, ("int", "tie", "dec")[t % 3])
            sense = int(rng.integers(0, 2))
            want = orc.revised_solve(A, np.abs(b), c, None, sense, max_iterations=120)
            got = api.revised_solve(A, np.abs(b), c, None, sense, max_iterations=120)
            ok = got["status"] == want["status"]
            if ok and want["status"] != -11:
                ok = (got["enter"].tolist() == want["enter"].tolist() and got["leave"].tolist() == want["leave"].tolist()
                      and same(got["Binv"], want["Binv"]) and same(got["xB"], want["xB"]))
            if not ok:
                bad += 1
                print("MISMATCH revised", m, n, want["status"], got["status"], flush=True)
        # 5. Branch & Bound: batches of small IPs (threaded host commit, cluster kernel for deep nodes)
        if rd % every == 0:
            m, n, bcnt = int(rng.integers(3, 9)), int(rng.integers(3, 10)), 48
            A = rng.integers(1, 12, size=(bcnt, m, n)).astype(float)
            b = rng.integers(3 * n, 12 * n, size=(bcnt, m)).astype(float)
            c = rng.integers(1, 15, size=(bcnt, n)).astype(float)
            got = api.bnb_simplex_batched(A, b, c)
            for k in range(bcnt):
                want = orc.bnb_simplex(A[k], b[k], c[k], node_cap=1 << 16)
                ok = (bool(got["found"][k]) == want["found"] and got["n_nodes"][k] == want["n_nodes"]
                      and got["lp_pivots"][k] == want["total_pivots"])
                if ok and want["found"]:
                    ok = same([got["best_z"][k]], [want["best_z"]]) and same(got["best_x"][k], want["best_x"])
                if not ok:
                    bad += 1
                    print("MISMATCH bnb", m, n, k, got["n_nodes"][k], want["n_nodes"], flush=True)
            # one deeper instance so that nodes leave one SM's shared memory
            A1 = rng.integers(1, 20, size=(40, 80)).astype(float)
            b1 = rng.integers(5 * 80, 15 * 80, size=40).astype(float)
            c1 = rng.integers(1, 30, size=80).astype(float)
            want = orc.bnb_simplex(A1, b1, c1, node_cap=1 << 16)
            g1 = api.bnb_simplex_batched(A1[None], b1[None], c1[None])
            if not (bool(g1["found"][0]) == want["found"] and g1["n_nodes"][0] == want["n_nodes"]
                    and g1["lp_pivots"][0] == want["total_pivots"]
                    and (not want["found"] or same(g1["best_x"][0], want["best_x"]))):
                bad += 1
                print("MISMATCH bnb deep", g1["n_nodes"][0], want["n_nodes"], flush=True)
        # 6. knapsack: integer (order-free sums) and fractional (ordered sums) data, several speculation settings
        if rd % every == (1 if every > 1 else 0):
            for t in range(6):
                nk = int(rng.integers(5, 120))
                w = rng.integers(1, 60, size=nk).astype(float)
                pf = rng.integers(1, 90, size=nk).astype(float)
                if t % 2:
                    w = np.round(w / 7.0, 3)
                    pf = np.round(pf / 3.0, 2)
                cap = float(np.floor(w.sum() * rng.choice([0.3, 0.5, 0.7])))
                want = orc.knapsack(pf, w, cap)
                for spec in ((0, 0), (1, 1), (4, 2), (32, 5)):
                    got = api.bnb_knapsack(pf, w, cap, spec_nodes=spec[0], spec_depth=spec[1])
                    ok = (got["found"] == want["found"] and got["n_evals"] == want["n_evals"]
                          and got["n_pops"] == want["n_pops"]
                          and (not want["found"] or (same([got["best"]], [want["best"]])
                                                     and list(got["best_x"]) == list(want["best_x"]))))
                    if not ok:
                        bad += 1
                        print("MISMATCH knapsack", nk, spec, got["n_evals"], want["n_evals"], flush=True)
        print(f"round {rd} done, mismatches so far {bad}, {time.time() - t0:.0f} s", flush=True)
    return bad


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bad = sweep(seed, rounds)
    print("FUZZ RESULT:", "OK" if bad == 0 else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
