"""Timing probe for the knapsack engine (GPU box): nodes/s for several batch sizes and data kinds."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
F.check(F.lib().lpx_init(0))

def run(count, kind, n=2000, reps=2):
    ps, ws, caps = zip(*[workloads.knapsack_c5(n=n, seed=13 + k, kind=kind) for k in range(count)])
    p, w, cap = np.stack(ps), np.stack(ws), np.array(caps)
    best = None
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        r = api.bnb_knapsack_batched(p, w, cap)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ev = int(r["n_evals"].sum())
    print(f"{kind:13s} count={count:4d}: {ev:9d} evals (max {int(r['n_evals'].max())}) in {best*1e3:8.2f} ms = {ev/best/1e6:7.3f} M nodes/s", flush=True)

for kind, counts in (("uncorrelated", (1, 16, 148, 592, 2368)), ("weak", (1, 16, 148, 592)), ("fractional", (1, 16, 148, 592, 1184))):
    for c in counts:
        run(c, kind)
