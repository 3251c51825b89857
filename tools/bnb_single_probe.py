"""Latency of ONE Branch & Bound tree without a callback (60 x 120, BASELINE config 4's instance shape)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
F.check(F.lib().lpx_init(0))
for seed in (11, 12, 13):
    A, b, c = workloads.ip_c4(seed=seed)
    api.bnb_simplex_batched(A[None], b[None], c[None])
    t0 = time.perf_counter()
    r = api.bnb_simplex_batched(A[None], b[None], c[None])
    dt = time.perf_counter() - t0
    print(f"seed {seed}: {int(r['n_nodes'][0])} nodes, {int(r['lp_pivots'][0])} pivots in {dt*1e3:.1f} ms = {r['n_nodes'][0]/dt/1e3:.1f} k nodes/s", flush=True)
