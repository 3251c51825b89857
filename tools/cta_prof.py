"""Per-phase cycle profile of the shared-memory per-CTA kernel on B&B-node-like shapes (LPX_CTA_PROF=1)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
F.check(F.lib().lpx_init(0))
for (m, n) in ((60, 120), (90, 120), (115, 120)):
    r = np.random.Generator(np.random.PCG64(5))
    A = r.integers(1, 20, size=(m, n)).astype(np.float64)
    b = r.integers(5 * n, 15 * n, size=m).astype(np.float64)
    c = r.integers(1, 30, size=n).astype(np.float64)
    for kernel in (F.KERNEL_CTA_SMEM,):
        try:
            g = api.primal_solve(A, b, c, kernel=kernel)
            print(m, n, "pivots", g["n_pivots"], flush=True)
        except Exception as e:
            print(m, n, "error", e)
