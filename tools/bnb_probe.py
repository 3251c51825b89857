"""C4 batch timing probe (GPU box): nodes/s and the per-kernel-family split (LPX_BNB_TRACE=1)."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
F.check(F.lib().lpx_init(0))
count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
As, bs, cs = zip(*[workloads.ip_c4(seed=11 + k) for k in range(count)])
A, b, c = np.stack(As), np.stack(bs), np.stack(cs)
api.bnb_simplex_batched(A, b, c)
for _ in range(2):
    t0 = time.perf_counter()
    r = api.bnb_simplex_batched(A, b, c)
    dt = time.perf_counter() - t0
    print(f"{count} instances: {int(r['n_nodes'].sum())} nodes, {int(r['lp_pivots'].sum())} pivots in {dt*1e3:.1f} ms = {r['n_nodes'].sum()/dt/1e3:.1f} k nodes/s", flush=True)
