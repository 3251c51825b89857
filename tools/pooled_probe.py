"""Mode B timing probe (GPU box): nodes/s of the pooled tree for a few instances and batch sizes."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
F.check(F.lib().lpx_init(0))
for (m, n, seed, batch) in ((60, 120, 11, 64), (60, 120, 11, 512), (60, 120, 12, 256), (60, 120, 12, 1024), (60, 120, 12, 4096), (40, 80, 5, 1024)):
    A, b, c = workloads.ip_c4(m=m, n=n, seed=seed)
    api.bnb_pooled(A, b, c, batch=batch)
    t0 = time.perf_counter()
    r = api.bnb_pooled(A, b, c, batch=batch)
    dt = time.perf_counter() - t0
    print(f"{m}x{n} seed {seed} batch {batch}: {r['n_nodes']} nodes, {r['rounds']} rounds, {r['total_pivots']} pivots, max pivots/node {int(r['pivots'].max())}, "
          f"{dt*1e3:.1f} ms = {r['n_nodes']/dt/1e3:.1f} k nodes/s, {dt/r['rounds']*1e3:.3f} ms/round", flush=True)
