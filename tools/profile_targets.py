"""Small, short-running targets for ncu captures of one kernel family each (see tools/final_profile_r02.sh).

    python tools/profile_targets.py bnb | knap | knapfrac | pooled
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads  # noqa: E402

F.check(F.lib().lpx_init(0))
what = sys.argv[1]
if what == "bnb":  # reference-exact B&B simplex, 64 instances of 60 x 120: cta_condensed_kernel
    As, bs, cs = zip(*[workloads.ip_c4(seed=11 + k) for k in range(64)])
    r = api.bnb_simplex_batched(np.stack(As), np.stack(bs), np.stack(cs))
    print("bnb", int(r["n_nodes"].sum()), "nodes")
elif what in ("knap", "knapfrac"):  # device-resident knapsack search, 148 instances of 2000 items: knap_search_kernel
    kind = "uncorrelated" if what == "knap" else "fractional"
    ps, ws, caps = zip(*[workloads.knapsack_c5(seed=13 + k, kind=kind) for k in range(148)])
    r = api.bnb_knapsack_batched(np.stack(ps), np.stack(ws), np.array(caps))
    print(what, int(r["n_evals"].sum()), "evaluations")
elif what == "pooled":  # Mode B, warm-started nodes
    A, b, c = workloads.ip_c4(seed=12)
    r = api.bnb_pooled(A, b, c, batch=4096)
    print("pooled", r["n_nodes"], "nodes")
