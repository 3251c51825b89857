"""C2 register kernel: ms per 4096-LP batch for every build, each checked against the default build's tableaux."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import gpu_probe  # noqa: F401  (initialises the library)
from gpu_probe import time_batched
from linear_programming_solver_lpr381_b200 import _ffi as F, api, workloads
A, b, c = workloads.batch_c2(count=256, seed=3)
ref = api.primal_solve_batched(A, b, c, kernel=F.KERNEL_CTA_REG, reg_variant=2)
for rv in (1, 2, 3, 4):
    got = api.primal_solve_batched(A, b, c, kernel=F.KERNEL_CTA_REG, reg_variant=rv)
    same = got["tableau"].tobytes() == ref["tableau"].tobytes() and np.array_equal(got["n_pivots"], ref["n_pivots"])
    print("reg_variant", rv, "bit-identical" if same else "MISMATCH", time_batched(F.KERNEL_CTA_REG, reg_variant=rv), flush=True)
