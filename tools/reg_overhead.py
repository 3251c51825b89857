"""C2 register kernel: how much of a 4096-LP batch is the per-LP fixed cost (build the tableau from A, b, c; write
the final tableau, x, z, basis) and how much the pivots?  Batches cut off after 0, 1, 8 pivots against the full solve."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_probe  # noqa: F401  (initialises the library)
from gpu_probe import time_batched
from linear_programming_solver_lpr381_b200 import _ffi as F
for mi in (0, 1, 8, 16, 10000):
    r = time_batched(F.KERNEL_CTA_REG, max_iterations=mi)
    print(f"max_iterations {mi:6d}: {r['ms']:.3f} ms per batch, {r['pivots']:.0f} pivots", flush=True)
