"""EXPERIMENT: the register kernel on 64 x 64 LPs (n + m = 128 columns, what a condensed 64 x 128 tableau holds):
the six-slot default build against four-slot builds at two and three CTAs per SM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_probe  # noqa: F401
from gpu_probe import time_batched
from linear_programming_solver_lpr381_b200 import _ffi as F
for rv in (0, 1, 3, 10, 11, 12, 13):
    try:
        r = time_batched(F.KERNEL_CTA_REG, reg_variant=rv, n=64)
        print(f"reg_variant {rv:2d}: {r['ms']:.3f} ms per batch, {r['pivots']:.0f} pivots, {r['mpivots_s']:.1f} M pivots/s", flush=True)
    except Exception as e:
        print("reg_variant", rv, "error", e, flush=True)
