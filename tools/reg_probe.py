"""C2 register kernel: ms per 4096-LP batch for both builds, plus barrier arrival stamps (LPX_REG_STAMPS=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_probe  # noqa: F401  (initialises the library)
from gpu_probe import time_batched
from linear_programming_solver_lpr381_b200 import _ffi as F
for rv in (1, 2):
    print("reg_variant", rv, time_batched(F.KERNEL_CTA_REG, reg_variant=rv), flush=True)
