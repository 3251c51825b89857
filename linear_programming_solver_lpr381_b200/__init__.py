"""B200-native LP/IP solve engine behind the solver entry points of
Jellyman750/Linear_Programming_Solver_LPR381 (Primal Simplex, Branch & Bound Simplex,
Branch & Bound Knapsack).

The product is `liblpx.so` (csrc/, C ABI in include/lpx.h) plus the C++ host layer that mirrors
the reference's controllers (host/).  This Python package is only the ctypes view used by the
tests and bench.py; it holds no solver logic and no CPU fallback.
"""
from . import _ffi, api, workloads  # noqa: F401

__all__ = ["_ffi", "api", "workloads"]
