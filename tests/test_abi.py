"""CPU suite: the C-ABI library loads, exports every symbol include/lpx.h declares, and fails
loudly (no CPU fallback) when there is no device.  No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from linear_programming_solver_lpr381_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lpx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpx_[a-z0-9_]+)\s*\(", src)) - {"lpx_bnb_node_fn", "lpx_knap_pop_fn"})


def test_library_exports_every_declared_symbol():
    L = F.lib()
    names = header_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(F.EXPORTS) == names


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "lpx.h")).read()
    assert "torch" not in src.lower().replace("no torch types", "") and "at::" not in src and "std::" not in src


def test_options_defaults_match_reference_constants():
    o = F.make_options()
    assert o.max_iterations == 10000  # PrimalSimplex.MaxIterations
    assert F.status_message(F.S_ITER_LIMIT) == "Iteration limit exceeded."
    assert F.status_message(F.S_GE_ROW).startswith("Constraint contains '>=' sign.")
    assert F.status_message(F.S_NEG_RHS).startswith("Constraint has a negative RHS value.")


def test_tableau_dims():
    rows, cols = C.c_int(), C.c_int()
    rel = np.array([0, 2, 0], dtype=np.int32)
    assert F.lib().lpx_tableau_dims(3, 4, F.ptr(rel), C.byref(rows), C.byref(cols)) == 0
    assert (rows.value, cols.value) == (5, 9)
    assert F.lib().lpx_tableau_dims(0, 4, None, C.byref(rows), C.byref(cols)) == F.E_BAD_ARGS


def test_bad_arguments_are_rejected_before_any_device_work():
    from linear_programming_solver_lpr381_b200 import api
    with pytest.raises(F.LpxError) as e:
        api.primal_solve(np.ones((2, 2)), np.ones(2), np.ones(2), rel=[0, 7])
    assert e.value.code == F.E_BAD_ARGS


def test_fails_loudly_without_a_gpu():
    if F.lib().lpx_device_count() > 0:
        pytest.skip("a GPU is present")
    from linear_programming_solver_lpr381_b200 import api
    with pytest.raises(F.LpxError) as e:
        api.primal_solve(np.ones((2, 2)), np.ones(2), np.ones(2))
    assert e.value.code == F.E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(F.LpxError):
        api.bnb_knapsack([1.0, 2.0], [1.0, 1.0], 1.0)
    with pytest.raises(F.LpxError):
        api.Session(np.ones((2, 2)), np.ones(2), np.ones(2))
