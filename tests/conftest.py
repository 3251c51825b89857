import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    import orc_ffi
    orc_ffi.lib()
    return orc_ffi


@pytest.fixture(scope="session")
def lpx():
    """The CUDA engine through its C ABI; initialises device 0.  GPU tests only."""
    from linear_programming_solver_lpr381_b200 import _ffi, api
    _ffi.check(_ffi.lib().lpx_init(0))
    return api


def unhex(v):
    if isinstance(v, list):
        return [unhex(x) for x in v]
    return float.fromhex(v)


def case_arrays(case):
    A = np.array(unhex(case["A"]), dtype=np.float64)
    b = np.array(unhex(case["b"]), dtype=np.float64)
    c = np.array(unhex(case["c"]), dtype=np.float64)
    rel = np.array(case["rel"], dtype=np.int32)
    return A, b, c, rel


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bits_equal(got, want, what=""):
    g, w = bits(np.asarray(got, dtype=np.float64)), bits(np.asarray(want, dtype=np.float64))
    assert g.shape == w.shape, f"{what}: shape {g.shape} != {w.shape}"
    if not np.array_equal(g, w):
        idx = np.argwhere(g != w)[0]
        raise AssertionError(f"{what}: first differing element at {tuple(idx)}: "
                             f"{np.asarray(got).reshape(g.shape)[tuple(idx)]!r} vs "
                             f"{np.asarray(want).reshape(w.shape)[tuple(idx)]!r} ({len(np.argwhere(g != w))} differ)")
