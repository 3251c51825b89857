"""A bounded, fixed-seed slice of tests/fuzz_parity.py inside the collected suite: every kernel family
(register-resident both builds, shared-memory, global-memory, cluster, streaming, dual, revised, batched
B&B, knapsack speculation settings) against the oracle on tie-provoking data, bit for bit."""
import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_slice(lpx, orc):
    import fuzz_parity
    assert fuzz_parity.sweep(seed=20261018, rounds=1, cnt=48, singles=18, revised=10, every=1, threads=8) == 0
