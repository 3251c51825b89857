"""CPU suite: host-layer parser / number formatting / dispatcher keys against the oracle and
against a third, decimal-module implementation of the .NET formatting rules."""
import math
from decimal import ROUND_HALF_EVEN, ROUND_HALF_UP, Decimal

import numpy as np
import pytest

import host_ffi as H

VALID = [
    "Max: 3x1 + 5x2\n1x1 + 0x2 <= 4\n0x1 + 2x2 <= 12\n3x1 + 2x2 <= 18\n",
    "min : -x1 - 2.5x2 + .5x3\r\n x1 + x2 + x3 = 10 \r\n\r\n x1 - x2 + 0x3 >= -2\n2x1+3x2-x3<=1,000\n",
    "MAX:x1\nx1<=1e3\n",
    "Max: 2x1 + 3x2\nx7 + x9 <= 4\n-x1 - 2x2 <= 5\n",
    "Max: 0x1 + -0x2\n1.x1 + 1x2 <= 2.50\n",
]
INVALID = [
    "Max: 3x1 + 5x2\n", "Maximise: 3x1\nx1 <= 4\n", "Max: 3y1\nx1 <= 4\n", "Max: 3x1\nx1 < 4\n", "Max: 3x1\nx1 <= four\n",
    "Max: 3x1 + 2e3x2\nx1 + x2 <= 4\n", "Max: 3x1\nx1 <=\n", "Max: .x1\nx1 <= 2\n", "\n\n", "Max: 3x\nx1 <= 2\n", "Max: 2x1 + 3x2\n-x1 - -x2 <= 5\n",
]


@pytest.mark.parametrize("text", VALID)
def test_parser_matches_oracle(orc, text):
    a, b = H.parse_text(text), orc.parse_text(text)
    assert a["sense"] == b["sense"]
    for k in ("A", "b", "c", "rel"):
        assert a[k].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize("text", INVALID)
def test_parser_errors_match_oracle(orc, text):
    with pytest.raises(ValueError) as e1:
        H.parse_text(text)
    with pytest.raises(ValueError) as e2:
        orc.parse_text(text)
    assert str(e1.value) == e2.value.args[0][1]


def test_parser_positional_coefficients():
    p = H.parse_text("Max: 2x1 + 3x2\nx7 + x9 <= 4\n-x1 - 2x2 <= 5\n")
    assert p["A"].tolist() == [[1.0, 1.0], [-1.0, -2.0]] and p["b"].tolist() == [4.0, 5.0]


def dec_custom(v, d=3):
    """ToString("0.###") by the book: 15 significant digits, then half-up at d decimals."""
    if v == 0:
        return "-0" if math.copysign(1, v) < 0 else "0"
    x = Decimal(abs(v))
    e = x.adjusted()
    x15 = x.quantize(Decimal(1).scaleb(e - 14), rounding=ROUND_HALF_EVEN)
    q = x15.quantize(Decimal(1).scaleb(-d), rounding=ROUND_HALF_UP)
    s = format(q, "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return ("-" if v < 0 else "") + s


def sample_values():
    rng = np.random.default_rng(2)
    vals = [0.0, -0.0, 0.5, -0.5, 0.0005, 0.00049999999, 0.0625, 2.9995, 9.9995, 999.9995, 1e-7, -1e-7, 1 / 3, -2 / 3, 36.0,
            1e11, 123456789012.3456, 1e14 + 0.5, 1e15, 1e16, 1.23e20, 4.35, 2.675, 1.0000000005, 0.1 + 0.2]
    vals += list(rng.normal(size=300) * 10.0 ** rng.integers(-6, 9, size=300))
    vals += list(np.round(rng.normal(size=200) * 100, 3)) + list(rng.integers(-50, 50, size=50) + 0.0005)
    return [float(v) for v in vals]


def test_custom_format_three_way(orc):
    for v in sample_values():
        want = dec_custom(v)
        assert H.fmt_custom(v) == want, v
        assert orc.fmt_custom(v) == want, v


def test_fixed_roundtrip_and_round_match_oracle(orc):
    for v in sample_values():
        assert H.fmt_fixed(v, 3) == orc.fmt_fixed(v, 3) == ("%.3f" % v)
        assert H.fmt_fixed(v, 6) == orc.fmt_fixed(v, 6)
        assert H.fmt_roundtrip(v) == orc.fmt_roundtrip(v), v
        assert H.math_round(v, 3) == orc.math_round(v, 3)
        assert H.math_round(v, 6) == orc.math_round(v, 6)
    assert H.fmt_roundtrip(1e15) == "1E+15" and H.fmt_roundtrip(1e-5) == "1E-05" and H.fmt_roundtrip(0.0001) == "0.0001"
    assert H.fmt_roundtrip(123456789012345.0) == "123456789012345" and H.fmt_roundtrip(2.5) == "2.5"
    assert H.fmt_custom(float("inf")) == "∞" and H.fmt_custom(-0.0004) == "-0"


def test_normalize_algorithm_key():
    assert H.normalize_key("  Primal   Simplex Algorithm ") == "primal simplex"
    assert H.normalize_key("Branch and Bound Simplex") == "branch and bound simplex"
    assert H.normalize_key("BNB") == "bnb"
    assert H.normalize_key("   ") == "!No algorithm selected."
