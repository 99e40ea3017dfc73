"""CPU checks of the arithmetic identities the round-2 integrator relies on (csrc/rwrt_device.cuh).  They restate the
device formulas in exact rational arithmetic (fractions -> one rounding to double = what a fused multiply-add
delivers) and compare with NumPy, i.e. with what the reference computes.  The CUDA code itself is compared with the
oracle on the GPU (tests/test_gpu_parity.py::test_rhs_on_adversarial_longitudes and the trajectory tests); these tests
pin the REASONING, so that a change of a constant or of a rounding step shows up on the build container already."""
import math
from fractions import Fraction as F

import numpy as np

TWOPI = 2.0 * math.pi


def fma(a, b, c):
    """round(a * b + c) with ONE rounding (finite arguments)."""
    return float(F(a) * F(b) + F(c))


def lon_direct(lon):
    """locate() / locate_cell(): the remainder on the signed value; None = 'take the general routine'."""
    q = math.floor(lon * (1.0 / TWOPI))
    r = fma(-q, TWOPI, lon)
    return r if (r >= 0.0 and r < TWOPI) else None


def test_signed_longitude_wrap_equals_numpy_remainder():
    """bs.py:519 + interpolation.py:80: whenever fma(-floor(lon / 2pi), 2pi, lon) lies in [0, 2pi) it IS
    np.remainder(lon, 2pi), bit for bit, for either sign -- and the general routine is needed only rarely."""
    rng = np.random.default_rng(7)
    mult = np.arange(-50, 51, dtype=np.float64) * TWOPI
    lons = np.concatenate([rng.uniform(-100.0, 100.0, 20000), rng.uniform(-1e7, 1e7, 5000), rng.uniform(-1e15, 1e15, 2000),
                           mult, np.nextafter(mult, np.inf), np.nextafter(mult, -np.inf),
                           [0.0, -0.0, 5e-324, -5e-324, 1e-300, -1e-300, -1e-20, np.nextafter(TWOPI, 0.0), -np.nextafter(TWOPI, 0.0)]])
    fallback = 0
    for lon in lons:
        got = lon_direct(float(lon))
        want = float(np.remainder(lon, TWOPI))
        if got is None:
            fallback += 1
            continue
        assert got == want and math.copysign(1.0, got) == 1.0, (lon, got, want)
        assert not want >= TWOPI                     # (the second % of interpolation.py:80 changes nothing on this path)
    assert fallback < 0.01 * len(lons), fallback    # exact multiples, off-by-one quotients, results that round to 2 pi
    # what must NOT be decided on this path: a remainder that rounds to 2 pi itself
    assert lon_direct(-1e-20) is None and float(np.remainder(-1e-20, TWOPI)) == TWOPI


def test_division_by_sqrt5_in_three_instructions_is_correctly_rounded():
    """rkf45.py:29-31 (the rms norm): q = x y, r = x - q c (exact), q + r y with c = RN(sqrt 5), y = RN(1 / c) equals
    x / c -- Markstein's final division step, used by the fast build instead of the 18-instruction division."""
    c, y = 2.23606797749978969641, 0.44721359549995793928
    assert c == math.sqrt(5.0) and y == 1.0 / c
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(0.0, 4.0, 20000) * 10.0 ** rng.integers(-12, 6, 20000), [0.0, 1.0, c, 5.0, 1e-300, 1e300]])
    for x in xs:
        x = float(x)
        q = x * y
        r = fma(-q, c, x)
        assert fma(r, y, q) == x / c, x


def test_integer_maximum_of_bit_patterns_is_numpy_maximum_of_magnitudes():
    """rkf45.py:443: np.maximum(|a|, |b|) propagates NaN; so does the integer maximum of the bit patterns with the sign
    bit cleared (every NaN pattern lies above +inf), and for numbers it picks the larger magnitude."""
    rng = np.random.default_rng(13)
    a = rng.standard_normal(20000) * 10.0 ** rng.integers(-300, 300, 20000)
    b = rng.standard_normal(20000) * 10.0 ** rng.integers(-300, 300, 20000)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 5e-324, -5e-324, 1.0, -1.0])
    a = np.concatenate([a, np.repeat(special, len(special))])
    b = np.concatenate([b, np.tile(special, len(special))])
    mask = np.int64(0x7FFFFFFFFFFFFFFF)
    got = np.maximum(a.view(np.int64) & mask, b.view(np.int64) & mask).view(np.float64)
    with np.errstate(invalid="ignore"):
        want = np.maximum(np.abs(a), np.abs(b))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.int64), want[ok].view(np.int64))


def test_clamped_far_index_is_one_unsigned_compare():
    """interpolation.py:112-119: x1 = clamp(x0u + 1) equals x0 = clamp(x0u) exactly when the unclamped index lies
    outside [0, W - 2] -- locate_cell()'s (unsigned)x0u > (unsigned)(W - 2)."""
    for W in (2, 3, 73, 145, 1441):
        x0u = np.concatenate([np.arange(-5, W + 5), [-(2 ** 31), 2 ** 31 - 2]]).astype(np.int64)
        x0 = np.clip(x0u, 0, W - 1)
        x1 = np.clip(x0u + 1, 0, W - 1)
        flag = (x0u.astype(np.int32).view(np.uint32) > np.uint32(W - 2))
        assert np.array_equal(flag, x1 == x0), W
