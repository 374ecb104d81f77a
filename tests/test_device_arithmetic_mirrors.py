"""CPU mirrors of three arithmetic shortcuts of the CUDA kernels, instruction for instruction, with every FMA and rounding
emulated exactly (rational arithmetic, one correctly rounded conversion per instruction).  They pin the CLAIMS made about
the shortcuts where the kernels state them; the kernels themselves are checked on the GPU (tests/test_gpu_parity.py).

* find_bin_uniform (tdvmc_b200/csrc/common.cuh): knot interval of a distance on uniform knots from r / h alone.
* sqrt_fast (tdvmc_b200/csrc/sweep_math.cuh): hardware seed + one Newton step, relative error <= 1.5 delta^2 (+ rounding),
  delta = the seed's error INCLUDING its empty low word (20 mantissa bits).
* Recip::divide (tdvmc_b200/csrc/tables.cu): the reference's IEEE divisions from one reciprocal (Markstein).
"""
import random
import struct
from fractions import Fraction

import numpy as np
import pytest

from tdvmc_b200 import splines

MAGIC = 6755399441055744.0  # 1.5 * 2^52


def rn(x):
    """Fraction -> nearest double (ties to even): Fraction.__float__ is correctly rounded."""
    return float(x)


def fma(a, b, c):
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def from_bits(u):
    return struct.unpack("<d", struct.pack("<Q", u))[0]


def empty_low_word(x):
    """MUFU.RCP64H / MUFU.RSQ64H write the high word only."""
    return from_bits(bits(x) & 0xFFFFFFFF00000000)


# ---- find_bin_uniform ------------------------------------------------------------------------------------------------
def bin_exact(knots, first_bin, r):
    """std::lower_bound(nodes, r) - 1  <=>  knots[bin] < r <= knots[bin + 1]  (BosonsBulk.cpp:197-198)."""
    return int(np.searchsorted(knots, r, side="left")) - 1


def bin_uniform(knots, first_bin, inv_h, guard, r):
    """find_bin_uniform: x = r * inv_h; nearest integer through the magic constant; the knots only inside the guard band."""
    x = r * inv_h
    y = x + MAGIC
    j = (bits(y) & 0xFFFFFFFF)                       # __double2loint(y)
    j = j - (1 << 32) if j >= (1 << 31) else j
    d = x - (y - MAGIC)
    if abs(d) > guard:
        return first_bin + j - (1 if d < 0.0 else 0), True
    return bin_exact(knots, first_bin, r), False


def check_uniform_index(knots, seed):
    K = len(knots) - 4
    fb = 3
    assert knots[fb] == 0.0
    nbins = K - fb
    h0 = (knots[K] - knots[fb]) / nbins                                           # capi.cu build_static_tables
    dev = max(abs(knots[j] - (j - fb) * h0) / h0 for j in range(fb, K + 1))
    assert dev < 1e-7
    guard = 2.0 * dev + 1e-10
    inv_h = 1.0 / h0
    rng = np.random.default_rng(seed)
    rs = list(rng.uniform(1e-6, knots[K], 20000))
    for j in range(fb + 1, K + 1):
        t = float(knots[j])
        rs += [t, np.nextafter(t, 0.0), np.nextafter(t, 1e9)]
        rs += [t + s * e * h0 for e in (1e-12, 1e-11, 1e-10, 3e-10, 1e-9) for s in (-1.0, 1.0)]
    fast = 0
    for r in rs:
        r = float(r)
        if not (0.0 < r <= knots[K]):
            continue
        got, was_fast = bin_uniform(knots, fb, inv_h, guard, r)
        fast += was_fast
        assert got == bin_exact(knots, fb, r), (r, got)
    assert fast > 19990


@pytest.mark.parametrize("n_param,half_length", [(201, 3.5), (50, 2.0), (201, 10.0), (100, 0.731)])
def test_uniform_interval_index_mirror(n_param, half_length):
    """The guard band of tdvmc_gpu_create (twice the largest deviation of a stored knot from the exact grid, in units of
    the spacing, + 1e-10) makes the knot-free index exact: random distances, distances ON knots, one ulp and 1e-12 ... 1e-9 h
    beside them.  The fast path must also be what nearly every distance takes."""
    knots = np.asarray(splines.uniform_knots(n_param, half_length), np.float64)   # (i * L / 2) / (P - 1), BosonsBulk.cpp:61-67
    check_uniform_index(knots, n_param)


@pytest.mark.parametrize("name", ["bosonsbulk_n343_equil", "nubosonsbulkpb_n1728_equil"])
def test_uniform_interval_index_mirror_on_reference_knots(golden, name):
    """The same on knot vectors the reference itself produced: BosonsBulk's own grid at the headline size and the NURBS_GRID
    of config/NUBosonsBulkPB3D.config (decimal literals 0.03 i: uniform to a few ulp, which the guard band absorbs)."""
    check_uniform_index(np.asarray(golden(name)["knots"], np.float64), 7)


# ---- sqrt_fast -------------------------------------------------------------------------------------------------------
def sqrt_fast(x, delta):
    """rsqrt seed with relative error delta and an empty low word, t = x y, t + (x - t^2) y / 2 (two FMAs)."""
    y = empty_low_word(rn(Fraction(1) / Fraction(np.sqrt(x)) * (1 + Fraction(delta))))
    t = x * y
    hy = from_bits(bits(y) - (0x00100000 << 32))     # exponent decrement on the high word
    assert hy == 0.5 * y
    return fma(fma(-t, t, x), hy, t)


def test_sampler_square_root_error_bound():
    """|sqrt_fast(x) / sqrt(x) - 1| <= 1.5 delta^2 + a few ulp, delta = error of the seed as the kernel sees it: the seed's own
    error (here up to 2^-21) plus the truncation to the 20 mantissa bits of its high word (< 2^-20) - i.e. <= 1.4e-12 for a
    seed that is exact before truncation, over the squared distances a box of L = 7 ... 20 produces.  (On the device the
    exponent change of a proposal comes out within 3e-13 of exact arithmetic,
    test_sampler_exponent_change_against_exact_arithmetic.)"""
    random.seed(3)
    worst = 0.0
    for _ in range(4000):
        x = random.uniform(1e-4, 300.0)
        delta = random.uniform(-1.0, 1.0) * 2.0 ** -21
        r = sqrt_fast(x, delta)
        exact = Fraction(x)
        # relative error of r against the exact root: (r^2 - x) / (2 x) to first order, evaluated exactly
        rel = abs(float((Fraction(r) * Fraction(r) - exact) / (2 * exact)))
        bound = 1.5 * (abs(delta) + 2.0 ** -20) ** 2 * 1.001 + 4 * 2.0 ** -53   # 2^-20: the seed's emptied low word
        worst = max(worst, rel / bound)
        assert rel <= bound, (x, delta, rel, bound)
    assert worst > 0.2  # the bound is not vacuous
    assert 1.5 * (2.0 ** -21 + 2.0 ** -20) ** 2 < 3.1e-12


# ---- Recip::divide ---------------------------------------------------------------------------------------------------
def recip(r, seed_err):
    y0 = empty_low_word((1.0 / r) * (1.0 + seed_err))
    e = fma(-r, y0, 1.0)
    y0 = fma(y0, e, y0)
    e = fma(-r, y0, 1.0)
    y0 = fma(y0, e, y0)
    e = fma(-r, y0, 1.0)
    return fma(y0, e, y0)


def divide(x, r, y):
    q0 = x * y
    return fma(fma(-r, q0, x), y, q0)


def all_ones_significand(r):
    return bits(r) & 0x000FFFFFFFFFFFFF == 0x000FFFFFFFFFFFFF


def test_markstein_division_is_the_ieee_quotient():
    """x / r from ONE reciprocal equals the correctly rounded (IEEE) quotient the reference computes, for every divisor whose
    significand is not all ones; for those (one double in 2^52 - the Newton step cannot reach RN(1 / r) there) the quotient may
    be one ulp off, which this test documents rather than hides."""
    random.seed(1)
    odd_seen = odd_off = 0
    for _ in range(6000):
        r = random.uniform(0.01, 20.0)
        special = random.random() < 0.25
        if special:  # awkward significands
            m = random.choice([0x000FFFFFFFFFFFFF, 0x000FFFFFFFFFFFFE, 0x1, 0x0008000000000001, 0x0007FFFFFFFFFFFF, 0x0])
            r = from_bits((bits(r) & 0xFFF0000000000000) | m)
        x = random.choice([random.uniform(-10.0, 10.0), 2.0, 1.0, 0.0, random.uniform(-1e-3, 1e-3)])
        y = recip(r, random.uniform(-1.0, 1.0) * 2.0 ** -20)
        q = divide(x, r, y)
        want = rn(Fraction(x) / Fraction(r))
        if all_ones_significand(r):
            odd_seen += 1
            odd_off += q != want
            assert abs(q - want) <= abs(np.spacing(want))
        else:
            assert y == rn(Fraction(1) / Fraction(r)), r.hex()
            assert q == want, (x.hex(), r.hex())
    assert odd_seen > 100 and odd_off > 0
