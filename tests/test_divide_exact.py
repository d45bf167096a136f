"""The numerical claim behind ``qd_div_u`` / ``qd_div_exact`` (csrc/qd_ops.cuh): with y = RN(1/b), the sequence
q0 = RN(x y); r = fma(-b, q0, x); q1 = fma(r, y, q0); r = fma(-b, q1, x); q = fma(r, y, q1) returns RN(x / b), i.e. the
bits of the IEEE division the reference performs, for every x whose first estimate lies in the guarded range
2^-500 <= |q0| < 2^500 and every divisor with 1e-100 < |b| < 1e100 (outside, the kernels take the plain division).
Checked here with an exact software FMA (rational arithmetic, correctly rounded conversion) on random and awkward
operands; the GPU parity tests exercise the same code on the device."""
from fractions import Fraction

import numpy as np


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))        # int/int -> float conversion is correctly rounded


def div_u(x, b):
    y = 1.0 / b
    q = x * y
    r = fma(-b, q, x)
    q = fma(r, y, q)
    r = fma(-b, q, x)
    return fma(r, y, q)


def _guarded(x, b):
    q0 = abs(x * (1.0 / b))
    return 2.0 ** -500 <= q0 < 2.0 ** 500


def test_two_fma_corrections_give_the_correctly_rounded_quotient():
    rng = np.random.default_rng(17)
    divisors = [3.0, 7.0, 1.0 / 3.0, 12.0, 2e-5, 1000.0, 86400.0, 5.670374e-8, 1004.0, 300.0, 37.0, 1.2 * 800.0, 917.0 * 3.34e5,
                2.0e7, 1e6, 0.5, 6.371e6 * 0.0021816615649929116, -9.81, 1e-99, 9.9e99]
    divisors += list(np.ldexp(rng.uniform(1.0, 2.0, 200), rng.integers(-300, 300, 200)))
    bad = 0
    n = 0
    for b in divisors:
        xs = np.ldexp(rng.uniform(1.0, 2.0, 100), rng.integers(-400, 400, 100)) * rng.choice([-1.0, 1.0], 100)
        # numerators that make the quotient land next to a rounding boundary: q*b for q just around representable values
        qs = np.ldexp(rng.uniform(1.0, 2.0, 50), rng.integers(-50, 50, 50))
        xs = np.concatenate([xs, qs * b, np.nextafter(qs * b, np.inf), np.nextafter(qs * b, -np.inf)])
        for x in xs:
            x = float(x)
            if not (1e-100 < abs(b) < 1e100) or not _guarded(x, b):
                continue
            n += 1
            if div_u(x, float(b)) != x / float(b):
                bad += 1
    assert n > 30000 and bad == 0, (n, bad)


def test_one_correction_is_not_enough():
    """Why there are two: a single residual step is faithful but not always correctly rounded."""
    rng = np.random.default_rng(3)
    miss = 0
    for _ in range(40000):
        b = float(rng.uniform(1.0, 2.0)); x = float(rng.uniform(1.0, 2.0))
        y = 1.0 / b
        q0 = x * y
        q1 = fma(fma(-b, q0, x), y, q0)
        q0_wrong = q0 != x / b
        miss += q0_wrong
        assert abs(q1 - x / b) <= np.spacing(x / b)              # faithful
    assert miss > 0                                              # the bare reciprocal multiply does misround
