"""Proof, by enumeration, that luma::div_const (luma_b200/csrc/lattice.cuh) is correctly rounded.

The reference divides by the run-time constants b1 = cs^2 and b2 = 2 cs^4 with IEEE `/`
(src/GridObj_ops_lbm_optimised.cpp:704).  The CUDA path evaluates

        q = RN(a*y);  r = fma(-q, b, a);  q' = fma(r, y, q)          with y = RN(1/b)

Claim: q' == RN(a/b) for every double a whose quotient is a normal number, for b in {b1, b2}.

Argument.  r = a - q*b is exact (q is within 1.5 ulp of a/b).  With y = (1+eps)/b, |eps| <= 2^-53,
    q + r*y = a/b + (a/b - q)*eps,   |(a/b - q)*eps| <= 1.5 * 2^-53 ulp(a/b).
So q' can differ from RN(a/b) only when a/b lies within 1.5*2^-53 ulp of a rounding boundary, i.e. of
a midpoint m = Mo * 2^(e-53) (Mo odd, 54 bits).  Writing a = A*2^-52, b = Bb*2^sb (integers), that is
    |A*2^k - Bb*Mo| = |I| <= 1.5*Bb*2^-52 < 3,   k = 1 - sb - e,
a linear congruence with a handful of solutions (A, Mo) per binade e of the quotient.  The test
enumerates every solution and replays the three operations in exact rational arithmetic.  By
scaling, a in [1,2) covers all normal a (powers of two do not change significands); the sign is
symmetric.  Random replay guards the argument itself.
"""
import math
import random
from fractions import Fraction

import pytest


def RN(x: Fraction) -> float:
    """round-to-nearest-even to double (Python's int/int true division is correctly rounded)."""
    return x.numerator / x.denominator


def fma(a: float, b: float, c: float) -> float:
    return RN(Fraction(a) * Fraction(b) + Fraction(c))


def div_const(a: float, b: float, y: float) -> float:
    q = a * y
    r = fma(-q, b, a)
    return fma(r, y, q)


def constants():
    cs = 1.0 / math.sqrt(3.0)          # src/stdafx.cpp:153
    cs2 = cs * cs
    return {"cs2": cs2, "2cs4": (2.0 * cs2) * cs2}


def decompose(b: float):
    m, e = math.frexp(b)               # b = m * 2^e, m in [0.5,1)
    Bb = int(m * (1 << 53))
    return Bb, e - 53                  # b = Bb * 2^sb


def candidates(b: float):
    """all a = A*2^-52 in [1,2) whose quotient a/b is within 3 units (of 2^(sb+e-53)) of a midpoint"""
    Bb, sb = decompose(b)
    out = set()
    lo, hi = Fraction(1) / Fraction(b), Fraction(2) / Fraction(b)
    e_lo = math.floor(math.log2(float(lo)))
    for e in (e_lo - 1, e_lo, e_lo + 1, e_lo + 2):
        k = 1 - sb - e
        if k <= 0:
            continue
        mod = 1 << k
        z = (Bb & -Bb).bit_length() - 1            # 2-adic valuation of Bb
        for I in (-2, -1, 1, 2):
            # A*2^k - Bb*Mo = I  ->  Bb*Mo = -I (mod 2^k)
            if I % (1 << z):
                continue
            m2 = mod >> z
            base = ((-I) >> z) * pow(Bb >> z, -1, m2) % m2 if m2 > 1 else 0
            Mo = base
            while Mo < (1 << 54):
                if Mo >= (1 << 53) and Mo % 2 == 1:
                    num = Bb * Mo + I
                    if num % mod == 0:
                        A = num // mod
                        if (1 << 52) <= A < (1 << 53):
                            out.add(A)
                Mo += m2
    return sorted(out)


@pytest.mark.parametrize("name", ["cs2", "2cs4"])
def test_div_const_is_correctly_rounded_on_every_candidate(name):
    b = constants()[name]
    y = 1.0 / b
    cands = candidates(b)
    # the enumeration must not be vacuous: neighbours of candidates are also replayed
    tested = 0
    for A in cands:
        for dA in (-1, 0, 1):
            a = (A + dA) / float(1 << 52)
            if not (1.0 <= a < 2.0):
                continue
            assert div_const(a, b, y) == RN(Fraction(a) / Fraction(b)), (name, a.hex())
            assert div_const(-a, b, y) == -RN(Fraction(a) / Fraction(b))
            tested += 1
    assert tested >= 1 or len(cands) == 0
    # r must be exact for the argument to hold: check on the candidates too
    for A in cands:
        a = A / float(1 << 52)
        q = a * y
        assert Fraction(fma(-q, b, a)) == Fraction(a) - Fraction(q) * Fraction(b)


@pytest.mark.parametrize("name", ["cs2", "2cs4"])
def test_div_const_random_replay(name):
    b = constants()[name]
    y = 1.0 / b
    rng = random.Random(20261017)
    for _ in range(40000):
        a = math.ldexp(rng.uniform(1.0, 2.0), rng.randint(-80, 20)) * rng.choice((-1.0, 1.0))
        assert div_const(a, b, y) == a / b, (name, a.hex())
    for a in (0.0, 1.0, 1.5, 0.05, 3.0, b, y, 2.0 ** -60):
        assert div_const(a, b, y) == a / b


def test_constants_match_the_reference_values():
    c = constants()
    assert c["cs2"].hex() == (1.0 / math.sqrt(3.0) * (1.0 / math.sqrt(3.0))).hex()
    assert abs(c["cs2"] - 1.0 / 3.0) < 2e-16 and c["cs2"] != 1.0 / 3.0    # SQ(cs) is NOT exactly 1/3
