"""Replays, in exact rational arithmetic, the two arithmetic shortcuts of atsc_b200/csrc/common.cuh:
div_1e5_int53 (ONE FMA refinement step of n * RN(1e-5) is the IEEE quotient n / 1e5 for every integer
|n| < 2^53) and round_half_away (trunc(y + copysign(pred(0.5), y)) == f64::round).
    python tools/div_check.py [samples_per_width]"""
import math
import random
import sys
from fractions import Fraction as F

B, Y = 100000.0, 1e-5
PRED_HALF = 0.49999999999999994


def fma(a, b, c):
    """RN(a * b + c): float(Fraction) rounds to nearest even, exactly like the hardware FMA."""
    return float(F(a) * F(b) + F(c))


def div_1e5_int53(n):
    q = n * Y
    return fma(fma(-B, q, n), Y, q)


def round_half_away(y):
    return float(math.trunc(y + math.copysign(PRED_HALF, y)))


def round_ref(y):
    r = float(math.trunc(y))
    if abs(y - r) >= 0.5:
        r += math.copysign(1.0, y)
    return r


def check_division(per_width, rng):
    bad = []
    for bits in range(1, 54):
        for _ in range(per_width):
            n = float(rng.getrandbits(bits) | (1 << (bits - 1)))
            for v in (n, -n):
                if div_1e5_int53(v) != v / B:
                    bad.append(v)
    for _ in range(per_width * 10):  # around the values whose quotient is a short dyadic
        n = float(rng.getrandbits(rng.randint(1, 40)) * 3125 + rng.randint(-2, 2))
        if div_1e5_int53(n) != n / B:
            bad.append(n)
    return bad


def check_rounding(count, rng):
    bad = []
    for _ in range(count):
        y = math.ldexp(rng.random() + 1, rng.randint(-5, 54)) * rng.choice([-1, 1])
        if rng.random() < 0.5:
            y = float(round(y)) + rng.choice([0.5, -0.5, PRED_HALF, -PRED_HALF, math.nextafter(0.5, 1), 0.25])
        if round_half_away(y) != round_ref(y):
            bad.append(y)
    for y in (0.0, -0.0, 0.5, -0.5, PRED_HALF, -PRED_HALF, 2.0 ** 52, 2.0 ** 52 + 1, 2.0 ** 51 + 0.5, -(2.0 ** 51 + 0.5), 1e300):
        r = round_half_away(y)
        if r != round_ref(y) or math.copysign(1.0, r) != math.copysign(1.0, round_ref(y)):
            bad.append(y)
    return bad


if __name__ == "__main__":
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rng = random.Random(1)
    bd, br = check_division(per, rng), check_rounding(per * 100, rng)
    print("division mismatches:", len(bd), bd[:5], "| rounding mismatches:", len(br), br[:5])
    sys.exit(1 if bd or br else 0)
