"""k_poly1s rounds the spline values of a TAME frame (every sample of one sign, 1e-200 <= |sample| <= 1e10) with a
frame-uniform addend: trunc(v * 1e5 + copysign(pred(0.5), vmin)) instead of f64::round's copysign(pred(0.5), v)
(poly.cuh: p1_trip).  The two differ only for a spline value on the other side of zero, and the claim is that the clamp
to [vmin, vmax] of round_and_limit_f64 (utils/mod.rs:66-74) hides the difference.  This replays both forms in IEEE
double arithmetic (numpy) on random and adversarial values:  python tools/half_check.py"""
import numpy as np

H = np.float64(0.49999999999999994)


def limit(out, vmin, vmax):
    out = np.where(out < vmin, vmin, out)   # same comparison order as the device code; NaN passes through
    return np.where(out > vmax, vmax, out)


def exact(v, vmin, vmax):
    return limit(np.trunc(v * 1e5 + np.copysign(H, v * 1e5)) / 1e5, vmin, vmax)


def uniform(v, vmin, vmax):
    return limit(np.trunc(v * 1e5 + np.copysign(H, vmin)) / 1e5, vmin, vmax)


def check(n=2_000_000, seed=1):
    rng = np.random.default_rng(seed)
    bad = 0
    for sign in (1.0, -1.0):
        for lo_exp, hi_exp in ((-200, -190), (-12, -4), (-6, 0), (-3, 3), (0, 10)):
            a = 10.0 ** rng.uniform(lo_exp, hi_exp, 2)
            vmin, vmax = (min(a), max(a)) if sign > 0 else (-max(a), -min(a))
            mag = 10.0 ** rng.uniform(-320, 11, n)
            mag[: n // 4] = rng.uniform(0, 3e-5, n // 4)           # around the first rounding steps
            mag[n // 4: n // 2] = (rng.integers(0, 50, n // 4) + 0.5) * 1e-5 * (1 + rng.integers(-2, 3, n // 4) * 2.0 ** -52)
            v = np.concatenate([mag, -mag, [0.0, -0.0, np.nan, np.inf, -np.inf]])
            with np.errstate(invalid="ignore", over="ignore"):
                e, u = exact(v, vmin, vmax), uniform(v, vmin, vmax)
            same = (e == u) | (np.isnan(e) & np.isnan(u))
            bad += int((~same).sum())
    return bad


if __name__ == "__main__":
    b = check()
    print("mismatches:", b)
    raise SystemExit(1 if b else 0)
