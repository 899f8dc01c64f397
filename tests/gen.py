"""Seeded synthetic series of the shapes SURVEY.md section 8(d) names (numpy, host side).
Used by the parity tests and by bench.py so GPU and CPU arms see identical inputs."""
import numpy as np


def periodic(n, seed, sigma=0.5):
    """C2: 100 + 20 sin(2 pi i/1440) + 5 sin(2 pi i/97) + sigma N(0,1); strictly positive."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64)
    return 100.0 + 20.0 * np.sin(2 * np.pi * i / 1440.0) + 5.0 * np.sin(2 * np.pi * i / 97.0) + \
        sigma * rng.standard_normal(n)


def gauge_walk(n, seed):
    """C3(a): integer gauge random walk, steps +-{0..3}*4096 around 5e7 (heap-like, I32)."""
    rng = np.random.default_rng(seed)
    steps = rng.integers(-3, 4, size=n) * 4096
    hold = rng.random(n) < 0.7  # gauges hold their value most of the time
    steps[hold] = 0
    return (5e7 + np.cumsum(steps)).astype(np.float64)


def utilisation(n, seed):
    """C3(b): clip(50 + 30 sin(2 pi i/4320) + AR(1) noise sigma=2, 0.01, 100), 2 decimals."""
    rng = np.random.default_rng(seed)
    i = np.arange(n, dtype=np.float64)
    e = rng.standard_normal(n) * 2.0 * np.sqrt(1 - 0.9 ** 2)
    ar = np.empty(n)
    acc = 0.0
    # AR(1) with phi = 0.9 (vectorised through lfilter-free recursion in chunks)
    from scipy.signal import lfilter
    ar = lfilter([1.0], [1.0, -0.9], e)
    x = np.clip(50.0 + 30.0 * np.sin(2 * np.pi * i / 4320.0) + ar, 0.01, 100.0)
    return np.round(x, 2)


def sawtooth(n, seed):
    """C3(c): integer sawtooth counter 0..255 with period 300 (U8)."""
    i = np.arange(n) + seed
    return np.minimum((i % 300), 255).astype(np.float64)


def constant(n, seed):
    return np.full(n, float(seed % 1000), dtype=np.float64)


def noisy(n, seed):
    """positive white noise with large relative spread: nothing reaches 5 % cheaply"""
    rng = np.random.default_rng(seed)
    return np.round(rng.uniform(1.0, 100.0, n), 3)


def steps(n, seed):
    """few long runs of a handful of integer levels: index-RLE territory"""
    rng = np.random.default_rng(seed)
    nseg = max(2, n // 500)
    cuts = np.sort(rng.choice(np.arange(1, n), size=min(nseg - 1, n - 1), replace=False))
    levels = rng.integers(0, 6, size=len(cuts) + 1).astype(np.float64) * 10.0 + 5.0
    out = np.empty(n)
    prev = 0
    for c, lv in zip(list(cuts) + [n], levels):
        out[prev:c] = lv
        prev = c
    return out


CLASSES = {
    "periodic": periodic, "gauge": gauge_walk, "util": utilisation, "saw": sawtooth,
    "constant": constant, "noisy": noisy, "steps": steps,
}


def make(kind, n, seed):
    return np.ascontiguousarray(CLASSES[kind](n, seed), dtype=np.float64)
