"""Checks the radix-3/4/8/9 butterfly decompositions used by tile_fft (fft.cuh) against np.fft."""
import numpy as np

def dft3(a, inv):
    c = 0.86602540378443864676
    t = a[1] + a[2]; d = a[1] - a[2]
    m = a[0] - 0.5 * t
    rot = (1j * c * d) if inv else (-1j * c * d)
    return [a[0] + t, m + rot, m - rot]

def dft4(a, inv):
    t0, t1, t2, t3 = a[0] + a[2], a[0] - a[2], a[1] + a[3], a[1] - a[3]
    j3 = 1j * t3 if inv else -1j * t3
    return [t0 + t2, t1 + j3, t0 - t2, t1 - j3]

def dft8(a, inv):
    s = -1.0 if not inv else 1.0
    r = 0.70710678118654752440
    w = [1, r * (1 + s * 1j), s * 1j, r * (-1 + s * 1j)]
    lo = [a[j] + a[j + 4] for j in range(4)]
    hi = [(a[j] - a[j + 4]) * w[j] for j in range(4)]
    e = dft4(lo, inv); o = dft4(hi, inv)
    y = [0] * 8
    for k2 in range(4):
        y[2 * k2] = e[k2]; y[2 * k2 + 1] = o[k2]
    return y

def dft9(a, inv):
    sg = 1.0 if inv else -1.0
    w9 = [np.exp(sg * 2j * np.pi * k / 9) for k in range(5)]
    t = [dft3([a[j], a[j + 3], a[j + 6]], inv) for j in range(3)]   # t[j][k1]
    t[1][1] *= w9[1]; t[1][2] *= w9[2]; t[2][1] *= w9[2]; t[2][2] *= w9[4]
    y = [0] * 9
    for k1 in range(3):
        z = dft3([t[0][k1], t[1][k1], t[2][k1]], inv)
        for k2 in range(3):
            y[k1 + 3 * k2] = z[k2]
    return y

rng = np.random.default_rng(0)
for n, f in ((3, dft3), (4, dft4), (8, dft8), (9, dft9)):
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.allclose(f(list(x), False), np.fft.fft(x)), n
    assert np.allclose(f(list(x), True), np.fft.ifft(x) * n), n
print("radix butterflies ok")
