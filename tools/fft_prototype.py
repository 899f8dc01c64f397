"""numpy prototype of the GPU FFT engine's index math (validated against np.fft).
Mirrors atsc_b200/csrc/fft.cuh: Stockham DIF stages, four-step M = M1*M2 with permuted
spectrum storage, real-input trick for even L, sparse-spectrum scatter coefficients."""
import numpy as np


def radices(n):
    r = []
    while n % 4 == 0:
        r.append(4); n //= 4
    while n % 2 == 0:
        r.append(2); n //= 2
    while n % 3 == 0:
        r.append(3); n //= 3
    assert n == 1
    return r


def stockham(x, inverse):
    """x: [len, batch] complex; DIF autosort, generic radix."""
    ln = x.shape[0]
    sign = 1.0 if inverse else -1.0
    n, s = ln, 1
    x = x.copy()
    for r in radices(ln):
        m = n // r
        y = np.zeros_like(x)
        for b in range(m * s):
            p, q = divmod(b, s)
            a = [x[q + s * (p + t * m)] for t in range(r)]
            for u in range(r):
                acc = 0
                for t in range(r):
                    acc = acc + a[t] * np.exp(sign * 2j * np.pi * t * u / r)
                y[q + s * (r * p + u)] = acc * np.exp(sign * 2j * np.pi * p * u / n)
        x = y
        n, s = m, s * r
    return x


def fourstep_forward(z, M1, M2):
    """natural order z[n1*M2+n2] -> permuted S[k1][k2] = Z[k1 + M1*k2]"""
    M = M1 * M2
    S = z.reshape(M1, M2).copy()
    S = stockham(S, False)  # column pass: FFT over axis0 (len M1), batch = columns
    k1 = np.arange(M1)[:, None]
    n2 = np.arange(M2)[None, :]
    S = S * np.exp(-2j * np.pi * k1 * n2 / M)
    S = stockham(S.T.copy(), False).T  # row pass: FFT over n2 (len M2), batch = rows
    return S


def fourstep_inverse(S, M1, M2):
    """permuted S[k1][k2] -> natural z[n1*M2+n2] (unnormalised)"""
    M = M1 * M2
    T = stockham(S.T.copy(), True).T  # row pass over k2
    k1 = np.arange(M1)[:, None]
    n2 = np.arange(M2)[None, :]
    T = T * np.exp(+2j * np.pi * k1 * n2 / M)
    T = stockham(T, True)  # column pass over k1
    return T.reshape(-1)


def test(L, M1, M2):
    rng = np.random.default_rng(L)
    x = rng.standard_normal(L)
    M = L // 2
    assert M1 * M2 == M
    # forward real trick
    z = x[0::2] + 1j * x[1::2]
    S = fourstep_forward(z, M1, M2)
    def Zs(k):
        k = k % M
        return S[k % M1, k // M1]
    X = np.zeros(M + 1, complex)
    for k in range(M + 1):
        Zk, Zmk = Zs(k), np.conj(Zs(M - k))
        twL = np.exp(-2j * np.pi * k / L)
        X[k] = 0.5 * ((Zk + Zmk) - 1j * twL * (Zk - Zmk))
    ref = np.fft.fft(x)[:M + 1]
    e1 = np.abs(X - ref).max()
    # inverse with sparse subset
    sel = rng.choice(M + 1, size=max(3, (M + 1) // 5), replace=False)
    sel = np.unique(np.concatenate([sel, [0, M]]))
    Zt = np.zeros((M1, M2), complex)
    for p in sel:
        zc = X[p]
        c, s = np.cos(2 * np.pi * p / L), np.sin(2 * np.pi * p / L)
        if p == 0:
            cD = zc.real * (1 + 1j); Zt[0, 0] += cD
        elif p == M:
            cM = zc.real * (1 - 1j); Zt[0, 0] += cM
        else:
            cD = zc * ((1 - s) + 1j * c)
            Zt[p % M1, p // M1] += cD
            cM = np.conj(zc) * ((1 + s) + 1j * c)
            k = M - p
            Zt[k % M1, k // M1] += cM
    zz = fourstep_inverse(Zt, M1, M2)
    xr = np.empty(L)
    xr[0::2] = zz.real
    xr[1::2] = zz.imag
    Xfull = np.zeros(L, complex)
    for p in sel:
        if p == 0 or p == M:
            Xfull[p] = X[p].real if p == 0 else 0
            if p == M:
                Xfull[M] = np.conj(X[M])
        else:
            Xfull[p] = X[p]
            Xfull[L - p] = np.conj(X[p])
    refx = (np.fft.ifft(Xfull) * L).real
    e2 = np.abs(xr - refx).max()
    print(L, M1, M2, "fwd err", e1, "inv err", e2)
    assert e1 < 1e-9 and e2 < 1e-9


def test_complex(L, M1, M2):
    rng = np.random.default_rng(L)
    x = rng.standard_normal(L)
    S = fourstep_forward(x.astype(complex), M1, M2)
    X = np.array([S[k % M1, k // M1] for k in range(L // 2 + 1)])
    ref = np.fft.fft(x)[:L // 2 + 1]
    e1 = np.abs(X - ref).max()
    sel = np.unique(np.concatenate([rng.choice(L // 2 + 1, size=5, replace=False), [0]]))
    Zt = np.zeros((M1, M2), complex)
    Xfull = np.zeros(L, complex)
    for p in sel:
        if p == 0:
            Zt[0, 0] = X[0].real; Xfull[0] = X[0].real
        else:
            Zt[p % M1, p // M1] = X[p]
            k = L - p
            Zt[k % M1, k // M1] = np.conj(X[p])
            Xfull[p] = X[p]; Xfull[L - p] = np.conj(X[p])
    zz = fourstep_inverse(Zt, M1, M2)
    e2 = np.abs(zz.real - (np.fft.ifft(Xfull) * L).real).max()
    print("complex", L, M1, M2, e1, e2)
    assert e1 < 1e-9 and e2 < 1e-9


if __name__ == "__main__":
    test(144, 9, 8)
    test(576, 18, 16)
    test(1152, 24, 24)
    test(4374, 27, 81)
    test_complex(243, 9, 27)
    test_complex(2187, 27, 81)
    print("ok")
