"""numpy prototype of the register-radix forward engine (fft2.cuh): two Stockham stages per
sub-FFT, W kept column-major [n2][k1], pass 2 pairs row k1 with M1-k1 so the real-input
post-process happens in registers.  Validated against np.fft.rfft."""
import numpy as np


def dft(a, axis=0):
    return np.fft.fft(a, axis=axis)


def forward(x, M1, M2, RA1, RB1, RA2, RB2):
    L = len(x); M = L // 2
    assert M1 * M2 == M and RA1 * RB1 == M1 and RA2 * RB2 == M2
    z = (x[0::2] + 1j * x[1::2]).reshape(M1, M2)          # z[n1][n2]
    W = np.zeros((M2, M1), complex)                       # W[n2][k1]
    # ---- pass 1, per column n2: stage 1 items p < RB1, stage 2 items q < RA1
    for n2 in range(M2):
        y = np.zeros(M1, complex)
        for p in range(RB1):
            a = np.array([z[p + RB1 * t, n2] for t in range(RA1)])
            b = dft(a)
            for q in range(RA1):
                y[RA1 * p + q] = b[q] * np.exp(-2j * np.pi * p * q / M1)
        for q in range(RA1):
            a = np.array([y[q + RA1 * t] for t in range(RB1)])
            b = dft(a)
            for u in range(RB1):
                k1 = q + RA1 * u
                W[n2, k1] = b[u] * np.exp(-2j * np.pi * k1 * n2 / M)
    # ---- pass 2: rows k1 paired with M1 - k1
    X = np.zeros(M + 1, complex)
    def row_stage1(k1):
        y = np.zeros(M2, complex)
        for p in range(RB2):
            a = np.array([W[p + RB2 * t, k1] for t in range(RA2)])
            b = dft(a)
            for q in range(RA2):
                y[RA2 * p + q] = b[q] * np.exp(-2j * np.pi * p * q / M2)
        return y
    def bfly2(y, q):
        return dft(np.array([y[q + RA2 * t] for t in range(RB2)]))   # -> k2 = q + RA2*u
    def post(k, Zk, Zmk):
        w = np.exp(-2j * np.pi * k / L)
        return 0.5 * ((Zk + np.conj(Zmk)) - 1j * w * (Zk - np.conj(Zmk)))
    npairs = (M1 - 1) // 2
    for k1 in range(1, npairs + 1):
        ya, yb = row_stage1(k1), row_stage1(M1 - k1)
        for q in range(RA2):
            A = bfly2(ya, q); B = bfly2(yb, RA2 - 1 - q)
            for u in range(RB2):
                k2 = q + RA2 * u
                k = k1 + M1 * k2
                Zk, Zmk = A[u], B[RB2 - 1 - u]
                X[k] = post(k, Zk, Zmk)
                X[M - k] = post(M - k, Zmk, Zk)
    # self-paired rows
    selfrows = [0] + ([M1 // 2] if M1 % 2 == 0 else [])
    for k1 in selfrows:
        y = row_stage1(k1)
        Zr = np.zeros(M2, complex)
        for q in range(RA2):
            A = bfly2(y, q)
            for u in range(RB2):
                Zr[q + RA2 * u] = A[u]
        for k2 in range(M2):
            k = k1 + M1 * k2
            if k1 == 0:
                pk2 = (M2 - k2) % M2
            else:
                pk2 = M2 - 1 - k2
            X[k] = post(k, Zr[k2], Zr[pk2])
        if k1 == 0:
            X[M] = Zr[0].real - Zr[0].imag
    return X


for (L, M1, M2, f) in [(2 * 18 * 27, 18, 27, (2, 9, 9, 3)), (2 * 36 * 27, 36, 27, (4, 9, 3, 9)), (2 * 9 * 27, 9, 27, (1, 9, 9, 3)),
                       (2 * 16 * 18 * 27, 288, 27, (16, 18, 9, 3))]:
    rng = np.random.default_rng(L)
    x = rng.standard_normal(L)
    X = forward(x, M1, M2, *f)
    ref = np.fft.rfft(x)
    print(L, M1, M2, np.abs(X - ref).max())
    assert np.abs(X - ref).max() < 1e-9
print("ok")
