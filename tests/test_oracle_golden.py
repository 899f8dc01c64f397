"""Pins the CPU oracle against every golden vector the reference's own tests hold for
the hot path and against the reference binary's own output
arrays embedded in atsc/demo/*.html (tests/golden/demo_html.npz, extracted by
tests/golden/make_golden.py).  CPU only.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def B(x):
    return bytes(x)


# ---------------------------------------------------------------- utils/mod.rs:81-101
def test_next_size_table():
    assert O.next_size(2048) == 2187
    assert O.next_size(512) == 576
    assert O.next_size(256) == 288
    assert O.next_size(128) == 144
    assert O.next_size(12432) == 13122


def test_round_and_limit():
    L = O.lib()
    assert L.atsc_oracle_round_and_limit_f64(3.0, 2.0, 4.0, 1) == 3.0
    assert L.atsc_oracle_round_and_limit_f64(5.0, 2.0, 4.0, 1) == 4.0
    assert L.atsc_oracle_round_and_limit_f64(1.0, 2.0, 4.0, 1) == 2.0
    assert L.atsc_oracle_round_and_limit_f64(3.123452312, 2.0, 4.0, 3) == 3.123


# ---------------------------------------------------------------- utils/error.rs:171-180
def test_mape():
    v1 = [1.0, 2.0, 3.0, 4.0, 5.0]
    v2 = [2.5, 4.0, 6.0, 8.0, 10.0]
    assert O.mape(v1, v1) == 0.0
    assert O.mape(v1, v2) == 1.1
    assert O.mape([1.0], [1.1]) < 0.101


# ---------------------------------------------------------------- optimizer/utils.rs:166-203
def test_stats():
    s = O.stats([1.0, 1.0, 1.0])
    assert (s["bitdepth"], s["min"], s["max"], s["mean"], s["min_loc"], s["max_loc"], s["fractional"]) == \
        (3, 1.0, 1.0, 1.0, 0, 0, False)
    s = O.stats([1.0, 4.0, 7.0])
    assert (s["bitdepth"], s["min"], s["max"], s["mean"], s["min_loc"], s["max_loc"], s["fractional"]) == \
        (3, 1.0, 7.0, 4.0, 0, 2, False)
    s = O.stats([1.5, 4.5, 9.0])
    assert (s["bitdepth"], s["min"], s["max"], s["mean"], s["min_loc"], s["max_loc"], s["fractional"]) == \
        (0, 1.5, 9.0, 5.0, 0, 2, True)


def test_stats_bitdepth_ladder():
    assert O.stats([0.0, 255.0])["bitdepth"] == 3
    assert O.stats([0.0, 256.0])["bitdepth"] == 2
    assert O.stats([-1.0, 5.0])["bitdepth"] == 2
    assert O.stats([-32768.0, 32767.0])["bitdepth"] == 2
    assert O.stats([-32769.0, 5.0])["bitdepth"] == 1
    assert O.stats([0.0, 32768.0])["bitdepth"] == 1
    assert O.stats([0.0, 2147483647.0])["bitdepth"] == 1
    assert O.stats([0.0, 2147483648.0])["bitdepth"] == 0
    assert O.stats([0.0, 0.5])["bitdepth"] == 0


# ---------------------------------------------------------------- optimizer/mod.rs:150-165
def test_chunk_sizes():
    assert O.chunk_sizes(131072 * 3 + 1765) == [131072, 131072, 131072, 1024, 512, 229]
    assert O.chunk_sizes(31) == [31]
    assert O.chunk_sizes(2048) == [2048]
    assert O.chunk_sizes(12032) == [8192, 2048, 1024, 512, 256]
    assert len(O.chunk_sizes(2049)) == 2
    assert len(O.chunk_sizes(132671)) == 4
    assert O.chunk_sizes(0) == []


# ---------------------------------------------------------------- constant.rs:150-178
def test_constant():
    assert O.compress(O.CONSTANT, [1.0] * 5) == B([30, 3, 1])
    assert O.compress(O.CONSTANT, [1.23456] * 5) == B([30, 0, 56, 50, 143, 252, 193, 192, 243, 63])
    out = O.decompress(O.CONSTANT, 5, B([30, 3, 1]))
    assert list(out) == [1.0] * 5


# ---------------------------------------------------------------- noop.rs:89-121
def test_noop():
    assert O.compress(O.NOOP, [1.0] * 5) == B([250, 5, 2, 2, 2, 2, 2])
    v = [1.0, 2.0, 3.0, 4.0, 1.0]
    assert list(O.decompress(O.NOOP, 5, O.compress(O.NOOP, v))) == v
    assert list(O.decompress(O.NOOP, 4, O.compress(O.NOOP, [1.5, 2.7, 3.3, 4.9]))) == [2.0, 3.0, 3.0, 5.0]


# ---------------------------------------------------------------- rle.rs:263-321
RLE_CASES = [
    ([1.0] * 512, [60, 3, 1, 1, 1, 0]),
    ([1.0, 2.0, 2.0, 3.0, 3.0, 3.0, 4.0, 4.0, 4.0, 4.0, 5.0, 5.0, 5.0, 5.0, 5.0],
     [60, 3, 5, 1, 1, 0, 2, 1, 1, 3, 1, 3, 4, 1, 6, 5, 1, 10]),
    ([1.0, 1.0, 1.0, 1.0, 2.0, 2.0, 1.0, 1.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0, 3.0, 3.0, 3.0, 3.0, 1.0, 1.0,
      1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0],
     [60, 3, 3, 1, 3, 0, 6, 18, 2, 2, 4, 10, 3, 1, 12]),
    ([1.23456] * 5, [60, 0, 1, 56, 50, 143, 252, 193, 192, 243, 63, 1, 0]),
]


@pytest.mark.parametrize("raw,enc", RLE_CASES)
def test_rle_roundtrip(raw, enc):
    assert O.compress(O.RLE, raw) == B(enc)
    assert list(O.decompress(O.RLE, len(raw), B(enc))) == raw


def test_index_rle_beats_regular():
    v = ([1.0] + [0.0] * 9) * 3 + [1.0]
    enc = O.compress(O.RLE, v)
    assert len(enc) < 16
    assert list(O.decompress(O.RLE, len(v), enc)) == v


# ---------------------------------------------------------------- polynomial.rs:436-599
V12 = [1.0, 0.0, 1.0, 1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
V17 = [1.0, 1.0, 1.0, 1.0, 2.0, 3.0, 5.0, 1.0, 2.0, 7.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
LIN12 = [float(i) for i in range(1, 13)]


def test_polynomial_bytes():
    assert O.compress(O.POLYNOMIAL, V12) == B(
        [0, 3, 4, 1, 2, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 20, 64, 4])
    v = V12[:-1] + [500.0]
    assert O.compress(O.POLYNOMIAL, v) == B(
        [0, 2, 4, 2, 4, 6, 251, 232, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 127, 64, 4])
    v = [40001.0, 40000.0, 40001.0, 40001.0, 40002.0, 40001.0, 40001.0, 40001.0, 40003.0, 40001.0,
         40001.0, 40005.0]
    assert O.compress(O.POLYNOMIAL, v) == B(
        [0, 1, 4, 252, 130, 56, 1, 0, 252, 132, 56, 1, 0, 252, 134, 56, 1, 0, 252, 138, 56, 1, 0, 0, 0,
         0, 0, 0, 136, 227, 64, 0, 0, 0, 0, 160, 136, 227, 64, 4])
    v = [1.1, 0.1, 1.1, 1.1, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
    assert O.compress(O.POLYNOMIAL, v) == B(
        [0, 0, 4, 154, 153, 153, 153, 153, 153, 241, 63, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0, 0, 0, 8,
         64, 0, 0, 0, 0, 0, 0, 20, 64, 154, 153, 153, 153, 153, 153, 185, 63, 0, 0, 0, 0, 0, 0, 20, 64, 4])
    assert O.compress(O.IDW, V12) == B(
        [1, 3, 4, 1, 2, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 20, 64, 4])
    assert O.compress(O.POLYNOMIAL, [1.0] * 4) == B(
        [0, 3, 0, 0, 0, 0, 0, 0, 0, 240, 63, 0, 0, 0, 0, 0, 0, 240, 63, 1])
    assert O.compress(O.IDW, [1.0] * 4) == B(
        [1, 3, 0, 0, 0, 0, 0, 0, 0, 240, 63, 0, 0, 0, 0, 0, 0, 240, 63, 1])


def test_polynomial_values():
    out = O.decompress(O.POLYNOMIAL, 17, O.compress(O.POLYNOMIAL, V17))
    assert list(out) == [1.0, 1.4, 1.8, 2.2, 2.6, 3.0, 2.824, 2.392, 1.848, 1.336, 1.0, 1.0, 1.0, 1.0,
                         1.0, 1.0, 5.0]
    out = O.decompress(O.POLYNOMIAL, 12, O.compress(O.POLYNOMIAL, LIN12))
    assert list(out) == LIN12


def test_idw_values():
    out = O.decompress(O.IDW, 17, O.compress(O.IDW, V17))
    assert list(out) == [1.0, 1.13167, 1.62573, 2.32782, 2.83429, 3.0, 2.8335, 2.34163, 1.68979, 1.184,
                         1.0, 1.18933, 1.64488, 1.9634, 1.77047, 1.0, 5.0]
    out = O.decompress(O.IDW, 12, O.compress(O.IDW, LIN12))
    assert list(out) == [1.0, 1.62873, 3.51429, 4.84995, 5.0, 5.40622, 7.05871, 8.64807, 9.0, 9.37719,
                         11.18119, 12.0]


def test_poly_allowed_error():
    b, err, _ = O.compress_bounded(O.POLYNOMIAL, V17, 0.05)
    assert O.mape(V17, O.decompress(O.POLYNOMIAL, 17, b)) <= 0.05
    b, err, _ = O.compress_bounded(O.IDW, V17, 0.02)
    assert O.mape(V17, O.decompress(O.IDW, 17, b)) <= 0.02


# ---------------------------------------------------------------- fft.rs:551-626
F12 = [1.0, 1.0, 1.0, 1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]


def test_fft_bytes():
    assert O.fft_set(F12, 2) == B(
        [15, 2, 0, 0, 0, 152, 65, 0, 0, 0, 0, 4, 0, 0, 96, 192, 102, 144, 138, 64, 0, 0, 160, 64, 0, 0,
         128, 63])


def test_fft_lossless_and_lossy():
    out = O.decompress(O.FFT, 12, O.fft_set(F12, 12))
    assert list(out) == F12
    out = O.decompress(O.FFT, 12, O.compress(O.FFT, F12))
    assert list(out) == [1.0, 1.87201, 2.25, 1.0, 1.82735, 1.689, 1.82735, 1.0, 2.75, 1.189, 1.0, 3.311]


def test_fft_allowed_error():
    b, err, _ = O.compress_bounded(O.FFT, F12, 0.01)
    assert O.mape(F12, O.decompress(O.FFT, 12, b)) <= 0.01


def test_fft_gibbs_sizing():
    v = [2.0] * 2048
    v[0] = 1.0
    v[2047] = 3.0
    g = O.gibbs_sizing(v)
    assert len(g) == 2187 and g[2] == 1.0 and g[2185] == 3.0


def test_fft_static_and_trim():
    v = [1.0] * 1024
    b = O.compress(O.FFT, v)
    assert b[1] == 0  # zero frequencies
    assert list(O.decompress(O.FFT, 1024, b)) == v


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 12, 64, 97, 127, 144, 243, 576, 2187, 17496])
def test_fft_matches_numpy(n):
    rng = np.random.default_rng(n)
    z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.fft(z.astype(np.complex128))
    got = O.fft_c32(z)
    tol = 3e-7 * np.log2(max(n, 2)) * np.abs(ref).max() + 1e-6
    assert np.abs(got - ref).max() < tol
    refi = np.fft.ifft(z.astype(np.complex128)) * n
    goti = O.fft_c32(z, inverse=True)
    assert np.abs(goti - refi).max() < tol


# ---------------------------------------------------------------- data.rs:145-154, header.rs
def test_stream_bytes():
    bro, comps = O.compress_stream([1.0] * 1024, compressor=O.CONSTANT)
    assert bro == B([66, 82, 82, 79, 1, 0, 0, 0, 1, 1, 41, 251, 0, 4, 3, 3, 30, 3, 1])
    assert list(O.decompress_stream(bro)) == [1.0] * 1024


def test_stream_rejects_bad_header():
    bro, _ = O.compress_stream([1.0] * 1024, compressor=O.CONSTANT)
    bad = bytearray(bro)
    bad[4] = 9
    p = np.frombuffer(bytes(bad), dtype=np.uint8).copy()
    import ctypes as C
    assert O.lib().atsc_oracle_decompress_stream(p.ctypes.data_as(C.POINTER(C.c_uint8)), len(p), None, 0) == -3
    bad = bytearray(bro)
    bad[0] = 0
    p = np.frombuffer(bytes(bad), dtype=np.uint8).copy()
    assert O.lib().atsc_oracle_decompress_stream(p.ctypes.data_as(C.POINTER(C.c_uint8)), len(p), None, 0) == -2


# ---------------------------------------------------------------- demo HTML: reference binary output
@pytest.fixture(scope="module")
def demo():
    return np.load(os.path.join(G, "demo_html.npz"))


@pytest.mark.parametrize("err", [1, 3])
@pytest.mark.parametrize("name", ["heap", "memory", "csv_iowait"])
@pytest.mark.parametrize("which,comp", [("polyData", O.POLYNOMIAL), ("idwData", O.IDW)])
def test_demo_poly_idw_bit_exact(demo, err, name, which, comp):
    """`atsc --compressor polynomial|idw --error E` then `atsc -u --verbose`
    (atsc/demo/run_demo.sh:15-22): oracle output must equal the reference's, bit for bit."""
    x = demo[f"e{err}_{name}_inputData"]
    want = demo[f"e{err}_{name}_{which}"]
    bro, _ = O.compress_stream(x, compressor=comp, error_pct=err)
    got = O.decompress_stream(bro)
    assert len(got) == len(want)
    assert np.array_equal(got, want), f"{np.sum(got != want)} of {len(want)} differ"


@pytest.mark.parametrize("err", [1, 3])
@pytest.mark.parametrize("name", ["heap", "memory", "csv_iowait"])
def test_demo_fft_within_f32_noise(demo, err, name):
    """FFT frames: rustfft's f32 butterfly order is build dependent; pinned to f32 noise
    (SURVEY.md H1): |ours - ref| <= 1e-5 + 4 * 2^-24 * log2(L) * max|x|."""
    x = demo[f"e{err}_{name}_inputData"]
    want = demo[f"e{err}_{name}_fftData"]
    bro, _ = O.compress_stream(x, compressor=O.FFT, error_pct=err)
    got = O.decompress_stream(bro)
    assert len(got) == len(want)
    tol = 1e-5 + 4 * 2.0 ** -24 * np.log2(2187) * np.abs(x[np.isfinite(x)]).max()
    assert np.abs(got - want).max() <= tol
    if name == "csv_iowait":
        assert np.array_equal(got, want)


def test_demo_iteration_counts(demo):
    """k / iteration counts quoted in SURVEY.md section 8c for the heap fixture, -e 3."""
    x = demo["e3_heap_inputData"]
    frames = [x[:2048], x[2048:2560], x[2560:]]
    e = float(np.float32(3) / np.float32(100.0))
    res = [O.compress_bounded(O.FFT, f, e) for f in frames]
    nfreq = [r[0][1] if r[0][1] < 251 else int.from_bytes(r[0][2:4], "little") for r in res]
    iters = [r[2] for r in res]
    assert nfreq == [200, 5, 11] and iters == [23, 1, 9]
