"""ctypes binding of the CPU oracle (oracle/libatsc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by atsc_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ODIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ODIR, "libatsc_oracle.so")

NOOP, FFT, IDW, CONSTANT, POLYNOMIAL, AUTO, RLE = range(7)
NAMES = ["Noop", "FFT", "Idw", "Constant", "Polynomial", "Auto", "RLE"]


def build():
    src = os.path.join(_ODIR, "atsc_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ODIR, "-s"])
    return _SO


def build_native():
    """The same C file built -O3 -march=native ON THIS MACHINE (timed CPU baseline only; not the checker).
    Returns the path, or None when the build is not possible (then the portable library is timed)."""
    so = os.path.join(_ODIR, "libatsc_oracle_native.so")
    try:
        subprocess.check_call(["make", "-C", _ODIR, "-s", "-B", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return so if os.path.exists(so) else None
    except Exception:
        return None


_lib = None
_native = False


def use_native():
    """Switches this process to the -O3 -march=native build (bench.py's CPU legs). True when it took effect."""
    global _lib, _native
    so = build_native()
    if so is None:
        return False
    _lib = None
    _native = True
    lib(so)
    return True


def lib(path=None):
    global _lib
    if _lib is None:
        _lib = C.CDLL(path or build())
        L = _lib
        dp = C.POINTER(C.c_double)
        bp = C.POINTER(C.c_uint8)
        L.atsc_oracle_next_size.restype = C.c_uint64
        L.atsc_oracle_next_size.argtypes = [C.c_uint64]
        L.atsc_oracle_prev_power_of_two.restype = C.c_uint64
        L.atsc_oracle_prev_power_of_two.argtypes = [C.c_uint64]
        L.atsc_oracle_round_f64.restype = C.c_double
        L.atsc_oracle_round_f64.argtypes = [C.c_double, C.c_uint32]
        L.atsc_oracle_round_and_limit_f64.restype = C.c_double
        L.atsc_oracle_round_and_limit_f64.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint32]
        L.atsc_oracle_mape.restype = C.c_double
        L.atsc_oracle_mape.argtypes = [dp, dp, C.c_uint64]
        L.atsc_oracle_stats.restype = None
        L.atsc_oracle_stats.argtypes = [dp, C.c_uint64, dp, dp, dp, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.atsc_oracle_compress.restype = C.c_int64
        L.atsc_oracle_compress.argtypes = [C.c_int, dp, C.c_uint64, bp, C.c_uint64]
        L.atsc_oracle_compress_bounded.restype = C.c_int64
        L.atsc_oracle_compress_bounded.argtypes = [C.c_int, dp, C.c_uint64, C.c_double, bp, C.c_uint64,
                                                   dp, C.POINTER(C.c_int)]
        L.atsc_oracle_compress_best.restype = C.c_int64
        L.atsc_oracle_compress_best.argtypes = [dp, C.c_uint64, C.c_float, C.c_uint32, C.POINTER(C.c_int),
                                                bp, C.c_uint64, dp, C.POINTER(C.c_uint64)]
        L.atsc_oracle_decompress.restype = C.c_int64
        L.atsc_oracle_decompress.argtypes = [C.c_int, C.c_uint64, bp, C.c_uint64, dp]
        L.atsc_oracle_fft_set.restype = C.c_int64
        L.atsc_oracle_fft_set.argtypes = [dp, C.c_uint64, C.c_uint64, bp, C.c_uint64]
        L.atsc_oracle_gibbs_sizing.restype = C.c_uint64
        L.atsc_oracle_gibbs_sizing.argtypes = [dp, C.c_uint64, dp, C.c_uint64]
        L.atsc_oracle_fft_c32.restype = None
        L.atsc_oracle_fft_c32.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int]
        L.atsc_oracle_clean_data.restype = C.c_uint64
        L.atsc_oracle_clean_data.argtypes = [dp, C.c_uint64, dp]
        L.atsc_oracle_chunk_sizes.restype = C.c_uint64
        L.atsc_oracle_chunk_sizes.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]
        L.atsc_oracle_compress_stream.restype = C.c_int64
        L.atsc_oracle_compress_stream.argtypes = [dp, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, bp,
                                                  C.c_uint64, C.POINTER(C.c_int), C.c_uint64]
        L.atsc_oracle_decompress_stream.restype = C.c_int64
        L.atsc_oracle_decompress_stream.argtypes = [bp, C.c_uint64, dp, C.c_uint64]
        L.atsc_oracle_compress_batch.restype = C.c_int
        L.atsc_oracle_compress_batch.argtypes = [dp, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32,
                                                 C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        L.atsc_oracle_decompress_batch.restype = C.c_int
        L.atsc_oracle_decompress_batch.argtypes = [bp, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64,
                                                   C.c_int, dp]
        L.atsc_oracle_set_idw_variant.restype = None
        L.atsc_oracle_set_idw_variant.argtypes = [C.c_int]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _cap(n):
    return int(n) * 16 + 4096


def next_size(n):
    return int(lib().atsc_oracle_next_size(n))


def stats(data):
    d = _f64(data)
    mn, mx, mean = C.c_double(), C.c_double(), C.c_double()
    il, al = C.c_uint64(), C.c_uint64()
    bd, fr = C.c_int(), C.c_int()
    lib().atsc_oracle_stats(_dp(d), len(d), C.byref(mn), C.byref(mx), C.byref(mean), C.byref(il),
                            C.byref(al), C.byref(bd), C.byref(fr))
    return dict(min=mn.value, max=mx.value, mean=mean.value, min_loc=il.value, max_loc=al.value,
                bitdepth=bd.value, fractional=bool(fr.value))


def mape(orig, gen):
    a, b = _f64(orig), _f64(gen)
    assert len(a) == len(b)
    return lib().atsc_oracle_mape(_dp(a), _dp(b), len(a))


def compress(comp, data):
    """Compressor::compress (compressor/mod.rs:63)."""
    d = _f64(data)
    out = np.zeros(_cap(len(d)), dtype=np.uint8)
    n = lib().atsc_oracle_compress(comp, _dp(d), len(d), _bp(out), len(out))
    assert n >= 0, n
    return out[:n].tobytes()


def compress_bounded(comp, data, max_error):
    """Compressor::get_compress_bounded_results -> (bytes, error, iterations)."""
    d = _f64(data)
    out = np.zeros(_cap(len(d)), dtype=np.uint8)
    err = C.c_double()
    it = C.c_int()
    n = lib().atsc_oracle_compress_bounded(comp, _dp(d), len(d), float(max_error), _bp(out), len(out),
                                           C.byref(err), C.byref(it))
    assert n >= 0, n
    return out[:n].tobytes(), err.value, it.value


def compress_best(data, max_error_f32, speed=0):
    """CompressorFrame::compress_best -> (compressor, bytes, cand_err[3], cand_size[3])."""
    d = _f64(data)
    out = np.zeros(_cap(len(d)), dtype=np.uint8)
    comp = C.c_int()
    ce = (C.c_double * 3)()
    cs = (C.c_uint64 * 3)()
    n = lib().atsc_oracle_compress_best(_dp(d), len(d), np.float32(max_error_f32), speed, C.byref(comp),
                                        _bp(out), len(out), ce, cs)
    assert n >= 0, n
    return comp.value, out[:n].tobytes(), list(ce), list(cs)


def decompress(comp, samples, payload):
    p = np.frombuffer(payload, dtype=np.uint8).copy()
    if comp == NOOP:
        out = np.zeros(max(samples, len(p)), dtype=np.float64)
    else:
        out = np.zeros(samples, dtype=np.float64)
    n = lib().atsc_oracle_decompress(comp, samples, _bp(p), len(p), _dp(out))
    assert n >= 0, n
    return out[:n]


def fft_set(data, freqs):
    d = _f64(data)
    out = np.zeros(_cap(len(d)), dtype=np.uint8)
    n = lib().atsc_oracle_fft_set(_dp(d), len(d), freqs, _bp(out), len(out))
    assert n >= 0
    return out[:n].tobytes()


def gibbs_sizing(data):
    d = _f64(data)
    out = np.zeros(len(d) * 2 + 16, dtype=np.float64)
    n = lib().atsc_oracle_gibbs_sizing(_dp(d), len(d), _dp(out), len(out))
    return out[:n]


def fft_c32(z, inverse=False):
    a = np.ascontiguousarray(np.asarray(z, dtype=np.complex64))
    buf = a.view(np.float32).copy()
    lib().atsc_oracle_fft_c32(buf.ctypes.data_as(C.POINTER(C.c_float)), len(a), int(inverse))
    return buf.view(np.complex64)


def clean_data(data):
    d = _f64(data)
    out = np.zeros(len(d), dtype=np.float64)
    n = lib().atsc_oracle_clean_data(_dp(d), len(d), _dp(out))
    return out[:n]


def chunk_sizes(n):
    k = lib().atsc_oracle_chunk_sizes(n, None, 0)
    out = (C.c_uint64 * max(k, 1))()
    lib().atsc_oracle_chunk_sizes(n, out, k)
    return [int(out[i]) for i in range(k)]


def compress_stream(samples, compressor=AUTO, error_pct=5, speed=0):
    """main.rs:130 compress_data -> (.bro bytes, [frame compressor ids])."""
    d = _f64(samples)
    out = np.zeros(_cap(len(d)) + 64 * (len(d) // 512 + 8), dtype=np.uint8)
    nfr = max(len(chunk_sizes(len(d))), 1)
    fc = (C.c_int * nfr)()
    n = lib().atsc_oracle_compress_stream(_dp(d), len(d), compressor, error_pct, speed, _bp(out), len(out),
                                          fc, nfr)
    assert n >= 0, n
    return out[:n].tobytes(), [fc[i] for i in range(len(chunk_sizes(len(clean_data(d)))))]


def decompress_stream(bro):
    p = np.frombuffer(bro, dtype=np.uint8).copy()
    n = lib().atsc_oracle_decompress_stream(_bp(p), len(p), None, 0)
    assert n >= 0, n
    out = np.zeros(n, dtype=np.float64)
    m = lib().atsc_oracle_decompress_stream(_bp(p), len(p), _dp(out), n)
    assert m == n, (m, n)
    return out


def compress_batch(samples2d, compressor=AUTO, error_pct=5, speed=0, threads=1):
    a = _f64(samples2d)
    ns, sl = a.shape
    sizes = np.zeros(ns, dtype=np.uint64)
    rc = lib().atsc_oracle_compress_batch(_dp(a), sl, ns, compressor, error_pct, speed, threads,
                                          sizes.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0
    return sizes


def set_idw_variant(v):
    lib().atsc_oracle_set_idw_variant(v)
