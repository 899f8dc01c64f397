"""atsc_b200 -- B200-native (sm_100a) implementation of ATSC's per-frame compressor
selection / fitting loop and decompression, behind the C ABI in include/atsc_gpu.h.

This Python package is only a ctypes binding of libatsc_gpu.so for tests and bench.py; the
product is the shared library (CUDA kernels + C ABI + C++ stream layer) and the `atsc` CLI
built from atsc_b200/host/.  There is no CPU fallback: without the built library or without a
CUDA device every call raises.

Names mirror the reference's operator interface (atsc/src/compressor/mod.rs,
frame/mod.rs, data.rs): Compressor, compress / compress_bounded / decompress,
compress_best, CompressedStream-level compress_data / decompress_data.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

NOOP, FFT, IDW, CONSTANT, POLYNOMIAL, AUTO, RLE = range(7)
COMPRESSOR_NAMES = ["Noop", "FFT", "Idw", "Constant", "Polynomial", "Auto", "RLE"]
TIE_FFT_LOOP, TIE_POLY_LOOP, TIE_SELECT, TIE_FFT_TOPK = 1, 2, 4, 8
ERRORS = {1: "ATSC_ERR_ARG", 2: "ATSC_ERR_CUDA", 3: "ATSC_ERR_CAPACITY", 4: "ATSC_ERR_UNSUPPORTED",
          5: "ATSC_ERR_FORMAT"}

API_SYMBOLS = [
    "atsc_gpu_create", "atsc_gpu_destroy", "atsc_gpu_last_error", "atsc_gpu_host_alloc", "atsc_gpu_host_free",
    "atsc_gpu_compress_frames", "atsc_gpu_decompress_frames", "atsc_plan_chunk_sizes",
    "atsc_gpu_compress_series", "atsc_gpu_decompress_series", "atsc_gpu_launch_count", "atsc_gpu_kernel_ms",
    "atsc_plan_shards", "atsc_wbro_decode", "atsc_wbro_encode", "atsc_csv_read_values", "atsc_gpu_last_call_ms",
    "atsc_vsri_new", "atsc_vsri_free", "atsc_day_elapsed_seconds", "atsc_vsri_update_for_point", "atsc_vsri_min",
    "atsc_vsri_max", "atsc_vsri_sample_count", "atsc_vsri_segment_count", "atsc_vsri_get_sample", "atsc_vsri_get_time",
    "atsc_vsri_get_next_sample", "atsc_vsri_get_previous_sample", "atsc_vsri_is_empty", "atsc_vsri_all_timestamps",
    "atsc_vsri_to_text", "atsc_vsri_from_text",
]
KERNEL_NAMES = ["stats", "poly", "rle", "fft_fwd", "select", "emit", "decode", "host_issue", "fft_small", "fft",
                "front", "reserved11"]


class AtscError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class FrameOut(C.Structure):
    _fields_ = [("compressor", C.c_uint8), ("near_tie", C.c_uint8), ("iterations", C.c_uint16),
                ("payload_len", C.c_uint32), ("payload_off", C.c_uint64), ("error", C.c_double),
                ("cand_error", C.c_double * 3), ("cand_size", C.c_uint32 * 3), ("reserved", C.c_uint32)]


class FrameIn(C.Structure):
    _fields_ = [("compressor", C.c_uint8), ("sample_count", C.c_uint32), ("payload_off", C.c_uint64),
                ("payload_len", C.c_uint32), ("out_off", C.c_uint64)]


_lib = None


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libatsc_gpu.so")


def load_library(build_if_missing=True):
    """Loads libatsc_gpu.so (building it in-tree when stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and _build.is_stale() and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        _build.build()
    if not os.path.exists(path):
        raise AtscError(2, f"{path} is missing: run `python -m atsc_b200.build` (no CPU fallback exists)")
    L = C.CDLL(path)
    vp, u8p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.atsc_gpu_create.restype = C.c_int
    L.atsc_gpu_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.atsc_gpu_destroy.restype = None
    L.atsc_gpu_destroy.argtypes = [vp]
    L.atsc_gpu_last_error.restype = C.c_char_p
    L.atsc_gpu_last_error.argtypes = [vp]
    L.atsc_gpu_host_alloc.restype = vp
    L.atsc_gpu_host_alloc.argtypes = [C.c_uint64]
    L.atsc_gpu_host_free.restype = None
    L.atsc_gpu_host_free.argtypes = [vp]
    L.atsc_gpu_launch_count.restype = C.c_uint64
    L.atsc_gpu_launch_count.argtypes = [vp]
    L.atsc_plan_shards.restype = None
    L.atsc_plan_shards.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p]
    L.atsc_wbro_decode.restype = C.c_int64
    L.atsc_wbro_decode.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
    L.atsc_wbro_encode.restype = C.c_uint64
    L.atsc_wbro_encode.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
    L.atsc_csv_read_values.restype = C.c_int64
    L.atsc_csv_read_values.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.c_char_p, vp, C.c_uint64]
    i32p = C.POINTER(C.c_int32)
    L.atsc_vsri_new.restype = vp
    L.atsc_vsri_new.argtypes = []
    L.atsc_vsri_free.restype = None
    L.atsc_vsri_free.argtypes = [vp]
    L.atsc_day_elapsed_seconds.restype = C.c_int32
    L.atsc_day_elapsed_seconds.argtypes = [C.c_int64]
    L.atsc_vsri_update_for_point.restype = C.c_int
    L.atsc_vsri_update_for_point.argtypes = [vp, C.c_int32]
    for nm in ("atsc_vsri_min", "atsc_vsri_max", "atsc_vsri_sample_count"):
        getattr(L, nm).restype = C.c_int32
        getattr(L, nm).argtypes = [vp]
    L.atsc_vsri_segment_count.restype = C.c_uint64
    L.atsc_vsri_segment_count.argtypes = [vp]
    for nm in ("atsc_vsri_get_sample", "atsc_vsri_get_time", "atsc_vsri_get_next_sample", "atsc_vsri_get_previous_sample"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [vp, C.c_int32, i32p]
    L.atsc_vsri_is_empty.restype = C.c_int
    L.atsc_vsri_is_empty.argtypes = [vp, C.c_int32, C.c_int32]
    L.atsc_vsri_all_timestamps.restype = C.c_uint64
    L.atsc_vsri_all_timestamps.argtypes = [vp, i32p, C.c_uint64]
    L.atsc_vsri_to_text.restype = C.c_uint64
    L.atsc_vsri_to_text.argtypes = [vp, C.c_char_p, C.c_uint64]
    L.atsc_vsri_from_text.restype = vp
    L.atsc_vsri_from_text.argtypes = [C.c_char_p, C.c_uint64]
    L.atsc_gpu_last_call_ms.restype = C.c_double
    L.atsc_gpu_last_call_ms.argtypes = [vp]
    L.atsc_gpu_kernel_ms.restype = None
    L.atsc_gpu_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    L.atsc_gpu_compress_frames.restype = C.c_int
    L.atsc_gpu_compress_frames.argtypes = [vp, vp, u64p, u32p, C.c_uint32, C.c_uint8, C.c_float, C.c_uint32,
                                           C.c_int, C.POINTER(FrameOut), vp, C.c_uint64, u64p]
    L.atsc_gpu_decompress_frames.restype = C.c_int
    L.atsc_gpu_decompress_frames.argtypes = [vp, C.POINTER(FrameIn), C.c_uint32, vp, C.c_uint64, vp]
    L.atsc_plan_chunk_sizes.restype = C.c_uint64
    L.atsc_plan_chunk_sizes.argtypes = [C.c_uint64, u32p, C.c_uint64]
    L.atsc_gpu_compress_series.restype = C.c_int
    L.atsc_gpu_compress_series.argtypes = [vp, vp, u64p, u64p, C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint32,
                                           vp, C.c_uint64, u64p, u64p, u8p]
    L.atsc_gpu_decompress_series.restype = C.c_int
    L.atsc_gpu_decompress_series.argtypes = [vp, vp, u64p, u64p, C.c_uint32, vp, u64p, u64p]
    _lib = L
    return L


class Vsri:
    """vsri::Vsri (vsri/src/lib.rs) through the C ABI; host only."""

    def __init__(self, handle=None):
        self.L = load_library()
        self.h = handle if handle is not None else self.L.atsc_vsri_new()

    @classmethod
    def from_text(cls, text):
        L = load_library()
        b = text.encode() if isinstance(text, str) else bytes(text)
        h = L.atsc_vsri_from_text(b, len(b))
        if not h:
            raise ValueError("malformed VSRI text")
        return cls(h)

    def __del__(self):
        try:
            if self.h:
                self.L.atsc_vsri_free(self.h)
                self.h = None
        except Exception:
            pass

    def update_for_point(self, y):
        return self.L.atsc_vsri_update_for_point(self.h, int(y)) == 0

    def _opt(self, fn, a):
        out = C.c_int32()
        return int(out.value) if fn(self.h, int(a), C.byref(out)) else None

    def get_sample(self, y):
        return self._opt(self.L.atsc_vsri_get_sample, y)

    def get_time(self, x):
        return self._opt(self.L.atsc_vsri_get_time, x)

    def get_next_sample(self, y):
        return self._opt(self.L.atsc_vsri_get_next_sample, y)

    def get_previous_sample(self, y):
        return self._opt(self.L.atsc_vsri_get_previous_sample, y)

    def is_empty(self, t0, t1):
        return bool(self.L.atsc_vsri_is_empty(self.h, int(t0), int(t1)))

    min = property(lambda self: int(self.L.atsc_vsri_min(self.h)))
    max = property(lambda self: int(self.L.atsc_vsri_max(self.h)))
    sample_count = property(lambda self: int(self.L.atsc_vsri_sample_count(self.h)))
    segment_count = property(lambda self: int(self.L.atsc_vsri_segment_count(self.h)))

    def all_timestamps(self):
        n = self.L.atsc_vsri_all_timestamps(self.h, None, 0)
        out = (C.c_int32 * max(int(n), 1))()
        self.L.atsc_vsri_all_timestamps(self.h, out, n)
        return [int(out[i]) for i in range(n)]

    def to_text(self):
        n = self.L.atsc_vsri_to_text(self.h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        self.L.atsc_vsri_to_text(self.h, buf, n)
        return buf.raw[:n].decode()


def day_elapsed_seconds(ts):
    return int(load_library().atsc_day_elapsed_seconds(int(ts)))


def chunk_sizes(n):
    """OptimizerPlan::get_chunks_sizes (optimizer/mod.rs:78-98); host-only, needs no GPU."""
    L = load_library()
    k = L.atsc_plan_chunk_sizes(n, None, 0)
    out = (C.c_uint32 * max(k, 1))()
    L.atsc_plan_chunk_sizes(n, out, k)
    return [int(out[i]) for i in range(k)]


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def plan_shards(frame_len, n_parts):
    """atsc_plan_shards: contiguous frame ranges balanced by sample count (host-only)."""
    L = load_library()
    fl = np.ascontiguousarray(frame_len, dtype=np.uint32)
    first = np.zeros(n_parts + 1, dtype=np.uint32)
    L.atsc_plan_shards(fl.ctypes.data_as(C.POINTER(C.c_uint32)), len(fl), n_parts,
                       first.ctypes.data_as(C.POINTER(C.c_uint32)))
    return [int(x) for x in first]


def wbro_decode(blob):
    """WavBrro::from_file on a file image (wavbrro/src/wavbrro.rs:103); host-only."""
    L = load_library()
    b = np.frombuffer(bytes(blob), dtype=np.uint8)
    n = L.atsc_wbro_decode(_ptr(b), len(b), None, 0)
    if n < 0:
        raise AtscError(5, f"not a WBRO file ({n})")
    out = np.empty(max(n, 1), dtype=np.float64)
    L.atsc_wbro_decode(_ptr(b), len(b), _ptr(out), n)
    return out[:n]


def wbro_encode(samples):
    """WavBrro::to_file_with_data (wavbrro/src/wavbrro.rs:114); host-only."""
    L = load_library()
    a = np.ascontiguousarray(samples, dtype=np.float64)
    if len(a) == 0:
        a = np.zeros(1)
        n = 0
    else:
        n = len(a)
    size = L.atsc_wbro_encode(_ptr(a), n, None, 0)
    out = np.empty(size, dtype=np.uint8)
    L.atsc_wbro_encode(_ptr(a), n, _ptr(out), size)
    return out.tobytes()


def csv_read_values(text, has_header=True, time_field="time", value_field="value"):
    """atsc/src/csv.rs:36-98; host-only."""
    L = load_library()
    t = text.encode() if isinstance(text, str) else bytes(text)
    n = L.atsc_csv_read_values(t, len(t), 1 if has_header else 0, time_field.encode(), value_field.encode(), None, 0)
    if n < 0:
        raise AtscError(5, f"csv error {n}")
    out = np.empty(max(n, 1), dtype=np.float64)
    L.atsc_csv_read_values(t, len(t), 1 if has_header else 0, time_field.encode(), value_field.encode(), _ptr(out), n)
    return out[:n]


class Context:
    """One atsc_ctx (streams + workspaces on the listed devices)."""

    def __init__(self, devices=None):
        self.L = load_library()
        self.h = C.c_void_p()
        if devices is None:
            rc = self.L.atsc_gpu_create(None, 0, C.byref(self.h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.L.atsc_gpu_create(arr, len(devices), C.byref(self.h))
        if rc:
            raise AtscError(rc, "atsc_gpu_create failed (is a CUDA device visible?)")

    def close(self):
        if self.h:
            self.L.atsc_gpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise AtscError(rc, self.L.atsc_gpu_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.L.atsc_gpu_launch_count(self.h))

    @property
    def last_call_ms(self):
        """Device span (CUDA events) of the last compress / decompress call, ms."""
        return float(self.L.atsc_gpu_last_call_ms(self.h))

    def kernel_ms(self, reset=True):
        """CUDA-event milliseconds per kernel since the last reset (dict by kernel name)."""
        arr = (C.c_double * 12)()
        self.L.atsc_gpu_kernel_ms(self.h, arr, 1 if reset else 0)
        return {k: arr[i] for i, k in enumerate(KERNEL_NAMES)}

    # ------------------------------------------------------------------ frame level (C ABI)
    def compress_frames(self, samples, frame_off, frame_len, compressor=AUTO, max_error=0.05, speed=0,
                        bounded=True, payload_cap=None, samples_ptr=None, payload_out=None):
        """atsc_gpu_compress_frames.  `samples` is a float64 numpy array (host) or, with
        samples_ptr, a raw device pointer.  Returns (FrameOut array, payload bytes ndarray)."""
        fo = np.ascontiguousarray(frame_off, dtype=np.uint64)
        fl = np.ascontiguousarray(frame_len, dtype=np.uint32)
        n = len(fl)
        out = (FrameOut * max(n, 1))()
        if payload_out is not None:
            payload, payload_cap = payload_out, len(payload_out)  # caller-owned (e.g. page-locked) buffer
        else:
            if payload_cap is None:
                payload_cap = int(fl.astype(np.uint64).sum()) * 16 + 64 * n + 64
            payload = np.empty(payload_cap, dtype=np.uint8)
        used = C.c_uint64()
        if samples_ptr is None:
            samples = np.ascontiguousarray(samples, dtype=np.float64)
            sp = _ptr(samples)
        else:
            sp = C.c_void_p(samples_ptr)
        rc = self.L.atsc_gpu_compress_frames(self.h, sp, fo.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             fl.ctypes.data_as(C.POINTER(C.c_uint32)), n, compressor,
                                             np.float32(max_error), speed, 1 if bounded else 0, out, _ptr(payload),
                                             payload_cap, C.byref(used))
        self._check(rc)
        return out, payload[:used.value]

    @staticmethod
    def frames_in(frames):
        """(compressor, sample_count, payload_off, payload_len, out_off) tuples -> atsc_frame_in array
        (build it once when the same table is decoded repeatedly)."""
        arr = (FrameIn * max(len(frames), 1))()
        for i, (c, sc, po, pl, oo) in enumerate(frames):
            arr[i].compressor, arr[i].sample_count, arr[i].payload_off = c, sc, po
            arr[i].payload_len, arr[i].out_off = pl, oo
        arr.n_frames = len(frames)
        arr.n_samples = max((oo + sc for _, sc, _, _, oo in frames), default=0)
        return arr

    def decompress_frames(self, frames, payloads, out=None, out_ptr=None, total=None):
        """atsc_gpu_decompress_frames.  frames: list of (compressor, sample_count, payload_off,
        payload_len, out_off), or the array Context.frames_in() made of it."""
        arr = frames if isinstance(frames, C.Array) else self.frames_in(frames)
        n, need = arr.n_frames, arr.n_samples
        payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
        if out_ptr is None:
            if out is None:
                out = np.empty(total if total is not None else need, dtype=np.float64)
            op = _ptr(out)
        else:
            op = C.c_void_p(out_ptr)
        rc = self.L.atsc_gpu_decompress_frames(self.h, arr, n, _ptr(payloads), len(payloads), op)
        self._check(rc)
        return out

    # ------------------------------------------------------------------ reference operator mirror
    def compress(self, compressor, data):
        """Compressor::compress (compressor/mod.rs:63)."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        out, payload = self.compress_frames(data, [0], [len(data)], compressor, 0.0, 0, bounded=False)
        return payload[out[0].payload_off:out[0].payload_off + out[0].payload_len].tobytes()

    def compress_bounded(self, compressor, data, max_error):
        """Compressor::get_compress_bounded_results (compressor/mod.rs:94) -> (bytes, FrameOut).
        max_error is taken as f32 then widened, exactly like frame/mod.rs:65-68."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        out, payload = self.compress_frames(data, [0], [len(data)], compressor, max_error, 0, bounded=True)
        return payload[out[0].payload_off:out[0].payload_off + out[0].payload_len].tobytes(), out[0]

    def compress_best(self, data, max_error, speed=0):
        """CompressorFrame::compress_best (frame/mod.rs:71) -> (compressor, bytes, FrameOut)."""
        b, o = self.compress_bounded(AUTO, data, max_error) if speed == 0 else self._best_speed(data, max_error, speed)
        return o.compressor, b, o

    def _best_speed(self, data, max_error, speed):
        data = np.ascontiguousarray(data, dtype=np.float64)
        out, payload = self.compress_frames(data, [0], [len(data)], AUTO, max_error, speed, bounded=True)
        return payload[out[0].payload_off:out[0].payload_off + out[0].payload_len].tobytes(), out[0]

    def decompress(self, compressor, samples, payload):
        """Compressor::decompress (compressor/mod.rs:109)."""
        p = np.frombuffer(bytes(payload), dtype=np.uint8)
        return self.decompress_frames([(compressor, samples, 0, len(p), 0)], p)

    # ------------------------------------------------------------------ stream level
    def compress_data(self, series, compressor=AUTO, error=3, speed=0, return_ties=False):
        """main.rs:130 compress_data for a list of series -> list of .bro byte strings."""
        series = [np.ascontiguousarray(s, dtype=np.float64) for s in series]
        lens = np.array([len(s) for s in series], dtype=np.uint64)
        offs = np.zeros(len(series), dtype=np.uint64)
        if len(series) > 1:
            offs[1:] = np.cumsum(lens)[:-1]
        flat = np.concatenate(series) if series else np.zeros(0)
        if len(flat) == 0:
            flat = np.zeros(1)
        cap = int(lens.sum()) * 16 + 4096 * len(series) + 4096
        buf = np.empty(cap, dtype=np.uint8)
        bo = np.zeros(len(series), dtype=np.uint64)
        bl = np.zeros(len(series), dtype=np.uint64)
        ties = np.zeros(max(len(series), 1), dtype=np.uint8)
        u64p = C.POINTER(C.c_uint64)
        rc = self.L.atsc_gpu_compress_series(self.h, _ptr(flat), offs.ctypes.data_as(u64p), lens.ctypes.data_as(u64p),
                                             len(series), compressor, error, speed, _ptr(buf), cap,
                                             bo.ctypes.data_as(u64p), bl.ctypes.data_as(u64p),
                                             ties.ctypes.data_as(C.POINTER(C.c_uint8)))
        self._check(rc)
        bros = [buf[int(bo[i]):int(bo[i] + bl[i])].tobytes() for i in range(len(series))]
        return (bros, ties[:len(series)]) if return_ties else bros

    def decompress_data(self, bros):
        """main.rs:168 decompress_data for a list of .bro byte strings -> list of f64 arrays."""
        lens = np.array([len(b) for b in bros], dtype=np.uint64)
        offs = np.zeros(len(bros), dtype=np.uint64)
        if len(bros) > 1:
            offs[1:] = np.cumsum(lens)[:-1]
        buf = np.frombuffer(b"".join(bros), dtype=np.uint8)
        cnt = np.zeros(len(bros), dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        rc = self.L.atsc_gpu_decompress_series(self.h, _ptr(buf), offs.ctypes.data_as(u64p), lens.ctypes.data_as(u64p),
                                               len(bros), None, None, cnt.ctypes.data_as(u64p))
        self._check(rc)
        oo = np.zeros(len(bros), dtype=np.uint64)
        if len(bros) > 1:
            oo[1:] = np.cumsum(cnt)[:-1]
        out = np.empty(max(int(cnt.sum()), 1), dtype=np.float64)
        rc = self.L.atsc_gpu_decompress_series(self.h, _ptr(buf), offs.ctypes.data_as(u64p), lens.ctypes.data_as(u64p),
                                               len(bros), _ptr(out), oo.ctypes.data_as(u64p), cnt.ctypes.data_as(u64p))
        self._check(rc)
        return [out[int(oo[i]):int(oo[i] + cnt[i])].copy() for i in range(len(bros))]
