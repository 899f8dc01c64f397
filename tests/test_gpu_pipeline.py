"""GPU tests of the pieces around the kernels: chunked statistics on awkward values, waves
rotating over the engines, the payload-overflow re-emit, the small-frame FFT kernel on every
short length, and the decoder's fast polynomial expansion.  Everything is compared with the CPU
oracle (bit-exact unless the frame is an FFT frame)."""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O
from test_gpu_parity import fft_tol, parse_fft, run_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    c = atsc_b200.Context()
    yield c
    c.close()


def awkward_frames():
    """Frames that exercise the stats pass: zero extremes of either sign, values below 2^-64, run
    ends on both sides of the varint thresholds (251, 65536), chunk edges (32768), odd offsets."""
    rng = np.random.default_rng(11)
    out = []
    n = 70001
    base = np.repeat(rng.integers(0, 50, size=n // 7 + 1), 7)[:n].astype(np.float64)
    out.append(("runs", base.copy()))
    a = base.copy(); a[40000] = -0.0; a[100] = 0.0                # +0 first: min is +0.0
    out.append(("poszero_first", a))
    a = base.copy() + 1.0; a[5] = -0.0; a[60000] = 0.0            # -0 first: min is -0.0
    out.append(("negzero_first", a))
    a = -base.copy() - 1.0; a[33000] = 0.0; a[32767] = -0.0       # max is a zero, -0 first (chunk edge)
    out.append(("max_negzero", a))
    a = base.copy(); a[12345] = 2.0 ** -70                        # positive sliver: NOT fractional (split_n)
    out.append(("tiny_pos", a))
    a = base.copy(); a[12345] = -(2.0 ** -70)                     # negative sliver: fractional (sign extension)
    out.append(("tiny_neg", a))
    a = base.copy(); a[69999] = 0.5                               # fractional in the last chunk only
    out.append(("late_fraction", a))
    a = np.arange(131072, dtype=np.float64) % 300                 # run ends at every index incl. 250/251, 65535/65536
    out.append(("all_runs", a))
    a = np.full(131072, 7.0); a[65535] = 8.0; a[250] = 9.0
    out.append(("threshold_runs", a))
    a = (rng.integers(0, 3, size=32769) * 1e9).astype(np.float64)  # I32 bitdepth, one sample past a chunk
    out.append(("chunk_plus_one", a))
    return out


@pytest.mark.parametrize("comp", [O.RLE, O.CONSTANT, O.POLYNOMIAL, O.NOOP])
def test_stats_awkward_values(ctx, comp):
    cs = awkward_frames()
    arrays = [a for _, a in cs]
    # an odd sample offset makes every frame 8-byte (not 16-byte) aligned: scalar load path
    for shift in (0, 1):
        arrs = ([np.zeros(1)] if shift else []) + arrays
        got = run_batch(ctx, arrs, comp, bounded=False)[shift:]
        for (name, a), (o, b) in zip(cs, got):
            assert b == O.compress(comp, a), f"{O.NAMES[comp]} {name} shift={shift}"


def test_waves_and_engines_match_single_wave(ctx):
    """The same call cut into many small waves (several in flight) gives the same records and bytes."""
    import atsc_b200
    kinds = ["periodic", "gauge", "util", "saw", "steps", "constant", "noisy"]
    series = [gen.make(k, 300_000 + 1000 * i, 50 + i) for i, k in enumerate(kinds * 2)]
    flat = np.concatenate(series)
    offs, lens, o0 = [], [], 0
    for s in series:
        for c in atsc_b200.chunk_sizes(len(s)):
            offs.append(o0); lens.append(c); o0 += c
    ref_out, ref_pay = ctx.compress_frames(flat, offs, lens, atsc_b200.AUTO, 0.05, 0, True)
    for engines, wave_mi in (("1", "1"), ("3", "1"), ("4", "2")):
        os.environ["ATSC_ENGINES"], os.environ["ATSC_WAVE_MI"] = engines, wave_mi
        try:
            c2 = atsc_b200.Context()
        finally:
            os.environ.pop("ATSC_ENGINES"); os.environ.pop("ATSC_WAVE_MI")
        out, pay = c2.compress_frames(flat, offs, lens, atsc_b200.AUTO, 0.05, 0, True)
        assert len(pay) == len(ref_pay)
        for i in range(len(lens)):
            a, b = out[i], ref_out[i]
            assert (a.compressor, a.payload_len, a.payload_off, a.iterations) == \
                   (b.compressor, b.payload_len, b.payload_off, b.iterations), f"frame {i} engines={engines}"
        assert np.array_equal(pay, ref_pay)
        # and back: decode waves rotate over the engines too
        frames = [(out[i].compressor, int(lens[i]), int(out[i].payload_off), int(out[i].payload_len), int(offs[i]))
                  for i in range(len(lens))]
        dec = c2.decompress_frames(frames, pay)
        assert np.array_equal(dec, ctx.decompress_frames(frames, ref_pay))
        c2.close()


def test_payload_overflow_reemit(ctx):
    """-e 0 on incompressible data stores every sample (8 B/sample): far beyond the 1 B/sample the
    payload buffer is sized for, so k_emit refuses and the host re-launches it into a grown buffer."""
    arrays = [gen.make("noisy", n, 3 + i) for i, n in enumerate([131072, 65536, 131072, 4096])]
    got = run_batch(ctx, arrays, O.AUTO, max_error=0.0)
    for a, (o, b) in zip(arrays, got):
        wc, wb, _, _ = O.compress_best(a, np.float32(0.0), 0)
        assert o.compressor == wc and b == wb
        assert len(b) > 8 * len(a)


@pytest.mark.parametrize("e", [0.01, 0.05])
def test_small_frame_fft_every_length_class(ctx, e):
    """k_fft_small: unpadded lengths (< 128, primes included) and every padded length up to 1152."""
    sizes = [1, 2, 7, 31, 97, 101, 127, 128, 143, 144, 161, 191, 215, 242, 255, 287, 323, 383, 431, 485, 511, 512,
             575, 576, 647, 728, 767, 863, 971, 1023, 1024]
    cs = [(k, n, 200 + n) for n in sizes for k in ("periodic", "gauge", "util")]
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.FFT, max_error=e)
    bad = 0
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want, werr, wit = O.compress_bounded(O.FFT, a, float(np.float32(e)))
        ge, gmx, gmn = parse_fft(b)
        we, wmx, wmn = parse_fft(want)
        assert (gmx, gmn) == (wmx, wmn)
        if len(ge) != len(we) or o.iterations != wit:
            assert o.near_tie & 9, f"{k} n={n}: k {len(ge)} vs {len(we)}, iters {o.iterations} vs {wit}, no tie flag"
            bad += 1
            continue
        if o.near_tie & 8:
            continue
        gd, wd = O.decompress(O.FFT, n, b), O.decompress(O.FFT, n, want)
        assert np.abs(gd - wd).max() <= fft_tol(a, n), f"{k} n={n}"
    assert bad <= 3


def test_decoder_polynomial_every_step(ctx):
    """poly_expand against the oracle for every step the refinement loop can store (1..133),
    regular and irregular last segments, tiny key counts."""
    frames, blobs, want = [], [], []
    po = oo = 0
    rng = np.random.default_rng(5)
    for n in (3, 4, 5, 9, 100, 101, 399, 400, 401, 1000, 4099, 13300, 70001):
        a = np.round(rng.normal(100, 20, size=n), 3)
        for e in (0.5, 0.05, 0.01, 0.002, 0.0):
            b, _, _ = O.compress_bounded(O.POLYNOMIAL, a, float(np.float32(e)))
            frames.append((O.POLYNOMIAL, n, po, len(b), oo))
            blobs.append(b)
            want.append(O.decompress(O.POLYNOMIAL, n, b))
            po += len(b)
            oo += n
    out = ctx.decompress_frames(frames, np.frombuffer(b"".join(blobs), dtype=np.uint8))
    oo = 0
    for (c, n, _, _, _), w in zip(frames, want):
        assert np.array_equal(out[oo:oo + n], w), f"poly decode n={n}: {np.sum(out[oo:oo + n] != w)} differ"
        oo += n
