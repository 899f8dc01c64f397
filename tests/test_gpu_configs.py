"""GPU parity at the shapes BASELINE.json's configs name: the bench measures
config 4 (auto on the mixed fleet); configs 2, 3 and 5 are checked here against the oracle."""
import numpy as np
import pytest

import gen
import oracle_lib as O
from test_gpu_parity import fft_tol, parse_fft

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


def frames_of(ctx, x, comp, e):
    import atsc_b200
    cs = atsc_b200.chunk_sizes(len(x))
    offs = np.concatenate([[0], np.cumsum(cs)[:-1]]).astype(np.uint64)
    out, pay = ctx.compress_frames(x, offs, cs, comp, e, 0, True)
    return cs, offs, out, pay


@pytest.mark.parametrize("epct", [1, 5, 10])
def test_config2_fft_only_1m_series(ctx, epct):
    """FFT compressor only: one 1M-sample sinusoid + noise series (sigma 0.5), -e 1/5/10."""
    x = gen.periodic(1_000_000, 42, sigma=0.5)
    e = float(np.float32(epct / 100.0))
    cs, offs, out, pay = frames_of(ctx, x, O.FFT, e)
    assert cs == [131072] * 7 + [65536, 16384, 512, 64]
    flagged = 0
    for i, (n, o0) in enumerate(zip(cs, offs)):
        a = x[int(o0):int(o0) + n]
        o = out[i]
        b = pay[o.payload_off:o.payload_off + o.payload_len].tobytes()
        want, werr, wit = O.compress_bounded(O.FFT, a, e)
        ge, gmx, gmn = parse_fft(b)
        we, wmx, wmn = parse_fft(want)
        assert (gmx, gmn) == (wmx, wmn)
        if len(ge) != len(we) or o.iterations != wit:
            assert o.near_tie & 9, f"frame {i} n={n}: k {len(ge)} vs {len(we)}, iters {o.iterations} vs {wit}, no tie flag"
            flagged += 1
            continue
        assert o.error == pytest.approx(werr, rel=2e-4, abs=1e-9)
        gd, wd = O.decompress(O.FFT, n, b), O.decompress(O.FFT, n, want)
        assert np.abs(gd - wd).max() <= fft_tol(a, n), f"frame {i} n={n}"
    assert flagged <= 2


def test_config3_polynomial_and_idw_64k_series(ctx):
    """Polynomial (Catmull-Rom) and IDW on 65536-sample monitoring series, -e 5: bytes are exact."""
    e = float(np.float32(0.05))
    series = [gen.gauge_walk(65536, 1000), gen.utilisation(65536, 1001), gen.sawtooth(65536, 1002)]
    flat = np.concatenate(series)
    offs = [0, 65536, 131072]
    for comp in (O.POLYNOMIAL, O.IDW):
        out, pay = ctx.compress_frames(flat, offs, [65536] * 3, comp, e, 0, True)
        for i, a in enumerate(series):
            o = out[i]
            b = pay[o.payload_off:o.payload_off + o.payload_len].tobytes()
            want, werr, wit = O.compress_bounded(comp, a, e)
            assert b == want or o.near_tie, f"{O.NAMES[comp]} series {i}: {len(b)} vs {len(want)} bytes"
            if b == want:
                assert o.iterations == wit
                # and the decoder reproduces the reference's expansion bit for bit
                got = ctx.decompress_frames([(comp, 65536, 0, len(b), 0)], np.frombuffer(b, dtype=np.uint8))
                assert np.array_equal(got, O.decompress(comp, 65536, want))


def test_config5_decompress_fleet_matches_oracle(ctx):
    """The BRO fleet auto -e 5 produces, expanded on the GPU and by the oracle."""
    kinds = ["constant", "periodic", "util", "gauge", "saw", "steps"]
    series = [gen.make(k, 300_000, 900 + i) for i, k in enumerate(kinds)]
    bros, ties = ctx.compress_data(series, compressor=O.AUTO, error=5, return_ties=True)
    dec = ctx.decompress_data(bros)
    for k, x, bro, d in zip(kinds, series, bros, dec):
        w = O.decompress_stream(bro)     # the oracle reads OUR stream (BRO compatibility)
        assert len(w) == len(d) == len(x)
        exact = np.array_equal(w, d)
        assert exact or np.abs(w - d).max() <= fft_tol(x, 131072), f"{k}: {np.abs(w - d).max()}"
