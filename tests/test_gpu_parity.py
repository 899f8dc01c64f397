"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the reference's golden vectors, and through round-trip properties.

Bars (north_star): bit-exact bytes for Constant / RLE / Noop / Polynomial / IDW frames and
all headers; same compressor + parameter counts for FFT except flagged near-ties; FFT values
within f32 FFT noise: |a - b| <= 1e-5 + 6 * 2^-24 * log2(L) * max|x| (SURVEY.md H1 states 4 for one
transform; two independent f32 transforms are compared here, and the achieved deviation is reported).
"""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    c = atsc_b200.Context()
    yield c
    c.close()


def run_batch(ctx, arrays, compressor, max_error=0.05, speed=0, bounded=True):
    lens = [len(a) for a in arrays]
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    flat = np.concatenate(arrays)
    out, payload = ctx.compress_frames(flat, offs, lens, compressor, max_error, speed, bounded)
    res = []
    for i in range(len(arrays)):
        o = out[i]
        res.append((o, payload[o.payload_off:o.payload_off + o.payload_len].tobytes()))
    return res


def fft_tol(x, n, factor=6):
    # Two f32 transforms are compared (the oracle's recursive mixed radix stands in for rustfft, whose
    # butterfly order is build dependent): each carries up to ~4 eps log2(L) max|x| of rounding noise in
    # the decoded samples (DC dominated frames reach it), and the noises are independent.
    # SURVEY H1 states the single-transform figure (factor 4); every comparison records what it actually
    # reached against that (FFT_DEV, reported by test_zz_fft_deviation_report).  Measured over the 727
    # comparisons of this suite on B200: worst 1.25 x the factor-4 bound (a 127-sample gauge frame at
    # |x| = 5e7), 9 comparisons above it -- so the tolerance is 6, not round 1's 8.
    L = O.next_size(n) if n >= 128 else max(n, 2)
    return 1e-5 + factor * 2.0 ** -24 * np.log2(L) * float(np.abs(x).max())


FFT_DEV = {"n": 0, "max_ratio_vs_4x": 0.0, "worst": "", "above_4x": 0}
FFT_SKIPPED = []


def fft_close(gd, wd, x, n, what):
    """max |gpu - oracle| of decoded FFT values against the tolerance; records the achieved deviation."""
    dev = float(np.abs(np.asarray(gd) - np.asarray(wd)).max()) if len(gd) else 0.0
    ratio = dev / fft_tol(x, n, factor=4)
    FFT_DEV["n"] += 1
    FFT_DEV["above_4x"] += int(ratio > 1.0)
    if ratio > FFT_DEV["max_ratio_vs_4x"]:
        FFT_DEV["max_ratio_vs_4x"], FFT_DEV["worst"] = ratio, f"{what}: |d| = {dev:.3e}"
    assert dev <= fft_tol(x, n), f"{what}: decoded diff {dev} > {fft_tol(x, n)}"
    return dev


def parse_fft(payload):
    assert payload[0] == 15
    p = 1
    def varint(p):
        b = payload[p]
        if b < 251:
            return b, p + 1
        if b == 251:
            return int.from_bytes(payload[p + 1:p + 3], "little"), p + 3
        raise AssertionError("bad varint")
    c, p = varint(p)
    ents = []
    for _ in range(c):
        pos, p = varint(p)
        re, im = np.frombuffer(payload[p:p + 8], dtype="<f4")
        ents.append((pos, float(re), float(im)))
        p += 8
    mx, mn = np.frombuffer(payload[p:p + 8], dtype="<f4")
    assert p + 8 == len(payload)
    return ents, float(mx), float(mn)


SIZES = [1, 2, 3, 5, 12, 17, 64, 127, 128, 200, 393, 512, 2048, 4096, 16384, 65536, 131072]
KINDS = ["periodic", "gauge", "util", "saw", "noisy", "steps", "constant"]


def cases(sizes=SIZES, kinds=KINDS):
    out = []
    for n in sizes:
        for k in kinds:
            out.append((k, n, 1000 + n % 97 + len(k)))
    return out


# ------------------------------------------------------------------ reference unit-test vectors
def test_reference_golden_bytes(ctx):
    A = __import__("atsc_b200")
    V12 = [1.0, 0.0, 1.0, 1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
    V17 = [1.0, 1.0, 1.0, 1.0, 2.0, 3.0, 5.0, 1.0, 2.0, 7.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
    assert ctx.compress(A.CONSTANT, [1.0] * 5) == bytes([30, 3, 1])
    assert ctx.compress(A.CONSTANT, [1.23456] * 5) == bytes([30, 0, 56, 50, 143, 252, 193, 192, 243, 63])
    assert ctx.compress(A.NOOP, [1.0] * 5) == bytes([250, 5, 2, 2, 2, 2, 2])
    assert ctx.compress(A.RLE, [1.0] * 512) == bytes([60, 3, 1, 1, 1, 0])
    assert ctx.compress(A.RLE, [1.0, 2.0, 2.0, 3.0, 3.0, 3.0, 4.0, 4.0, 4.0, 4.0, 5.0, 5.0, 5.0, 5.0, 5.0]) == \
        bytes([60, 3, 5, 1, 1, 0, 2, 1, 1, 3, 1, 3, 4, 1, 6, 5, 1, 10])
    assert ctx.compress(A.RLE, [1.23456] * 5) == bytes([60, 0, 1, 56, 50, 143, 252, 193, 192, 243, 63, 1, 0])
    assert ctx.compress(A.POLYNOMIAL, V12) == bytes(
        [0, 3, 4, 1, 2, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 20, 64, 4])
    assert ctx.compress(A.POLYNOMIAL, V12[:-1] + [500.0]) == bytes(
        [0, 2, 4, 2, 4, 6, 251, 232, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 127, 64, 4])
    assert ctx.compress(A.IDW, V12) == bytes(
        [1, 3, 4, 1, 2, 3, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 20, 64, 4])
    assert ctx.compress(A.POLYNOMIAL, [1.0] * 4) == bytes(
        [0, 3, 0, 0, 0, 0, 0, 0, 0, 240, 63, 0, 0, 0, 0, 0, 0, 240, 63, 1])
    out = ctx.decompress(A.POLYNOMIAL, 17, ctx.compress(A.POLYNOMIAL, V17))
    assert list(out) == [1.0, 1.4, 1.8, 2.2, 2.6, 3.0, 2.824, 2.392, 1.848, 1.336, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 5.0]
    out = ctx.decompress(A.IDW, 17, ctx.compress(A.IDW, V17))
    assert list(out) == [1.0, 1.13167, 1.62573, 2.32782, 2.83429, 3.0, 2.8335, 2.34163, 1.68979, 1.184, 1.0,
                         1.18933, 1.64488, 1.9634, 1.77047, 1.0, 5.0]
    F12 = [1.0, 1.0, 1.0, 1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0, 1.0, 5.0]
    out = ctx.decompress(A.FFT, 12, ctx.compress(A.FFT, F12))
    want = [1.0, 1.87201, 2.25, 1.0, 1.82735, 1.689, 1.82735, 1.0, 2.75, 1.189, 1.0, 3.311]
    assert np.abs(np.array(out) - np.array(want)).max() <= 1.001e-5
    b = ctx.compress(A.FFT, [1.0] * 1024)
    assert b[1] == 0 and list(ctx.decompress(A.FFT, 1024, b)) == [1.0] * 1024
    bros = ctx.compress_data([np.ones(1024)], compressor=A.CONSTANT)
    assert bros[0] == bytes([66, 82, 82, 79, 1, 0, 0, 0, 1, 1, 41, 251, 0, 4, 3, 3, 30, 3, 1])
    assert list(ctx.decompress_data(bros)[0]) == [1.0] * 1024


# ------------------------------------------------------------------ bit-exact compressors
@pytest.mark.parametrize("comp", [O.CONSTANT, O.NOOP, O.RLE, O.POLYNOMIAL, O.IDW])
def test_unbounded_bytes_exact(ctx, comp):
    cs = cases()
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, comp, bounded=False)
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want = O.compress(comp, a)
        assert b == want, f"{O.NAMES[comp]} {k} n={n}: {len(b)} vs {len(want)} bytes"


@pytest.mark.parametrize("e", [0.0, 0.01, 0.03, 0.05, 0.2])
def test_polynomial_bounded_bytes_exact(ctx, e):
    cs = cases()
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.POLYNOMIAL, max_error=e)
    bad = 0
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want, werr, wit = O.compress_bounded(O.POLYNOMIAL, a, float(np.float32(e)))
        if b != want:
            assert o.near_tie, f"poly e={e} {k} n={n}: bytes differ without a near-tie flag"
            bad += 1
        else:
            assert o.iterations == wit
            assert o.error == pytest.approx(werr, rel=1e-9, abs=1e-15, nan_ok=True)
    assert bad <= 1


@pytest.mark.parametrize("e", [0.01, 0.05])
def test_idw_bounded_bytes_exact(ctx, e):
    cs = cases(sizes=[3, 12, 17, 127, 128, 393, 512, 2048], kinds=["periodic", "gauge", "util", "saw", "steps"])
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.IDW, max_error=e)
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want, werr, wit = O.compress_bounded(O.IDW, a, float(np.float32(e)))
        assert b == want or o.near_tie, f"idw e={e} {k} n={n}"


# ------------------------------------------------------------------ FFT
def compare_fft(a, n, gpu_payload, o, want_payload, wit, what):
    ge, gmx, gmn = parse_fft(gpu_payload)
    we, wmx, wmn = parse_fft(want_payload)
    assert (gmx, gmn) == (wmx, wmn), what
    if len(ge) != len(we) or o.iterations != wit:
        assert o.near_tie & 9, f"{what}: k {len(ge)} vs {len(we)}, iters {o.iterations} vs {wit}, no tie flag"
        FFT_SKIPPED.append(f"{what}: near-tie flag {o.near_tie}, k {len(ge)} vs {len(we)}, iterations {o.iterations} vs {wit}")
        return False
    if o.near_tie & 8:
        FFT_SKIPPED.append(f"{what}: equal |z| at a top-k cut")
        return False  # equal |z| at the top-k cut: BinaryHeap pop order is unspecified
    gd = O.decompress(O.FFT, n, gpu_payload)
    wd = O.decompress(O.FFT, n, want_payload)
    fft_close(gd, wd, a, n, what)
    gp = {p for p, _, _ in ge}
    wp = {p for p, _, _ in we}
    if gp != wp:
        assert len(gp ^ wp) <= max(2, len(we) // 50) or (o.near_tie & 8), f"{what}: bin sets differ by {len(gp ^ wp)}"
    # positions are u16-wrapped on disk (fft.rs:242): a 131072-sample frame can hold two
    # entries with the same stored position, so compare per position in stored order
    wm, gm = {}, {}
    for p, r, i in we:
        wm.setdefault(p, []).append(complex(r, i))
    for p, r, i in ge:
        gm.setdefault(p, []).append(complex(r, i))
    scale = max(abs(complex(r, i)) for _, r, i in we) if we else 1.0
    for p, gl in gm.items():
        wl = wm.get(p)
        if wl is None or len(wl) != len(gl):
            FFT_SKIPPED.append(f"{what}: stored position {p} holds {len(gl)} entries here, {0 if wl is None else len(wl)} in the oracle")
            continue
        for gz, wz in zip(gl, wl):
            assert abs(gz - wz) <= 4e-6 * scale + 1e-6, f"{what}: bin {p}"
    return True


@pytest.mark.parametrize("e", [0.01, 0.05, 0.1])
def test_fft_bounded(ctx, e):
    cs = cases(kinds=["periodic", "gauge", "util", "saw", "noisy", "steps"])
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.FFT, max_error=e)
    same = 0
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want, werr, wit = O.compress_bounded(O.FFT, a, float(np.float32(e)))
        same += compare_fft(a, n, b, o, want, wit, f"fft e={e} {k} n={n}")
    assert same >= len(cs) - 3, "not compared value by value:\n" + "\n".join(x for x in FFT_SKIPPED if f"fft e={e} " in x)


def test_fft_unbounded_small_and_pow(ctx):
    cs = cases(sizes=[3, 12, 64, 127, 144, 1024, 2187, 4096], kinds=["periodic", "gauge", "saw"])
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.FFT, bounded=False)
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        want = O.compress(O.FFT, a)
        ge, _, _ = parse_fft(b)
        we, _, _ = parse_fft(want)
        assert len(ge) == len(we)
        assert [p for p, _, _ in ge] == [p for p, _, _ in we] or (o.near_tie & 8)


# ------------------------------------------------------------------ auto selection
@pytest.mark.parametrize("speed", [0, 3, 6])
@pytest.mark.parametrize("e", [0.0, 0.03, 0.05])
def test_auto_selection(ctx, e, speed):
    cs = cases()
    arrays = [gen.make(k, n, s) for k, n, s in cs]
    got = run_batch(ctx, arrays, O.AUTO, max_error=e, speed=speed)
    mism = 0
    for (k, n, s), a, (o, b) in zip(cs, arrays, got):
        wc, wb, werr, wsize = O.compress_best(a, np.float32(e), speed)
        what = f"auto e={e} c={speed} {k} n={n}: gpu {O.NAMES[o.compressor]}({len(b)}) oracle {O.NAMES[wc]}({len(wb)}) " \
               f"cand {list(o.cand_size)} {list(o.cand_error)} vs {wsize} {werr}"
        if o.compressor != wc:
            assert o.near_tie, what
            mism += 1
            continue
        if wc == O.FFT:
            ge, _, _ = parse_fft(b)
            we, _, _ = parse_fft(wb)
            if len(ge) != len(we):
                assert o.near_tie, what
                mism += 1
            else:
                gd, wd = O.decompress(O.FFT, n, b), O.decompress(O.FFT, n, wb)
                fft_close(gd, wd, a, n, what)
        else:
            assert b == wb, what
    assert mism <= 2


# ------------------------------------------------------------------ decompression
@pytest.mark.parametrize("comp", [O.CONSTANT, O.NOOP, O.RLE, O.POLYNOMIAL, O.IDW, O.FFT])
def test_decompress_oracle_payloads(ctx, comp):
    sizes = SIZES if comp != O.IDW else [3, 12, 17, 127, 128, 393, 512, 2048, 4096]
    cs = cases(sizes=sizes)
    frames, blobs, want = [], [], []
    po = oo = 0
    for k, n, s in cs:
        a = gen.make(k, n, s)
        if comp in (O.POLYNOMIAL, O.IDW, O.FFT):
            b, _, _ = O.compress_bounded(comp, a, float(np.float32(0.03)))
        else:
            b = O.compress(comp, a)
        frames.append((comp, n, po, len(b), oo))
        blobs.append(b)
        want.append(O.decompress(comp, n, b))
        po += len(b)
        oo += n
    out = ctx.decompress_frames(frames, np.frombuffer(b"".join(blobs), dtype=np.uint8))
    oo = 0
    for (k, n, s), w in zip(cs, want):
        g = out[oo:oo + n]
        if comp == O.FFT:
            a = gen.make(k, n, s)
            fft_close(g, w, a, n, f"fft decode {k} n={n}")
        else:
            assert np.array_equal(g, w), f"{O.NAMES[comp]} decode {k} n={n}: {np.sum(g != w)} differ"
        oo += n


# ------------------------------------------------------------------ streams + fixtures
def test_fixture_streams(ctx):
    fx = np.load(os.path.join(G, "fixtures.npz"))
    names = list(fx.keys())
    series = [fx[k] for k in names]
    for comp in (O.AUTO, O.POLYNOMIAL, O.RLE, O.NOOP, O.CONSTANT, O.FFT):
        for e in (0, 3, 5):
            bros, ties = ctx.compress_data(series, compressor=comp, error=e, return_ties=True)
            dec = ctx.decompress_data(bros)
            for nm, x, bro, tie, d in zip(names, series, bros, ties, dec):
                wbro, wcomps = O.compress_stream(x, compressor=comp, error_pct=e)
                wdec = O.decompress_stream(wbro)
                what = f"{nm} comp={O.NAMES[comp]} e={e}"
                assert len(d) == len(wdec), what
                if comp in (O.POLYNOMIAL, O.RLE, O.NOOP, O.CONSTANT):
                    assert bro == wbro or tie, what
                clean = x[np.isfinite(x)]
                if bro == wbro:
                    continue
                tol = fft_tol(clean, 2048)
                if not tie:
                    assert np.abs(d - wdec).max() <= tol, what
                # reference e2e criterion (atsc/tests/e2e.rs:244-264): whole-file MAPE <= e
                if comp in (O.AUTO, O.POLYNOMIAL, O.FFT) and e > 0 and np.all(clean != 0):
                    assert O.mape(clean, d) <= e / 100 + 1e-3 or comp == O.FFT, what


def test_demo_html_poly_idw(ctx):
    """The reference binary's own decompressed output (atsc/demo/*.html) through the GPU path."""
    demo = np.load(os.path.join(G, "demo_html.npz"))
    for err in (1, 3):
        for name in ("heap", "memory", "csv_iowait"):
            x = demo[f"e{err}_{name}_inputData"]
            for which, comp in (("polyData", O.POLYNOMIAL), ("idwData", O.IDW)):
                bro = ctx.compress_data([x], compressor=comp, error=err)
                got = ctx.decompress_data(bro)[0]
                want = demo[f"e{err}_{name}_{which}"]
                assert np.array_equal(got, want), f"{name} e={err} {which}: {np.sum(got != want)} differ"
            bro = ctx.compress_data([x], compressor=O.FFT, error=err)
            got = ctx.decompress_data(bro)[0]
            want = demo[f"e{err}_{name}_fftData"]
            fft_close(got, want, x[np.isfinite(x)], 2048, f"reference binary's fftData {name} e={err}")


# ------------------------------------------------------------------ full-size properties
def test_roundtrip_properties_full_size(ctx):
    """Size-independent properties at the bench's frame size: lossless round trips and the
    error bound on the decompressed stream."""
    n = 1_000_000
    series = [gen.make(k, n, 7 + i) for i, k in enumerate(["constant", "periodic", "util", "gauge", "saw", "steps"])]
    bros = ctx.compress_data(series, compressor=O.AUTO, error=0)
    dec = ctx.decompress_data(bros)
    for x, d in zip(series, dec):
        # -e 0: auto must be lossless up to the 5-decimal rounding of the codecs (e2e.rs:158-164)
        assert np.array_equal(np.round(x, 5), np.round(d, 5)) or np.abs(x - d).max() <= 1e-5
    bros = ctx.compress_data(series, compressor=O.AUTO, error=5)
    dec = ctx.decompress_data(bros)
    for x, d, b in zip(series, dec, bros):
        assert len(d) == n
        nz = x != 0
        assert O.mape(x[nz], d[nz]) <= 0.0505
        assert len(b) < n * 8
    # decompress is idempotent w.r.t. recompressing constants / rle
    bros = ctx.compress_data(series, compressor=O.RLE)
    for x, d in zip(series, ctx.decompress_data(bros)):
        assert np.array_equal(x, d)
    bros = ctx.compress_data([np.round(s) for s in series], compressor=O.NOOP)
    for x, d in zip(series, ctx.decompress_data(bros)):
        assert np.array_equal(np.round(x), d)


# ------------------------------------------------------------------ the bench fleet itself
@pytest.mark.parametrize("speed", [0, 6])
def test_bench_fleet_vs_oracle(ctx, speed):
    """bench.py's own generator (constant / periodic sigma=0.05 / utilisation, 1 M samples) through both arms:
    per frame the winner, the iteration count, and the bytes (FFT frames: entry count + values)."""
    import atsc_b200
    import bench
    series = [bench.make_series(c, 5000 + c + 3 * r) for r in range(2) for c in range(3)]
    flat = np.concatenate(series)
    offs, lens = bench.frame_table(len(series))
    out, pay = ctx.compress_frames(flat, offs, lens, atsc_b200.AUTO, 0.05, speed, True)
    ties = mism = 0
    for i in range(len(lens)):
        a = flat[int(offs[i]):int(offs[i]) + int(lens[i])]
        o = out[i]
        b = pay[o.payload_off:o.payload_off + o.payload_len].tobytes()
        wc, wb, werr, wsize = O.compress_best(a, np.float32(0.05), speed)
        what = f"bench fleet c={speed} frame {i} (n={lens[i]}): gpu {O.NAMES[o.compressor]}({len(b)}) oracle {O.NAMES[wc]}({len(wb)})"
        ties += bool(o.near_tie)
        if o.compressor != wc:
            assert o.near_tie, what
            mism += 1
        elif wc == O.FFT:
            ge, we = parse_fft(b)[0], parse_fft(wb)[0]
            if len(ge) != len(we):
                assert o.near_tie, what
                mism += 1
            else:
                fft_close(O.decompress(O.FFT, len(a), b), O.decompress(O.FFT, len(a), wb), a, len(a), what)
        else:
            assert b == wb, what
    assert mism <= 1 and ties <= 3, f"{mism} mismatches, {ties} near-tie flags in {len(lens)} frames"


def test_zz_fft_deviation_report():
    """Runs last in this file: the deviation every FFT comparison above actually reached, against SURVEY H1's
    single-transform tolerance (factor 4).  The assertion documents the headroom of the factor-6 tolerance."""
    import json
    rep = {k: (v if isinstance(v, str) else float(v)) for k, v in FFT_DEV.items()}
    rep["skipped"] = len(FFT_SKIPPED)
    print("\nFFT deviation report:", json.dumps(rep))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "fft_deviation.json"), "w") as f:
            json.dump(dict(rep, skipped_cases=FFT_SKIPPED), f, indent=1)
    assert FFT_DEV["n"] > 100
    assert FFT_DEV["max_ratio_vs_4x"] <= 1.5  # == the factor-6 tolerance every comparison asserted
