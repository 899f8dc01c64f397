"""The front ends of the big frames against each other and against the CPU oracle: k_front (ATSC_FRONT=1: stats +
first Polynomial step + FFT probe fold in one read), k_sfold (ATSC_FRONT=2, the default: stats + probe fold, with
k_probe), the separate passes (ATSC_FRONT=0), and the first-step work items (k_poly1s, k_poly1): same frame records,
same payload bytes.
Reference behaviour under test: frame/mod.rs:71-149, polynomial.rs:209-277, optimizer/utils.rs:39-89."""
import os

import numpy as np
import pytest

import gen
import oracle_lib as O
from test_gpu_parity import run_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    os.environ["ATSC_FRONT"] = "1"  # the fused front end is opt-in
    try:
        on = atsc_b200.Context()
    finally:
        os.environ.pop("ATSC_FRONT")
    os.environ["ATSC_FRONT"] = "0"
    try:
        off = atsc_b200.Context()
    finally:
        os.environ.pop("ATSC_FRONT")
    yield on, off
    on.close()
    off.close()


@pytest.fixture(scope="module")
def sfold_ctxs():
    """k_sfold (ATSC_FRONT=2: stats + FFT probe in one read, sfold.cuh) against the separate passes."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    made = []
    for v in ("2", "0"):
        os.environ["ATSC_FRONT"] = v
        try:
            made.append(atsc_b200.Context())
        finally:
            os.environ.pop("ATSC_FRONT")
    yield made
    for c in made:
        c.close()


def front_frames():
    """Frames of >= 16384 samples that exercise every branch of the streaming pass."""
    rng = np.random.default_rng(77)
    out = []
    for n in (16384, 32768, 65536, 131072):
        for k in ("periodic", "gauge", "util", "saw", "noisy", "steps", "constant"):
            out.append((f"{k}{n}", gen.make(k, n, 300 + n % 89 + len(k))))
    # even lengths that are not powers of two (no fold, ragged last tile) and one ragged fast geometry
    for n in (16386, 20000, 70000, 100002, 131070):
        out.append((f"util{n}", gen.make("util", n, n)))
        out.append((f"gauge{n}", gen.make("gauge", n, n + 1)))
    a = gen.make("periodic", 131072, 5)
    a[:9000] = a[0]                                   # constant tiles first: skipped work -> fallback
    out.append(("const_prefix", a))
    a = gen.make("util", 65536, 6)
    a[:4096] = a[0]
    out.append(("const_tile0", a))
    a = gen.make("periodic", 65536, 7) - 100.0         # sign changes: not tame
    out.append(("signs", a))
    a = gen.make("util", 131072, 8)
    a[70000] = 0.0                                     # one zero sample: not tame, MAPE inf
    out.append(("one_zero", a))
    a = gen.make("periodic", 131072, 9, )
    a[::5000] += 400.0                                 # spikes: Catmull-Rom overshoot -> parked samples
    out.append(("spikes", a))
    a = np.where(np.arange(65536) % 2000 < 1000, 10.0, 90.0) + rng.normal(0, 0.01, 65536)  # square wave: clamp acts
    out.append(("square", np.round(a, 3)))
    a = np.linspace(1.0, 1e6, 131072)                  # trend: the published range moves with every tile
    out.append(("ramp", np.round(a, 2)))
    a = np.full(131072, 7.0); a[65535] = 8.0; a[250] = 9.0; a[249] = 9.5
    out.append(("threshold_runs", a))
    a = np.arange(131072, dtype=np.float64) % 300
    out.append(("all_runs", a))
    a = gen.make("util", 32768, 10); a[5] = np.nan     # NaN sample (frame level API does not clean)
    out.append(("nan", a))
    a = gen.make("noisy", 65536, 11) * 1e12            # beyond the tame magnitude
    out.append(("huge", a))
    return out


def same_records(on, off, names, tag):
    for name, (a, pa), (b, pb) in zip(names, on, off):
        what = f"{tag} {name}"
        assert (a.compressor, a.payload_len, a.iterations) == (b.compressor, b.payload_len, b.iterations), what
        assert pa == pb, what
        if np.isfinite(b.error):
            assert abs(a.error - b.error) <= 1e-12 * max(1.0, abs(b.error)), what
        else:
            assert not np.isfinite(a.error) or a.error == b.error, what


@pytest.mark.parametrize("comp,err,speed", [
    (O.AUTO, 0.05, 0), (O.AUTO, 0.01, 0), (O.AUTO, 0.05, 6), (O.AUTO, 0.0, 0),
    (O.POLYNOMIAL, 0.05, 0), (O.POLYNOMIAL, 0.001, 0), (O.FFT, 0.05, 0), (O.RLE, 0.05, 0), (O.CONSTANT, 0.05, 0),
])
def test_front_matches_separate_passes(ctxs, comp, err, speed):
    on, off = ctxs
    cs = front_frames()
    if comp == O.FFT:
        cs = [c for c in cs if len(c[1]) in (16384, 32768, 65536, 131072)][:12]
    arrays = [a for _, a in cs]
    names = [n for n, _ in cs]
    r_on = run_batch(on, arrays, comp, max_error=err, speed=speed)
    r_off = run_batch(off, arrays, comp, max_error=err, speed=speed)
    same_records(r_on, r_off, names, f"{O.NAMES[comp]} e={err} c={speed}")
    k = on.kernel_ms(reset=True)
    assert k["front"] > 0.0, "k_front did not run"


def test_front_unbounded_and_idw(ctxs):
    on, off = ctxs
    cs = [c for c in front_frames() if len(c[1]) == 16384]
    arrays = [a for _, a in cs]
    names = [n for n, _ in cs]
    for comp, bounded in ((O.POLYNOMIAL, False), (O.IDW, True), (O.NOOP, False)):
        r_on = run_batch(on, arrays, comp, bounded=bounded)
        r_off = run_batch(off, arrays, comp, bounded=bounded)
        same_records(r_on, r_off, names, O.NAMES[comp])


def test_front_against_oracle(ctxs):
    """Auto at -e 5: winner, sizes and bytes against the oracle for the streamed frames."""
    on, _ = ctxs
    cs = [c for c in front_frames() if not c[0].startswith(("nan",))]
    arrays = [a for _, a in cs]
    got = run_batch(on, arrays, O.AUTO, max_error=0.05)
    bad = []
    for (name, a), (o, b) in zip(cs, got):
        wc, wb, _, _ = O.compress_best(a, np.float32(0.05), 0)
        if o.compressor != wc or (wc != O.FFT and b != wb):
            if not o.near_tie:
                bad.append(name)
    assert not bad, bad


def test_front_stats_bytes(ctxs):
    """RLE / Constant / Polynomial(unbounded) bytes of streamed frames equal the oracle's: the stats
    (min, max, bitdepth, run counts and their varint classes) are exact."""
    on, _ = ctxs
    cs = [c for c in front_frames() if c[0] in ("threshold_runs", "all_runs", "square", "steps131072", "saw65536",
                                                "gauge32768", "ramp", "signs", "huge")]
    arrays = [a for _, a in cs]
    for comp in (O.RLE, O.CONSTANT, O.POLYNOMIAL):
        got = run_batch(on, arrays, comp, bounded=False)
        for (name, a), (o, b) in zip(cs, got):
            assert b == O.compress(comp, a), f"{O.NAMES[comp]} {name}"


@pytest.mark.parametrize("err,speed", [(0.05, 0), (0.01, 0), (0.0, 0), (0.05, 6), (0.3, 0)])
def test_sfold_matches_separate_passes(sfold_ctxs, err, speed):
    on, off = sfold_ctxs
    cs = front_frames()
    arrays = [a for _, a in cs]
    names = [n for n, _ in cs]
    r_on = run_batch(on, arrays, O.AUTO, max_error=err, speed=speed)
    r_off = run_batch(off, arrays, O.AUTO, max_error=err, speed=speed)
    same_records(r_on, r_off, names, f"sfold Auto e={err} c={speed}")
    if speed == 0:
        assert on.kernel_ms(reset=True)["front"] > 0.0, "k_sfold did not run"


def test_sfold_against_oracle(sfold_ctxs):
    on, _ = sfold_ctxs
    cs = [c for c in front_frames() if not c[0].startswith(("nan",))]
    got = run_batch(on, [a for _, a in cs], O.AUTO, max_error=0.05)
    bad = []
    for (name, a), (o, b) in zip(cs, got):
        wc, wb, _, _ = O.compress_best(a, np.float32(0.05), 0)
        if (o.compressor != wc or (wc != O.FFT and b != wb)) and not o.near_tie:
            bad.append(name)
    assert not bad, bad


def test_sfold_mixed_batch_and_other_compressors(sfold_ctxs):
    """Frames k_sfold does not take (other compressors, odd lengths) share a wave with frames it does."""
    on, off = sfold_ctxs
    cs = front_frames()[:14]
    arrays = [a for _, a in cs] + [gen.make("util", 5000, 1), gen.make("periodic", 131071, 2)]
    names = [n for n, _ in cs] + ["util5000", "periodic131071"]
    for comp in (O.AUTO, O.RLE, O.POLYNOMIAL, O.FFT):
        r_on = run_batch(on, arrays, comp, max_error=0.05)
        r_off = run_batch(off, arrays, comp, max_error=0.05)
        same_records(r_on, r_off, names, f"sfold {O.NAMES[comp]}")


@pytest.fixture(scope="module")
def poly_item_ctxs():
    """First Polynomial step of frames >= 65536 samples in balanced work items (k_poly1s, poly.cuh) on / off."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    made = []
    for v in ("1", "0"):
        os.environ["ATSC_POLY_ITEMS"] = v
        os.environ["ATSC_FRONT"] = "0"
        try:
            made.append(atsc_b200.Context())
        finally:
            os.environ.pop("ATSC_POLY_ITEMS")
            os.environ.pop("ATSC_FRONT")
    yield made
    for c in made:
        c.close()


@pytest.mark.parametrize("comp,err", [(O.AUTO, 0.05), (O.AUTO, 0.002), (O.POLYNOMIAL, 0.05), (O.POLYNOMIAL, 0.001),
                                      (O.POLYNOMIAL, 0.0), (O.IDW, 0.05)])
def test_poly_items_match_whole_frames(poly_item_ctxs, comp, err):
    on, off = poly_item_ctxs
    cs = [c for c in front_frames() if len(c[1]) >= 65536 or c[0] in ("util20000", "gauge16386")]
    if comp == O.IDW:
        cs = cs[:2]
    arrays = [a for _, a in cs]
    names = [n for n, _ in cs]
    r_on = run_batch(on, arrays, comp, max_error=err)
    r_off = run_batch(off, arrays, comp, max_error=err)
    same_records(r_on, r_off, names, f"poly items {O.NAMES[comp]} e={err}")
    if comp == O.POLYNOMIAL:
        for (name, a), (o, b) in zip(cs, r_on):
            if not o.near_tie and not name.startswith("nan"):
                assert b == O.compress_bounded(comp, a, np.float32(err))[0], name


def test_poly1_static_matches_items():
    """k_poly1s (k_plan's compacted self-contained item list, static schedule, raw keys one item ahead by cp.async,
    samples one trip ahead in registers, one partial sum per warp, the left-over samples in poly_frame) against the
    queue-driven k_poly1
    (ATSC_POLY1_STATIC=0): same per-sample arithmetic, another summation tree -- same records and payload bytes,
    errors equal to 1e-12 relative; non-tame frames (zeros, sign changes, NaN), constant frames between the big
    ones (never listed by k_plan) and lengths that move the item boundaries included; vs the oracle too."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    made = []
    for v in ("1", "0"):
        os.environ["ATSC_POLY1_STATIC"] = v
        try:
            made.append(atsc_b200.Context())
        finally:
            os.environ.pop("ATSC_POLY1_STATIC")
    try:
        rng = np.random.default_rng(5)
        cs = [c for c in front_frames() if len(c[1]) >= 65536]
        for n in (65536, 65537, 70001, 98304, 99999, 131071, 131072):
            cs.append((f"util{n}", gen.make("util", n, n)))
            cs.append((f"const{n}", np.full(n, 3.25)))
            cs.append((f"sign{n}", np.sin(np.arange(n) / 900.0) * 3.0 + rng.normal(0, 0.01, n)))  # not tame
        # tame frames (one sign, away from zero) whose spline overshoots ACROSS zero between spikes at the key
        # positions: k_poly1s rounds them with the frame's sign (poly.cuh: p1_trip), the clamp must hide it
        for n, sgn in ((131072, 1.0), (100000, -1.0), (65536, 1.0)):
            a = np.full(n, 0.001)
            a[::300] = 10.0
            a[150::300] = 0.004
            a += rng.uniform(0, 1e-6, n)
            cs.append((f"overshoot{n}_{int(sgn)}", sgn * a))
        # enough items that every CTA of k_poly1s walks several (descriptor / key double buffering, ring hand-over)
        for i in range(12):
            cs.append((f"periodic131072_{i}", gen.make("periodic", 131072, 100 + i)))
        arrays = [a for _, a in cs]
        names = [n for n, _ in cs]
        for comp, err in ((O.AUTO, 0.05), (O.POLYNOMIAL, 0.05), (O.POLYNOMIAL, 0.002), (O.POLYNOMIAL, 0.0)):
            r_on = run_batch(made[0], arrays, comp, max_error=err)
            r_off = run_batch(made[1], arrays, comp, max_error=err)
            same_records(r_on, r_off, names, f"poly1 static {O.NAMES[comp]} e={err}")
            if comp == O.POLYNOMIAL:
                for (name, a), (o, b) in zip(cs, r_on):
                    if not o.near_tie and not name.startswith("nan"):
                        assert b == O.compress_bounded(comp, a, np.float32(err))[0], name
    finally:
        for c in made:
            c.close()


def test_probe_kernel_matches_in_kernel_probe():
    """k_probe (small probe tails, decided frames skip k_fft_fwd's probe) against ATSC_PROBE_KERNEL=0."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    made = []
    for v in ("1", "0"):
        os.environ["ATSC_PROBE_KERNEL"] = v
        try:
            made.append(atsc_b200.Context())
        finally:
            os.environ.pop("ATSC_PROBE_KERNEL")
    try:
        cs = front_frames()
        # sparse spectra: an exact DC step and a two-level square wave leave most probed bins at zero
        a = np.zeros(131072); a[:65536] = 5.0; a += 10.0
        cs.append(("dc_step", a))
        cs.append(("lin_ramp_int", np.arange(65536, dtype=np.float64) + 1.0))
        arrays = [x for _, x in cs]
        names = [n for n, _ in cs]
        for err in (0.05, 0.0005):
            r1 = run_batch(made[0], arrays, O.AUTO, max_error=err)
            r0 = run_batch(made[1], arrays, O.AUTO, max_error=err)
            same_records(r1, r0, names, f"probe kernel e={err}")
    finally:
        for c in made:
            c.close()
