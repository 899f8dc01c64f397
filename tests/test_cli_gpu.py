"""End-to-end through the `atsc` binary (reference: atsc/tests/e2e.rs, integration_test.rs)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")
BIN = os.path.join(ROOT, "atsc_b200", "atsc")


def run(*args):
    r = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


@pytest.fixture()
def wbro(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    p = tmp_path / "go_gc_heap_goal_bytes.wbro"
    shutil.copy(os.path.join(G, "go_gc_heap_goal_bytes.wbro"), p)
    return p


@pytest.mark.parametrize("comp,oc", [("noop", O.NOOP), ("rle", O.RLE), ("constant", O.CONSTANT),
                                     ("polynomial", O.POLYNOMIAL), ("idw", O.IDW), ("auto", O.AUTO)])
def test_cli_bro_matches_oracle(wbro, comp, oc):
    import atsc_b200
    x = atsc_b200.wbro_decode(open(wbro, "rb").read())
    run("--compressor", comp, "-e", "5", str(wbro))
    bro = open(wbro.with_suffix(".bro"), "rb").read()
    want, comps = O.compress_stream(x, compressor=oc, error_pct=5)
    if oc != O.AUTO or O.FFT not in comps:
        assert bro == want
    os.remove(wbro)
    run("-u", str(wbro.with_suffix(".bro")))
    got = atsc_b200.wbro_decode(open(wbro, "rb").read())
    wdec = O.decompress_stream(want)
    assert len(got) == len(wdec)
    if bro == want:
        assert np.array_equal(got, wdec)


def test_cli_lossless_and_lossy_e2e(wbro):
    """e2e.rs: -e 0 round trips exactly (polynomial / auto); -e 5 keeps whole-file MAPE <= 5 %."""
    import atsc_b200
    x = atsc_b200.wbro_decode(open(wbro, "rb").read())
    for comp in ("polynomial", "auto", "rle"):
        run("--compressor", comp, "-e", "0", str(wbro))
        os.remove(wbro)
        run("-u", str(wbro.with_suffix(".bro")))
        assert np.array_equal(atsc_b200.wbro_decode(open(wbro, "rb").read()), x), comp
    for comp in ("fft", "polynomial", "idw", "auto"):
        run("--compressor", comp, "-e", "5", "-c", "3", str(wbro))
        os.remove(wbro)
        run("-u", str(wbro.with_suffix(".bro")))
        d = atsc_b200.wbro_decode(open(wbro, "rb").read())
        assert O.mape(x, d) <= 0.05, comp
        shutil.copy(os.path.join(G, "go_gc_heap_goal_bytes.wbro"), wbro)


def test_cli_csv_and_verbose(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    fx = np.load(os.path.join(G, "fixtures.npz"))
    vals = np.round(fx["csv_cpu_utilization"])
    p = tmp_path / "cpu.csv"
    p.write_text("time,value\n" + "".join(f"{1730419200 + 20 * i},{float(v)!r}\n" for i, v in enumerate(vals)))
    out = run("--csv", "--compressor", "noop", "--verbose", str(p))
    assert out.startswith("Input=[60.0, 65.0, 69.0")
    out = run("-u", "--verbose", str(p.with_suffix(".bro")))
    assert out.startswith("Output=[60.0, 65.0, 69.0")
    import atsc_b200
    assert np.array_equal(atsc_b200.wbro_decode(open(p.with_suffix(".wbro"), "rb").read()), vals)
    q = tmp_path / "vals.csv"
    q.write_text("".join(f"{float(v)!r}\n" for v in vals))
    run("--csv", "--no-header", "--compressor", "noop", str(q))
    assert open(q.with_suffix(".bro"), "rb").read() == open(p.with_suffix(".bro"), "rb").read()


def test_cli_directory_fleet(tmp_path):
    """main.rs:50-68 directory mode: every file of a directory, batched into one GPU call; each
    .bro equals what the single-file path (and the oracle) produces, and -u on the directory
    restores every series."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import atsc_b200
    import gen
    d = tmp_path / "fleet"
    d.mkdir()
    kinds = ["constant", "gauge", "util", "saw", "steps", "periodic"]
    series = {}
    for i, k in enumerate(kinds * 2):
        x = gen.make(k, 3000 + 517 * i, 40 + i)
        series[f"s{i:02d}"] = x
        (d / f"s{i:02d}.wbro").write_bytes(atsc_b200.wbro_encode(x))
    run("--compressor", "auto", "-e", "5", str(d))
    for name, x in series.items():
        bro = (d / f"{name}.bro").read_bytes()
        want, comps = O.compress_stream(x, compressor=O.AUTO, error_pct=5)
        if O.FFT not in comps:
            assert bro == want, name
        else:
            assert len(O.decompress_stream(bro)) == len(x)
        os.remove(d / f"{name}.wbro")
    run("-u", str(d))
    for name, x in series.items():
        got = atsc_b200.wbro_decode((d / f"{name}.wbro").read_bytes())
        assert len(got) == len(x)
        nz = x != 0
        assert O.mape(x[nz], got[nz]) <= 0.0505, name
