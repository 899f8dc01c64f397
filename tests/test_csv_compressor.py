"""The `csv-compressor` tool (reference: csv-compressor/src/{main,csv,metric}.rs).
Host-only behaviour (index, WavBrro, error paths) runs on CPU; the compress / uncompress round trip
goes through the GPU library and is checked against the oracle's BRO stream."""
import os
import subprocess

import numpy as np
import pytest

import atsc_b200
import oracle_lib as O
from test_vsri_cpu import RefVsri

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "atsc_b200", "csv-compressor")


def run(*args, ok=True):
    r = subprocess.run([BIN, *map(str, args)], capture_output=True, text=True, timeout=300)
    if ok:
        assert r.returncode == 0, r.stderr
    return r


def write_csv(path, ts_ms, values, header="timestamp,value"):
    with open(path, "w") as f:
        f.write(header + "\n")
        for t, v in zip(ts_ms, values):
            f.write(f"{t},{v!r}\n")


DAY0 = 1730419200000  # 2024-11-01T00:00:00Z in ms


def test_no_compression_writes_index_and_wavbrro(tmp_path):
    ts = [DAY0 + 1000 * s for s in (0, 15, 30, 45, 700, 715, 716)]
    vals = [1.01, 1.22, 5.0, 1e-5, 1e20, -0.5, 3.25]
    p = tmp_path / "m.csv"
    write_csv(p, ts, vals)
    run("--no-compression", "--output-vsri", "--output-wavbrro", p)
    assert sorted(os.listdir(tmp_path)) == ["m.csv", "m.vsri", "m.wavbro"]
    r = RefVsri()
    for t in ts:
        assert r.update((t // 1000) % 86400)
    assert open(tmp_path / "m.vsri").read() == r.text()
    assert open(tmp_path / "m.wavbro", "rb").read() == atsc_b200.wbro_encode(np.array(vals))


def test_output_base_and_column_order(tmp_path):
    p = tmp_path / "in.csv"
    with open(p, "w") as f:
        f.write("value,extra,timestamp\n2.5,x,%d\n3.5,y,%d\n" % (DAY0 + 60000, DAY0 + 120000))
    out = tmp_path / "sub.dir" / "res.anything"
    os.makedirs(out.parent)
    run("--no-compression", "--output-vsri", "-o", out, p)
    assert open(out.with_suffix(".vsri")).read() == "60\n120\n60,0,60,2\n"


@pytest.mark.parametrize("body,msg", [
    ("timestamp,value\n5000,1\n1000,2\n", "updating for point failed, sample: Sample { timestamp: 1000, value: 2.0 }"),
    ("time,value\n5000,1\n", "missing field `timestamp`"),
    ("timestamp,value\n5000,abc\n", "invalid value"),
    ("timestamp,value\n1.5,2\n", "invalid timestamp"),
])
def test_bad_input_is_a_panic_status(tmp_path, body, msg):
    p = tmp_path / "bad.csv"
    p.write_text(body)
    r = run("--no-compression", p, ok=False)
    assert r.returncode == 101 and msg in r.stderr
    assert os.listdir(tmp_path) == ["bad.csv"]


def test_input_must_be_a_file_and_no_cpu_path(tmp_path):
    assert run(tmp_path, ok=False).returncode == 101
    assert run(tmp_path / "missing.csv", ok=False).returncode == 101
    import torch
    if not torch.cuda.is_available():
        p = tmp_path / "m.csv"
        write_csv(p, [DAY0], [1.0])
        r = run(p, ok=False)
        assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("comp,oc,err", [("auto", O.AUTO, 5), ("polynomial", O.POLYNOMIAL, 0), ("noop", O.NOOP, 5),
                                         ("constant", O.CONSTANT, 5), ("idw", O.IDW, 3), ("fft", O.FFT, 5)])
def test_round_trip_matches_oracle(tmp_path, comp, oc, err):
    rng = np.random.default_rng(7)
    n = 3000                                        # strictly increasing seconds within one day: the reference's
    secs = np.cumsum(rng.choice([15, 15, 15, 15, 30, 45], size=n))  # index drops repeated timestamps (lib.rs:375-388)
    assert secs[-1] < 86400
    vals = np.round(50.0 + 10.0 * np.sin(np.arange(n) / 40.0) + rng.normal(0, 0.2, n), 2)
    if comp == "constant":
        vals[:] = 42.0
    p = tmp_path / "metric.csv"
    write_csv(p, [DAY0 + 1000 * int(s) + 7 for s in secs], vals.tolist())
    run("--compressor", comp, "-e", err, "--output-vsri", p)
    bro = open(tmp_path / "metric.bro", "rb").read()
    want, comps = O.compress_stream(vals, compressor=oc, error_pct=err)
    if O.FFT not in comps:
        assert bro == want
    r = RefVsri()
    for s in secs:
        assert r.update(int(s))
    assert open(tmp_path / "metric.vsri").read() == r.text()

    os.remove(p)
    run("-u", tmp_path / "metric.bro")
    got = atsc_b200.wbro_decode(open(tmp_path / "metric.wbro", "rb").read())
    dec = O.decompress_stream(bro)
    if O.FFT in comps:                              # f32 inverse transform: noise-bounded
        assert len(got) == len(dec) and np.abs(got - dec).max() <= 1e-3 * np.abs(dec).max()
    else:
        assert np.array_equal(got, dec)
    lines = open(tmp_path / "metric.csv").read().splitlines()
    assert lines[0] == "timestamp,value" and len(lines) == n + 1
    for i in (0, 1, 2, n // 2, n - 1):
        t, v = lines[i + 1].split(",")
        assert int(t) == r.get_time(i) and float(v) == got[i]
    if err == 0:
        assert np.array_equal(got, vals)


@pytest.mark.gpu
def test_uncompress_needs_the_index(tmp_path):
    p = tmp_path / "m.csv"
    write_csv(p, [DAY0 + 1000 * i for i in range(100)], [float(i % 7) for i in range(100)])
    run(p)                                          # no --output-vsri: only the .bro is written
    assert sorted(os.listdir(tmp_path)) == ["m.bro", "m.csv"]
    r = run("-u", tmp_path / "m.bro", ok=False)
    assert r.returncode == 101 and "failed to read vsri" in r.stderr
