#!/usr/bin/env python
"""Extract golden vectors from the reference tree into small committed fixtures.

Run HERE (the build container), never on the GPU box:
    python tests/golden/make_golden.py
Reads /root/reference (read-only) and writes tests/golden/*.npz.

Sources (all data, no code):
  * atsc/demo/comparison-error-{1,3}-{heap,memory,csv-iowait}.html -- arrays printed by
    the reference binary itself (`--verbose`): inputData and the reference's own
    decompressed output for --compressor fft / idw / polynomial at -e 1 and -e 3
    (atsc/demo/run_demo.sh:9-22).
  * atsc/tests/wbros/*.wbro, atsc/tests/csv/*.csv -- the reference's test fixtures
    (decoded to plain f64 arrays).
"""
import os
import re
import struct
import sys

import numpy as np

REF = "/root/reference/atsc"
OUT = os.path.dirname(os.path.abspath(__file__))


def html_arrays(path):
    txt = open(path).read()
    out = {}
    for name in ("inputData", "fftData", "idwData", "polyData"):
        m = re.search(r"const %s = \[(.*?)\];" % name, txt, re.S)
        vals = [float(x) for x in m.group(1).split(",") if x.strip()]
        out[name] = np.array(vals, dtype=np.float64)
    return out


def read_wbro(path):
    """WBRO = 12-byte header + rkyv 0.7 archive (wavbrro/src/wavbrro.rs:36-45)."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"WBRO" and raw[8:12] == b"WBRO"
    b = raw[12:]
    root = len(b) - 16
    rel, n_chunks, sample_count, bitdepth = struct.unpack_from("<iIIB", b, root)
    arr = root + rel
    chunks = []
    for c in range(n_chunks):
        at = arr + 8 * c
        crel, clen = struct.unpack_from("<iI", b, at)
        chunks.append(np.frombuffer(b, dtype="<f8", count=clen, offset=at + crel))
    data = np.concatenate(chunks) if chunks else np.zeros(0)
    assert len(data) == sample_count, (len(data), sample_count)
    return data.astype(np.float64)


def read_csv_values(path, header=True, col="value"):
    rows = open(path).read().strip().splitlines()
    if header:
        names = [c.strip() for c in rows[0].split(",")]
        idx = names.index(col)
        rows = rows[1:]
    else:
        idx = 0
    return np.array([float(r.split(",")[idx]) for r in rows], dtype=np.float64)


def main():
    demo = {}
    for err in (1, 3):
        for name in ("heap", "memory", "csv-iowait"):
            arrs = html_arrays(f"{REF}/demo/comparison-error-{err}-{name}.html")
            for k, v in arrs.items():
                demo[f"e{err}_{name.replace('-', '_')}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "demo_html.npz"), **demo)

    fx = {}
    for n in ("go_gc_heap_goal_bytes", "memory_used", "uptime"):
        fx["wbro_" + n] = read_wbro(f"{REF}/tests/wbros/{n}.wbro")
    fx["csv_cpu_utilization"] = read_csv_values(f"{REF}/tests/csv/cpu_utilization.csv")
    fx["csv_iowait"] = read_csv_values(f"{REF}/tests/csv/iowait.csv")
    fx["csv_cpu_utilization_no_headers"] = read_csv_values(
        f"{REF}/tests/csv/cpu_utilization_no_headers_only_values.csv", header=False)
    np.savez_compressed(os.path.join(OUT, "fixtures.npz"), **fx)
    # one raw WBRO fixture (binary data file, 23 KB) for the host WBRO reader test
    import shutil
    shutil.copy(f"{REF}/tests/wbros/go_gc_heap_goal_bytes.wbro", os.path.join(OUT, "go_gc_heap_goal_bytes.wbro"))
    for k, v in {**demo, **fx}.items():
        print(k, v.shape, v[:3])


if __name__ == "__main__":
    sys.exit(main())
