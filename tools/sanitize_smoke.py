"""Small end-to-end pass over every kernel family for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen, atsc_b200
os.environ.setdefault("ATSC_WAVE_MI", "1")
ctx = atsc_b200.Context([0])
kinds = ["periodic", "gauge", "util", "saw", "steps", "constant", "noisy"]
series = [gen.make(k, 150_000 + 777 * i, 60 + i) for i, k in enumerate(kinds)]
for comp, err, speed in ((atsc_b200.AUTO, 5, 0), (atsc_b200.AUTO, 0, 0), (atsc_b200.AUTO, 5, 6), (atsc_b200.FFT, 3, 0),
                         (atsc_b200.POLYNOMIAL, 3, 0), (atsc_b200.RLE, 0, 0), (atsc_b200.NOOP, 0, 0)):
    bros = ctx.compress_data(series, compressor=comp, error=err, speed=speed)
    dec = ctx.decompress_data(bros)
    assert all(len(a) == len(b) for a, b in zip(series, dec))
    print("ok", atsc_b200.COMPRESSOR_NAMES[comp], err, speed, sum(len(b) for b in bros))
small = [gen.make(k, n, 5) for k in ("periodic", "gauge") for n in (1, 7, 127, 393, 512, 1024, 2048, 4096)]
bros = ctx.compress_data(small, compressor=atsc_b200.IDW, error=5)
ctx.decompress_data(bros)
print("ok idw")
ctx.close()
