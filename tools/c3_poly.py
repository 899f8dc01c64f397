"""C3 of the bench (Polynomial on 3072 x 65536-sample series) alone, for A/B runs of the Polynomial kernels:
    ATSC_POLY_ITEMS_MAXF=1000000 python tools/c3_poly.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import atsc_b200, bench, gen
S3 = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
h3 = np.empty((S3, 65536))
for s in range(S3):
    h3[s] = (gen.gauge_walk, gen.utilisation, gen.sawtooth)[s % 3](65536, 1000 + s)
c3 = torch.from_numpy(h3).cuda()
ctx = atsc_b200.Context([0])
offs, lens = bench.frame_table(S3, 65536)
pbuf = np.empty(512 << 20, dtype=np.uint8)
for r in range(4):
    ctx.kernel_ms(reset=True)
    out, pay = ctx.compress_frames(None, offs, lens, atsc_b200.POLYNOMIAL, 0.05, 0, True, samples_ptr=c3.data_ptr(), payload_out=pbuf)
    k = ctx.kernel_ms(reset=True)
    print(r, "call ms", round(ctx.last_call_ms, 3), "bytes", len(pay), {a: round(b, 3) for a, b in k.items() if b and b > 0.01}, flush=True)
ctx.close()
