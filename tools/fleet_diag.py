"""Which frames of the bench fleet reach which kernel: per (class, frame length) the winners, how many frames
had their FFT candidate evaluated to the end (k_fft / k_fft_small), polynomial iterations, near-tie flags."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from collections import Counter, defaultdict
import atsc_b200, bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 36
speed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = atsc_b200.Context([0])
host = np.empty((S, bench.SERIES_LEN))
bench.make_fleet(S, 5000, host)
offs, lens = bench.frame_table(S)
out, pay = ctx.compress_frames(host.reshape(-1), offs, lens, atsc_b200.AUTO, 0.05, speed, True)
per = len(lens) // S
rows = defaultdict(lambda: dict(n=0, win=Counter(), fft_eval=0, it=Counter(), tie=0, fft_it=Counter(), bytes=0))
for i in range(len(lens)):
    cls = ("const", "periodic", "util")[(i // per) % 3]
    r = rows[(cls, int(lens[i]))]
    o = out[i]
    r["n"] += 1; r["win"][atsc_b200.COMPRESSOR_NAMES[o.compressor]] += 1
    r["fft_eval"] += 1 if o.cand_size[0] else 0
    r["it"][o.iterations] += 1
    r["tie"] += 1 if o.near_tie else 0
    r["bytes"] += o.payload_len
for k in sorted(rows):
    r = rows[k]
    print(k, r["n"], dict(r["win"]), "fft evaluated:", r["fft_eval"], "iters:", dict(r["it"]), "ties:", r["tie"], "bytes:", r["bytes"])
print(ctx.kernel_ms(reset=True))
