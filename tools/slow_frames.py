"""Diagnostic: which frames of a bench fleet keep the FFT candidate alive (not pruned)?"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, atsc_b200, bench
S = 288; r = int(sys.argv[1]) if len(sys.argv) > 1 else 1
host = np.empty((S, bench.SERIES_LEN)); bench.make_fleet(S, 5000 + r * S, host)
offs, lens = bench.frame_table(S)
ctx = atsc_b200.Context([0])
out, pay = ctx.compress_frames(host.reshape(-1), offs, lens, atsc_b200.AUTO, 0.05, 0, True)
for i in range(len(lens)):
    o = out[i]
    if lens[i] > 1024 and o.cand_size[2] > 0:
        print("series", i // 11, "class", (i // 11) % 3, "frame", i % 11, "len", lens[i], "winner", atsc_b200.COMPRESSOR_NAMES[o.compressor],
              "iters", o.iterations, "cand", list(o.cand_size), [round(e, 4) for e in o.cand_error])
