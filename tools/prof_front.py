"""Small driver for profiling k_front: S series of the bench fleet, device resident, a few compress calls.
    python tools/prof_front.py [series] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import atsc_b200, bench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 72
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = atsc_b200.Context([0])
host = np.empty((S, bench.SERIES_LEN))
bench.make_fleet(S, 5000, host)
dev = torch.from_numpy(host.reshape(-1)).cuda()
offs, lens = bench.frame_table(S)
pbuf = np.empty(64 << 20, dtype=np.uint8)
for r in range(reps):
    ctx.kernel_ms(reset=True)
    out, pay = ctx.compress_frames(None, offs, lens, atsc_b200.AUTO, 0.05, 0, True, samples_ptr=dev.data_ptr(), payload_out=pbuf)
    k = ctx.kernel_ms(reset=True)
    print(r, "call ms", round(ctx.last_call_ms, 3), {a: round(b, 3) for a, b in k.items() if b}, flush=True)
ctx.close()
