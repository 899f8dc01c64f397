"""Driver for profiling k_fft: `python tools/prof_fft.py noisy|c2 [series]` -- auto on an all-noise fleet, or the FFT
compressor alone on the C2 class (sinusoid + noise, sigma 0.5)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import atsc_b200, bench, gen

kind = sys.argv[1] if len(sys.argv) > 1 else "noisy"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = atsc_b200.Context([0])
dev = torch.device("cuda", 0)
if kind == "noisy":
    fleet = bench.make_fleet_device(S, 9000, dev, kinds=(3,)); comp = atsc_b200.AUTO
else:
    fleet = torch.from_numpy(np.stack([gen.periodic(bench.SERIES_LEN, 42 + s, sigma=0.5) for s in range(S)])).to(dev); comp = atsc_b200.FFT
offs, lens = bench.frame_table(S)
pbuf = np.empty(512 << 20, dtype=np.uint8)
for r in range(3):
    ctx.kernel_ms(reset=True)
    out, pay = ctx.compress_frames(None, offs, lens, comp, 0.05, 0, True, samples_ptr=fleet.data_ptr(), payload_out=pbuf)
    k = ctx.kernel_ms(reset=True)
    print(r, kind, "call ms", round(ctx.last_call_ms, 3), {a: round(b, 3) for a, b in k.items() if b > 0.01}, flush=True)
its = np.array([out[i].iterations for i in range(len(lens))])
print("iterations histogram", {int(a): int(b) for a, b in zip(*np.unique(its, return_counts=True))})
ctx.close()
