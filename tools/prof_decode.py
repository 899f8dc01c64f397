"""Driver for profiling k_decode: `python tools/prof_decode.py c3|mixed [series]` -- decompression of the C3
polynomial fleet (65,536-sample gauge / utilisation / sawtooth series) or of the mixed 1 M-sample fleet."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import atsc_b200, bench, gen

kind = sys.argv[1] if len(sys.argv) > 1 else "c3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
ctx = atsc_b200.Context([0])
dev = torch.device("cuda", 0)
if kind == "c3":
    n = 65536
    h = np.empty((S, n))
    for s in range(S):
        h[s] = (gen.gauge_walk, gen.utilisation, gen.sawtooth)[s % 3](n, 1000 + s)
    fleet = torch.from_numpy(h).to(dev); comp = atsc_b200.POLYNOMIAL
else:
    n = bench.SERIES_LEN
    S = min(S, 288)
    fleet = bench.make_fleet_device(S, 5000, dev); comp = atsc_b200.AUTO
offs, lens = bench.frame_table(S, n)
pbuf = np.empty(512 << 20, dtype=np.uint8)
out, pay = ctx.compress_frames(None, offs, lens, comp, 0.05, 0, True, samples_ptr=fleet.data_ptr(), payload_out=pbuf)
pay = pay.copy()
steps = np.array([0])
frames = ctx.frames_in([(out[i].compressor, int(lens[i]), int(out[i].payload_off), int(out[i].payload_len), int(offs[i])) for i in range(len(lens))])
dout = torch.empty(S * n, dtype=torch.float64, device=dev)
for r in range(4):
    ctx.kernel_ms(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.decompress_frames(frames, pay, out_ptr=dout.data_ptr())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    k = ctx.kernel_ms(reset=True)["decode"]
    print(r, kind, "call ms", round(dt * 1e3, 3), "kernel ms", round(k, 3), "GB/s wall", round(S * n * 8 / dt / 1e9), "kernel", round(S * n * 8 / (k * 1e-3) / 1e9), "payload MB", round(len(pay) / 1e6, 1), flush=True)
comps = np.array([out[i].compressor for i in range(len(lens))])
print({atsc_b200.COMPRESSOR_NAMES[c]: int((comps == c).sum()) for c in np.unique(comps)}, "avg payload/frame", len(pay) // len(lens))
ctx.close()
