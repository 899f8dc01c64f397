"""Per-class kernel-time breakdown of the auto path (diagnostic, GPU box only).
    python tools/class_profile.py [--series 48] [--classes constant,periodic,util,...]"""
import argparse, os, sys, time, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, gen, atsc_b200

ap = argparse.ArgumentParser()
ap.add_argument("--series", type=int, default=48)
ap.add_argument("--len", type=int, default=1_000_000)
ap.add_argument("--classes", default="constant,periodic,util,gauge,saw,noisy")
ap.add_argument("--error", type=int, default=5)
ap.add_argument("--speed", type=int, default=0)
ap.add_argument("--comp", default="auto")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
comp = {"auto": atsc_b200.AUTO, "fft": atsc_b200.FFT, "poly": atsc_b200.POLYNOMIAL, "idw": atsc_b200.IDW,
        "rle": atsc_b200.RLE}[a.comp]
ctx = atsc_b200.Context([0])
cs = atsc_b200.chunk_sizes(a.len)
for cls in a.classes.split(","):
    sigma = {}
    arr = np.empty((a.series, a.len))
    for s in range(a.series):
        arr[s] = gen.periodic(a.len, 5000 + s, sigma=0.05) if cls == "periodic" else gen.make(cls, a.len, 5000 + s)
    dev = torch.from_numpy(arr.reshape(-1)).cuda()
    offs, lens = [], []
    for s in range(a.series):
        o = s * a.len
        for c in cs:
            offs.append(o); lens.append(c); o += c
    offs = np.array(offs, dtype=np.uint64); lens = np.array(lens, dtype=np.uint32)
    for _ in range(2):
        out, pay = ctx.compress_frames(None, offs, lens, comp, a.error / 100.0, a.speed, True, samples_ptr=dev.data_ptr())
    ctx.kernel_ms(reset=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.reps):
        out, pay = ctx.compress_frames(None, offs, lens, comp, a.error / 100.0, a.speed, True, samples_ptr=dev.data_ptr())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / a.reps
    k = ctx.kernel_ms(reset=True)
    n = a.series * a.len
    big = [i for i in range(len(lens)) if lens[i] == 131072]
    win = collections.Counter(atsc_b200.COMPRESSOR_NAMES[out[i].compressor] for i in range(len(lens)))
    its = collections.Counter((atsc_b200.COMPRESSOR_NAMES[out[i].compressor], out[i].iterations) for i in big)
    print(f"== {cls}: {n/dt/1e9:.2f} Gsamples/s wall {dt*1e3:.3f} ms; kernels/step:",
          {kk: round(v / a.reps, 3) for kk, v in k.items() if v}, "payload", len(pay))
    print("   winners", dict(win), "| big-frame (winner, iters):", dict(its))
    i = big[0] if big else 0
    print("   frame0 cand sizes [fft,poly,rle]", list(out[i].cand_size), "errs", [round(e, 5) for e in out[i].cand_error])
    del dev
ctx.close()
