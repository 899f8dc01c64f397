"""Wave-size sweep on the bench fleet (diagnostic, GPU box only): device span per step for several
ATSC_WAVE_MI settings, four engines.   python tools/wave_size_sweep.py [--mi 48,56,60,64,72,80]"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, atsc_b200, bench

ap = argparse.ArgumentParser()
ap.add_argument("--mi", default="48,56,60,64,72,80")
ap.add_argument("--series", type=int, default=288)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
base = np.empty((24, bench.SERIES_LEN)); bench.make_fleet(24, 1000, base)
host = np.tile(base, (a.series // 24, 1))       # classes stay s % 3; timing only
dev = torch.from_numpy(host.reshape(-1)).cuda()
offs, lens = bench.frame_table(a.series)
n = int(lens.astype(np.int64).sum())
L = atsc_b200.load_library(); pcap = 64 << 20
pbuf = np.ctypeslib.as_array(C.cast(L.atsc_gpu_host_alloc(pcap), C.POINTER(C.c_uint8)), shape=(pcap,))
for mi in a.mi.split(","):
    os.environ["ATSC_WAVE_MI"] = mi
    ctx = atsc_b200.Context([0])
    run = lambda: ctx.compress_frames(None, offs, lens, atsc_b200.AUTO, 0.05, 0, True, samples_ptr=dev.data_ptr(), payload_out=pbuf)
    for _ in range(3):
        run()
    span = 0.0
    for _ in range(a.steps):
        run(); span += ctx.last_call_ms
    print(f"ATSC_WAVE_MI={mi}: {span / a.steps:.3f} ms/step = {n / (span / a.steps) / 1e6:.1f} Gsamples/s", flush=True)
    ctx.close()
