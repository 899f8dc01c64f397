"""A/B of kernel variants (diagnostic, GPU box only): one process, fleets built once, variants switched
through ATSC_STATS_VARIANT / ATSC_POLY_VARIANT between calls.  Timing runs on the clean bench fleet;
parity compares every variant's frame table and payload with variant (0, 0) on a small hostile fleet
(zeros, -0.0, integer ramps, odd alignments, ragged lengths).
    python tools/ab_variants.py [--series 288] [--steps 10] [--combos s:p,s:p,...]"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, atsc_b200, bench

ap = argparse.ArgumentParser()
ap.add_argument("--series", type=int, default=288)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--combos", default="0:0,1:0,2:0,1:1,1:2,1:3,1:4,1:5,1:6")
a = ap.parse_args()
L = atsc_b200.load_library()
pcap = 128 << 20
pbuf = np.ctypeslib.as_array(C.cast(L.atsc_gpu_host_alloc(pcap), C.POINTER(C.c_uint8)), shape=(pcap,))

host = np.empty((a.series, bench.SERIES_LEN))
bench.make_fleet(a.series, 1000, host)
dev = torch.from_numpy(host.reshape(-1)).cuda()
offs, lens = bench.frame_table(a.series)
n = int(lens.astype(np.int64).sum())

HS = 12
hh = host[:HS].copy()
hh[1, 5:4000] = 0.0
hh[2, 70000:70010] = -0.0
hh[4, 131072:131072 + 300] = np.arange(300)
hh[5, :200000] = np.round(hh[5, :200000])
hh[7, 300000:300100] = -hh[7, 300000:300100]
hdev = torch.from_numpy(hh.reshape(-1)).cuda()
hoffs, hlens = bench.frame_table(HS)
for i in range(0, len(hlens), 3):  # odd alignment, ragged lengths
    if hlens[i] > 8:
        hoffs[i] += 1; hlens[i] -= 3


def table(out, pay):
    return (np.array([(o.compressor, o.iterations, o.payload_len, o.payload_off, *o.cand_size) for o in out], dtype=np.int64),
            bytes(pay))


ref = None
for combo in a.combos.split(","):
    sv, pv, *ev = combo.split(":")
    many = ev[0] if ev else "4"
    os.environ["ATSC_STATS_VARIANT"] = sv; os.environ["ATSC_POLY_VARIANT"] = pv
    res = {}
    for eng in ("1", many):
        os.environ["ATSC_ENGINES"] = eng
        ctx = atsc_b200.Context([0])
        run = lambda: ctx.compress_frames(None, offs, lens, atsc_b200.AUTO, bench.ERROR_PCT / 100.0, 0, True,
                                          samples_ptr=dev.data_ptr(), payload_out=pbuf)
        for _ in range(3):
            run()
        ctx.kernel_ms(reset=True)
        span = 0.0
        for _ in range(a.steps):
            run(); span += ctx.last_call_ms
        k = ctx.kernel_ms(reset=True)
        res[eng] = (span / a.steps, {kk: round(v / a.steps, 3) for kk, v in k.items() if v})
        if eng == many:
            cur = table(*ctx.compress_frames(None, hoffs, hlens, atsc_b200.AUTO, 0.05, 0, True, samples_ptr=hdev.data_ptr(), payload_out=pbuf))
            cur0 = table(*ctx.compress_frames(None, hoffs, hlens, atsc_b200.POLYNOMIAL, 0.0, 0, True, samples_ptr=hdev.data_ptr(), payload_out=pbuf))
        ctx.close()
    if ref is None:
        ref = (cur, cur0); same = "reference"
    else:
        ok = all(np.array_equal(x[0], y[0]) and x[1] == y[1] for x, y in zip((cur, cur0), ref))
        same = "IDENTICAL" if ok else f"DIFFERENT ({int((cur[0] != ref[0][0]).any(axis=1).sum())}+{int((cur0[0] != ref[1][0]).any(axis=1).sum())} frames)"
    print(f"stats={sv} poly={pv}: {many} engines {res[many][0]:.3f} ms/step = {n / res[many][0] / 1e6:.1f} Gsamples/s | 1 engine {res['1'][0]:.3f} ms "
          f"{ {k: v for k, v in res['1'][1].items() if k in ('stats', 'poly')} } | {same}", flush=True)
