"""Where does a compress step's wall time go?  (run on the GPU box)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, atsc_b200, bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = atsc_b200.Context([0])
host = np.empty((S, bench.SERIES_LEN)); bench.make_fleet(S, 5000, host)
dev = torch.from_numpy(host.reshape(-1)).cuda()
offs, lens = bench.frame_table(S)
n = len(lens)
out = (atsc_b200.FrameOut * n)()
payload = np.empty(64 << 20, dtype=np.uint8)
used = C.c_uint64()
def raw():
    return ctx.L.atsc_gpu_compress_frames(ctx.h, C.c_void_p(dev.data_ptr()), offs.ctypes.data_as(C.POINTER(C.c_uint64)),
        lens.ctypes.data_as(C.POINTER(C.c_uint32)), n, 5, np.float32(0.05), 0, 1, out, C.c_void_p(payload.ctypes.data), len(payload), C.byref(used))
for _ in range(3): assert raw() == 0
ctx.kernel_ms(True)
t0 = time.perf_counter()
for _ in range(10): raw()
t1 = time.perf_counter()
k = ctx.kernel_ms(True)
print("raw C call ms/step", (t1 - t0) / 10 * 1e3, "kernel ms/step", sum(k.values()) / 10, {a: round(b / 10, 3) for a, b in k.items() if b})
t0 = time.perf_counter()
for _ in range(10): ctx.compress_frames(None, offs, lens, 5, 0.05, 0, True, samples_ptr=dev.data_ptr())
t1 = time.perf_counter()
print("python wrapper ms/step", (t1 - t0) / 10 * 1e3)
