#!/usr/bin/env python
"""bench.py -- throughput of the ATSC hot path on B200 (contract: see the task statement).

Workload (config.workload): BASELINE.json config 4 ("auto selection at -c 0 ... mixed
constant/periodic/noisy") scaled to one GPU: S series x 1,000,000 samples per GPU, one third
constant, one third periodic (SURVEY.md C2 formula, sigma 0.05), one third noisy (C3 class b
utilisation gauge), `atsc --compressor auto -e 5 -c 0`; the 100k x 1M fleet of the config is
800 GB and does not fit, so each GPU processes a stated subsample per step (weak scaling: the
per-GPU share is fixed as N grows, frames never communicate, no collective).  The fleet is
generated ON THE DEVICE (same formulas as tests/gen.py, torch RNG) so that a step can be large:
1152 series = 9.2 GB per GPU, 16 waves of the library's pipeline.

One "step" = one pass of the hot path over the GPU's whole batch:
  value   : Msamples/s, inputs resident in HBM when the clock starts, payloads + frame
            records copied back to the host inside the timed region.
  e2e     : same call with the samples in pinned HOST memory (H2D inside the timed region).
  roofline: dominant kernel (largest summed CUDA-event time) -- algorithmic bytes = 8 B x samples of
            the frames it reads / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the oracle port (C restatement of the reference, rebuilt -O3 -march=native on this
            machine) on the host cores, bounded sample.
  decompress: GB/s of f64 output for the BRO fleet produced by the compress step.
  configs : (N = 1 only) the other named shapes of BASELINE.json -- FFT only on one 1 M-sample series,
            Polynomial / IDW on 65,536-sample series, auto at -c 6, an all-noise fleet, decompression
            of the polynomial fleet -- each with its throughput, fraction of the HBM line and the CPU
            port beside it on a bounded sample.
`--single-process`: ONE context sharding a call's frames over all visible GPUs (host buffers in,
payload gathered to the host) instead of one process per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SERIES_LEN = 1_000_000
ERROR_PCT = 5
SPEED = 0
DEFAULT_SERIES = 1152


def make_series(cls, seed):
    import gen
    if cls == 0:
        return gen.constant(SERIES_LEN, seed)
    if cls == 1:
        return gen.periodic(SERIES_LEN, seed, sigma=0.05)
    return gen.utilisation(SERIES_LEN, seed)


def make_fleet(n_series, seed0, out):
    """Fills out[n_series, SERIES_LEN] with the mixed fleet (class = series index mod 3); numpy, host."""
    for s in range(n_series):
        out[s, :] = make_series(s % 3, seed0 + s)


def make_fleet_device(n_series, seed0, device, kinds=(0, 1, 2), n=SERIES_LEN):
    """The same three classes generated on the GPU (torch RNG): constant (seed mod 1000), periodic
    100 + 20 sin(2 pi i / 1440) + 5 sin(2 pi i / 97) + 0.05 N(0,1), utilisation clip(50 + 30 sin(2 pi i / 4320) +
    AR(1; phi 0.9, sigma 2), 0.01, 100) rounded to 2 decimals (the AR(1) filter as a 256-tap convolution:
    0.9^256 = 2e-12).  kinds: class of series s = kinds[s % len(kinds)]; 3 = positive white noise (gen.noisy)."""
    import torch
    out = torch.empty((n_series, n), dtype=torch.float64, device=device)
    g = torch.Generator(device=device)
    i = torch.arange(n, dtype=torch.float64, device=device)
    per = 100.0 + 20.0 * torch.sin(2 * np.pi * i / 1440.0) + 5.0 * torch.sin(2 * np.pi * i / 97.0)
    util = 50.0 + 30.0 * torch.sin(2 * np.pi * i / 4320.0)
    taps = (0.9 ** torch.arange(255, -1, -1, dtype=torch.float64, device=device)).view(1, 1, -1)
    for s in range(n_series):
        cls = kinds[s % len(kinds)]
        g.manual_seed(seed0 + s)
        if cls == 0:
            out[s].fill_(float((seed0 + s) % 1000))
        elif cls == 1:
            out[s] = per + 0.05 * torch.randn(n, generator=g, dtype=torch.float64, device=device)
        elif cls == 2:
            e = torch.randn(n + 255, generator=g, dtype=torch.float64, device=device) * (2.0 * np.sqrt(1 - 0.81))
            ar = torch.nn.functional.conv1d(e.view(1, 1, -1), taps).view(-1)
            out[s] = torch.round(torch.clamp(util + ar, 0.01, 100.0), decimals=2)
        else:
            out[s] = torch.round(1.0 + 99.0 * torch.rand(n, generator=g, dtype=torch.float64, device=device), decimals=3)
    return out


def frame_table(n_series, series_len=SERIES_LEN):
    import atsc_b200
    cs = atsc_b200.chunk_sizes(series_len)
    offs, lens = [], []
    for s in range(n_series):
        o = s * series_len
        for c in cs:
            offs.append(o)
            lens.append(c)
            o += c
    return np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint32)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML at 20 Hz -- a 500 Hz poller per
    rank showed up in the 8-rank numbers of round 1; falls back to polling nvidia-smi)."""

    def __init__(self, gpu):
        self.gpu = gpu
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop = False
        self.th = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)  # first call is slow: not inside the timed region
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while True:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
            except Exception:
                pass
            if self.stop:  # at least one sample, taken while the region is still open
                break
            time.sleep(0.05)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                self.sm.append(float(r[0]))
                self.mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run_nvml if self.nvml else self._run_smi, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def bind_cores(local, world):
    """Each rank on the cores `nvidia-smi topo -m` lists as its GPU's CPU affinity, split among the ranks that
    share them (8 ranks on one socket contended for the same cores in round 1).  Best effort."""
    try:
        rows = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()
        hdr = next(r for r in rows if "CPU Affinity" in r)
        col = hdr.split("\t").index("CPU Affinity") if "\t" in hdr else None
        sets = {}
        for r in rows:
            if r.startswith("GPU") and col is not None:
                cells = r.split("\t")
                gi = int(cells[0].strip()[3:])
                cores = set()
                for part in cells[col].strip().split(","):
                    if "-" in part:
                        a, b = part.split("-")
                        cores.update(range(int(a), int(b) + 1))
                    elif part.strip().isdigit():
                        cores.add(int(part))
                sets[gi] = cores
        mine = sorted(sets[local] & os.sched_getaffinity(0))
        sharers = sorted(g for g in sets if g < world and sets[g] == sets[local])
        k = sharers.index(local)
        share = mine[k * len(mine) // len(sharers):(k + 1) * len(mine) // len(sharers)]
        if len(share) >= 2:
            os.sched_setaffinity(0, share)
            return f"{share[0]}-{share[-1]} ({len(share)} cores)"
    except Exception:
        pass
    return None


_native = None


def oracle():
    """The oracle library for the TIMED CPU legs: rebuilt -O3 -march=native on this machine when possible."""
    global _native
    import oracle_lib as O
    if _native is None:
        _native = O.use_native()
        if not _native:
            O.lib()
    return O, ("-O3 -march=native build of oracle/atsc_oracle.c" if _native else "portable -O2 build of oracle/atsc_oracle.c")


def cpu_port_rate(arr2d, compressor, error_pct, speed, threads, budget_s=6.0):
    """Msamples/s of the CPU port on a small sample (rows = series), best of up to 3."""
    O, how = oracle()
    n, sl = arr2d.shape
    t0 = time.perf_counter()
    O.compress_batch(arr2d, compressor, error_pct, speed, threads)
    dt = time.perf_counter() - t0
    reps = 1
    while dt * (reps + 1) / reps < budget_s and reps < 3:
        t1 = time.perf_counter()
        O.compress_batch(arr2d, compressor, error_pct, speed, threads)
        dt = min(dt, time.perf_counter() - t1)
        reps += 1
    return n * sl / dt / 1e6


def cpu_baseline(threads, budget_s=20.0):
    """Oracle port timed on the host cores on a bounded sample of the same workload."""
    O, how = oracle()
    per_class = 4 * max(1, threads // 3)
    n = per_class * 3
    arr = np.empty((n, SERIES_LEN))
    make_fleet(n, 5000, arr)
    t0 = time.perf_counter()
    O.compress_batch(arr, O.AUTO, ERROR_PCT, SPEED, threads)
    dt = time.perf_counter() - t0
    reps = 1
    # repeat while cheap so the figure is not a single noisy shot
    while dt * (reps + 1) / reps < budget_s and reps < 3:
        t1 = time.perf_counter()
        O.compress_batch(arr, O.AUTO, ERROR_PCT, SPEED, threads)
        dt = min(dt, time.perf_counter() - t1)
        reps += 1
    return {"value": n * SERIES_LEN / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": f"{n} series x {SERIES_LEN} samples ({per_class} per class), best of {reps}, "
                      f"{threads} threads, one series per thread at a time (reference is single-threaded per series); {how}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust
    reference cannot be built in this image) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O, how = oracle()
    threads = os.cpu_count() or 1
    per_class = 4 * max(1, threads // 3)
    n = per_class * 3
    arr = np.empty((n, SERIES_LEN))
    make_fleet(n, 5000, arr)
    for _ in range(args.warmup):
        O.compress_batch(arr[:3], O.AUTO, ERROR_PCT, SPEED, min(3, threads))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.compress_batch(arr, O.AUTO, ERROR_PCT, SPEED, threads)
    dt = time.perf_counter() - t0
    v = args.steps * n * SERIES_LEN / dt / 1e6
    sample = f"{n} series x {SERIES_LEN} samples per step ({per_class} per class), {threads} threads; {how}"
    line = {
        "impl": "reference", "metric": "auto_compress_msamples_per_s", "value": v, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": workload_config(args.series, args.gpus),
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(series_per_gpu, n_gpus):
    return {
        "workload": f"atsc --compressor auto -e {ERROR_PCT} -c {SPEED}: {series_per_gpu} series/GPU x {SERIES_LEN} samples "
                    "(1/3 constant, 1/3 periodic sigma=0.05, 1/3 noisy utilisation gauge), "
                    "frames [131072x7,65536,16384,512,64] per series; subsample of BASELINE config 4 (100k x 1M)",
        "series_per_gpu": series_per_gpu, "series_len": SERIES_LEN, "error_pct": ERROR_PCT, "speed": SPEED,
        "parallelism": f"frames sharded by series over {n_gpus} GPU(s), no collective",
        "l2_policy": f"inputs ({series_per_gpu * SERIES_LEN * 8 / 1e6:.0f} MB per GPU per step) exceed the 126 MB L2",
    }


def extra_configs(ctx, torch, dev_index, peak, threads, host_threads_note):
    """The other named shapes of BASELINE.json (N = 1): throughput through the same C ABI with device-resident
    inputs, fraction of the HBM line (8 B per sample), the CPU port on a bounded sample beside it."""
    import atsc_b200
    A = atsc_b200
    device = torch.device("cuda", dev_index)
    pbuf = np.empty(512 << 20, dtype=np.uint8)
    res = {}

    def run(name, fleet, comp, err, speed, reps, cpu_rows, cpu_comp=None, note=""):
        S, n = fleet.shape
        offs, lens = frame_table(S, n)
        call = lambda: ctx.compress_frames(None, offs, lens, comp, err / 100.0, speed, True,  # noqa: E731
                                           samples_ptr=fleet.data_ptr(), payload_out=pbuf)
        for _ in range(2):
            out, pay = call()
        ctx.kernel_ms(reset=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        span = 0.0
        for _ in range(reps):
            out, pay = call()
            span += ctx.last_call_ms
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        k = {a: round(b / reps, 4) for a, b in ctx.kernel_ms(reset=True).items() if b and not a.startswith("reserved")}
        comps = np.array([out[i].compressor for i in range(len(lens))])
        host = fleet[:cpu_rows].cpu().numpy()
        cpu = cpu_port_rate(host, cpu_comp if cpu_comp is not None else comp, err, speed, min(threads, cpu_rows))
        ms = dt / reps * 1e3
        res[name] = {
            "msamples_per_s": S * n / dt * reps / 1e6, "ms_per_call": ms, "device_ms_per_call": span / reps,
            "samples_per_call": int(S * n), "frac_of_hbm_line": S * n * 8 / (dt / reps) / 1e9 / peak,
            "winners": {A.COMPRESSOR_NAMES[c]: int((comps == c).sum()) for c in np.unique(comps)},
            "compressed_bytes": int(len(pay)), "near_tie_frames": int(sum(1 for i in range(len(lens)) if out[i].near_tie)),
            "kernel_ms_per_call": k,
            "cpu_port_msamples_per_s": cpu, "cpu_sample": f"{cpu_rows} series x {n} samples, {min(threads, cpu_rows)} threads{host_threads_note}",
            "note": note,
        }
        return out, pay, offs, lens

    # C2: FFT compressor only, one synthetic 1 M-sample sinusoid + noise series (sigma 0.5), -e 1 / 5 / 10
    import gen
    c2 = torch.from_numpy(gen.periodic(SERIES_LEN, 42, sigma=0.5)).to(device).view(1, -1)
    for e in (1, 5, 10):
        run(f"C2_fft_1M_e{e}", c2, A.FFT, e, 0, 10, 1, note="one series: 11 frames, latency of the refinement loop, not throughput")
    c2f = torch.from_numpy(np.stack([gen.periodic(SERIES_LEN, 42 + s, sigma=0.5) for s in range(48)])).to(device)
    run("C2_fft_48x1M_e5", c2f, A.FFT, 5, 0, 5, min(threads, 16), note="the same class as a fleet: 48 series, FFT compressor only")
    del c2f
    # C3: Polynomial and IDW on 65,536-sample monitoring series (gauge walk / utilisation / sawtooth), -e 5
    S3 = 3072
    h3 = np.empty((S3, 65536))
    for s in range(S3):
        h3[s] = (gen.gauge_walk, gen.utilisation, gen.sawtooth)[s % 3](65536, 1000 + s)
    c3 = torch.from_numpy(h3).to(device)
    out3, pay3, offs3, lens3 = run("C3_polynomial_3072x64k", c3, A.POLYNOMIAL, 5, 0, 5, min(threads, 48) * 2,
                                   note="subsample of the 10k x 64k fleet (3072 series, three classes)")
    # the payload stays page-locked for C5 (its H2D is inside the timed call); pbuf is reused by the next configs
    q3 = ctx.L.atsc_gpu_host_alloc(max(len(pay3), 1))
    import ctypes as C
    pin3 = np.ctypeslib.as_array(C.cast(q3, C.POINTER(C.c_uint8)), shape=(max(len(pay3), 1),))
    pin3[:len(pay3)] = pay3
    pay3 = pin3[:len(pay3)]
    run("C3_idw_96x64k", c3[:96], A.IDW, 5, 0, 3, max(3, min(threads, 12)), note="IDW is O(N K) per frame: 96 series")
    # C5: decompression of the C3 polynomial fleet
    frames_in = ctx.frames_in([(out3[i].compressor, int(lens3[i]), int(out3[i].payload_off), int(out3[i].payload_len), int(offs3[i]))
                               for i in range(len(lens3))])
    dout = torch.empty(S3 * 65536, dtype=torch.float64, device=device)
    for _ in range(2):
        ctx.decompress_frames(frames_in, pay3, out_ptr=dout.data_ptr())
    ctx.kernel_ms(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ctx.decompress_frames(frames_in, pay3, out_ptr=dout.data_ptr())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    dec_ms = ctx.kernel_ms(reset=True)["decode"] / 10
    # CPU port: decompress 96 series of the same fleet
    O, how = oracle()
    bros = ctx.compress_data([h3[s] for s in range(96)], compressor=A.POLYNOMIAL, error=5)
    boff = np.concatenate([[0], np.cumsum([len(b) for b in bros])]).astype(np.uint64)
    blob = np.frombuffer(b"".join(bros), dtype=np.uint8).copy()
    hout = np.empty((96, 65536))
    t1 = time.perf_counter()
    O.lib().atsc_oracle_decompress_batch(blob.ctypes.data_as(C.POINTER(C.c_uint8)), boff.ctypes.data_as(C.POINTER(C.c_uint64)),
                                         96, 65536, min(threads, 96), hout.ctypes.data_as(C.POINTER(C.c_double)))
    cdt = time.perf_counter() - t1
    res["C5_decompress_C3_polynomial"] = {
        "gb_per_s_f64_out": S3 * 65536 * 8 / dt / 1e9, "ms_per_call": dt * 1e3, "kernel_gb_per_s": S3 * 65536 * 8 / (dec_ms * 1e-3) / 1e9,
        "frac_of_hbm_line": S3 * 65536 * 8 / dt / 1e9 / peak, "payload_bytes": int(len(pay3)),
        "cpu_port_gb_per_s": 96 * 65536 * 8 / cdt / 1e9, "cpu_sample": f"96 series x 65536 samples, {min(threads, 96)} threads",
    }
    del c3, dout
    ctx.L.atsc_gpu_host_free(q3)
    # C4 at -c 6 (sampled selection) and an all-noise fleet at -c 0
    c4 = make_fleet_device(288, 5000, device)
    run("C4_auto_c6_288x1M", c4, A.AUTO, 5, 6, 10, min(threads, 12), note="sampled selection: 128-sample probe frames pick the compressor")
    del c4
    noisy = make_fleet_device(24, 9000, device, kinds=(3,))
    run("noisy_auto_c0_24x1M", noisy, A.AUTO, 5, 0, 3, min(threads, 12),
        note="positive white noise: no candidate reaches 5 % cheaply, the FFT refinement loop runs all 23 iterations")
    return res


def run_single_process(args):
    """One context over every visible GPU: frames sharded by the library, payload gathered to the host."""
    import torch
    import atsc_b200
    import ctypes as C
    n = min(args.gpus, torch.cuda.device_count()) if args.gpus > 1 else torch.cuda.device_count()
    # host-resident fleet: 288 series (2.3 GB of page-locked memory) per GPU unless --series asks for less
    args.series = min(args.series, 288)
    S = args.series * n
    ctx = atsc_b200.Context(list(range(n)))
    L = ctx.L
    hptr = L.atsc_gpu_host_alloc(S * SERIES_LEN * 8)
    host = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(S, SERIES_LEN))
    for g in range(n):
        torch.cuda.set_device(g)
        fl = make_fleet_device(args.series, 5000 + g * args.series, torch.device("cuda", g))
        host[g * args.series:(g + 1) * args.series] = fl.cpu().numpy()
        del fl
    offs, lens = frame_table(S)
    pcap = max(256 << 20, S * SERIES_LEN // 8)  # the mixed fleet compresses to ~0.05 B per sample
    pptr = L.atsc_gpu_host_alloc(pcap)
    pbuf = np.ctypeslib.as_array(C.cast(pptr, C.POINTER(C.c_uint8)), shape=(pcap,))
    call = lambda: ctx.compress_frames(host.reshape(-1), offs, lens, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True, payload_out=pbuf)  # noqa: E731
    for _ in range(max(args.warmup, 2)):
        out, pay = call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, pay = call()
    dt = time.perf_counter() - t0
    v = S * SERIES_LEN * args.steps / dt / 1e6
    line = {"metric": "auto_compress_msamples_per_s", "value": v, "unit": "Msamples/s", "n_gpus": n, "steps": args.steps,
            "warmup": max(args.warmup, 2), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "mode": "single-process: one atsc_ctx over all GPUs, host buffers in (H2D inside), payload gathered to the host",
            "config": dict(workload_config(args.series, n), parallelism=f"one context sharding a call's frames over {n} GPU(s) by frame range"),
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": S * SERIES_LEN * 8, "d2h_bytes_per_step": int(len(pay)) + len(lens) * 64},
            "compressed_bytes": int(len(pay)), "gpu_launches": int(ctx.launches)}
    print(json.dumps(line))
    L.atsc_gpu_host_free(hptr)
    L.atsc_gpu_host_free(pptr)
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--series", type=int, default=DEFAULT_SERIES, help="series per GPU (multiple of 3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other named shapes (N = 1 runs them by default)")
    ap.add_argument("--single-process", action="store_true")
    ap.add_argument("--as-rank", type=int, default=None, help="diagnostic: generate the fleet rank R of a multi-GPU run would get")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import atsc_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    if args.single_process:
        return run_single_process(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cores = bind_cores(local, world) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout when the first communicator is created: send fd 1 to
        # stderr meanwhile, so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    warmup = max(args.warmup, 3)
    S = args.series
    ctx = atsc_b200.Context([local])
    L = ctx.L
    device = torch.device("cuda", local)

    # ---- inputs: device fleet (for value) + pinned host copy (for e2e)
    nbytes = S * SERIES_LEN * 8
    dev2d = make_fleet_device(S, 5000 + (rank if args.as_rank is None else args.as_rank) * S, device)
    dev = dev2d.view(-1)
    hptr = L.atsc_gpu_host_alloc(nbytes)
    import ctypes as C
    host = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(S, SERIES_LEN))
    torch.from_numpy(host).copy_(dev2d)
    torch.cuda.synchronize()
    offs, lens = frame_table(S)
    n_samples = S * SERIES_LEN

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # page-locked result buffer: payload bytes land in it straight from the device
    pcap = max(64 << 20, S * 80_000)
    pptr = L.atsc_gpu_host_alloc(pcap)
    pbuf = np.ctypeslib.as_array(C.cast(pptr, C.POINTER(C.c_uint8)), shape=(pcap,))

    def step_dev():
        return ctx.compress_frames(None, offs, lens, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                   samples_ptr=dev.data_ptr(), payload_out=pbuf)

    def step_host():
        return ctx.compress_frames(host.reshape(-1), offs, lens, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                   payload_out=pbuf)

    dev_ms = {}

    def timed(fn, steps, tag=None):
        """K steps bracketed by barrier + synchronize; returns (wall seconds, max over ranks; last result).
        The library's own CUDA events give the device span of every call (dev_ms[tag] = their sum, max over
        ranks): the two differ only by the host gaps between calls."""
        barrier()
        t0 = time.perf_counter()
        span = 0.0
        for _ in range(steps):
            r = fn()
            span += ctx.last_call_ms
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, span], dtype=torch.float64, device="cuda")
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            if tag:
                dev_ms[tag + "_per_rank"] = [round(float(x[1].item()) / steps, 4) for x in allt]
            dt, span = max(float(x[0].item()) for x in allt), max(float(x[1].item()) for x in allt)
        if tag:
            dev_ms[tag] = span / steps
        return dt, r

    for _ in range(warmup):
        out, payload = step_dev()
    ctx.kernel_ms(reset=True)
    l0 = ctx.launches
    with ClockSampler(local) as clk:
        dt, (out, payload) = timed(step_dev, args.steps, "compress")
    launches = ctx.launches - l0
    kms = ctx.kernel_ms(reset=True)
    value = world * n_samples * args.steps / dt / 1e6

    # ---- e2e: host buffers, H2D inside
    e2e_steps = max(3, min(args.steps, 10))
    step_host()
    dt_e2e, _ = timed(step_host, e2e_steps, "compress_e2e")
    e2e_v = world * n_samples * e2e_steps / dt_e2e / 1e6
    d2h = int(len(payload)) + len(lens) * 176  # payload + one FrameWork record per frame (csrc/common.cuh)

    # ---- roofline of the dominant kernel
    comps = np.array([out[i].compressor for i in range(len(lens))])
    near_ties = int(sum(1 for i in range(len(lens)) if out[i].near_tie))
    nonconst = comps != atsc_b200.CONSTANT
    fft_samples = int(lens[nonconst].astype(np.int64).sum())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(("stats", "poly", "rle", "fft_fwd", "fft_small", "fft", "select", "emit", "front"), key=lambda k: kms[k])  # host_issue is not a kernel
    dom_ms = kms[dom] / args.steps
    front_mode = os.environ.get("ATSC_FRONT", "2")  # api.cu: 2 = k_sfold (default), 1 = k_front, 0 = separate passes
    kname = {"front": "k_sfold" if front_mode == "2" else "k_front",
             "poly": "k_poly" if os.environ.get("ATSC_POLY_ITEMS", "1") == "0" else
                     "k_plan+k_poly1s+k_poly" if os.environ.get("ATSC_POLY1_STATIC", "1") != "0" else "k_plan+k_poly1+k_poly"}
    p1name = "k_poly1s" if os.environ.get("ATSC_POLY1_STATIC", "1") != "0" else "k_poly1"
    big_samples = int(lens[lens >= 16384].astype(np.int64).sum())  # the frames the front-end kernel takes
    fftwin_samples = int(lens[comps == atsc_b200.FFT].astype(np.int64).sum())  # frames k_fft_fwd transforms in full
    dom_samples = {"stats": n_samples - (big_samples if front_mode != "0" else 0), "select": n_samples, "front": big_samples,
                   "fft_fwd": fft_samples if front_mode == "0" else fftwin_samples}.get(dom, fft_samples)
    achieved = dom_samples * 8 / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # the same kernels timed without other waves sharing the GPU (one engine): explains how much of
    # the in-pipeline duration above is contention
    iso = None
    try:
        os.environ["ATSC_ENGINES"] = "1"
        ctx1 = atsc_b200.Context([local])
        offs1, lens1 = frame_table(288)
        for _ in range(2):
            ctx1.compress_frames(None, offs1, lens1, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                 samples_ptr=dev.data_ptr(), payload_out=pbuf)
        ctx1.kernel_ms(reset=True)
        for _ in range(3):
            ctx1.compress_frames(None, offs1, lens1, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                 samples_ptr=dev.data_ptr(), payload_out=pbuf)
        k1 = ctx1.kernel_ms(reset=True)
        ctx1.close()
        iso = {k: v / 3 for k, v in k1.items() if v}
    except Exception as ex:  # the figure is explanatory only
        iso = {"error": str(ex)}
    finally:
        os.environ.pop("ATSC_ENGINES", None)
    # dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed
    # ncu --set full capture (profiles/traffic.json): that launch is one 72-series wave, whose
    # algorithmic bytes are stated beside it
    traffic, traffic_detail = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if dom == "poly" and p1name in tj and os.environ.get("ATSC_POLY_ITEMS", "1") != "0":
            # the 'poly' slot times k_plan + k_poly1s + k_poly: one wave's launches of the two sample-reading kernels
            pk = "k_poly_after_poly1s" if p1name == "k_poly1s" and "k_poly_after_poly1s" in tj else "k_poly"
            traffic_detail = {p1name: tj[p1name], "k_poly": {k: v for k, v in tj[pk].items() if k != "whole_frames"}}
            traffic = tj[p1name]["bytes_per_launch"] + tj[pk]["bytes_per_launch"]
        else:
            traffic_detail = tj.get(kname.get(dom, "k_" + dom))
            if traffic_detail:
                traffic = traffic_detail["bytes_per_launch"]
    except Exception:
        pass
    # algorithmic bytes of a 288-series call: every sample once for the stats pass (k_stats + the front-end kernel),
    # the non-constant frames for k_poly1s + k_poly, the frames transformed in full for k_fft_fwd
    per_series = 288.0 / S
    alg1 = {"stats": 288 * SERIES_LEN * 8, "poly": 192 * SERIES_LEN * 8}
    if front_mode == "0":  # with k_sfold + k_probe the forward kernel only sees surviving frames: a latency chain, no roofline
        alg1["fft_fwd"] = fft_samples * per_series * 8
    one_engine = None
    if iso and "error" not in iso:
        iso1 = dict(iso)
        iso1["stats"] = iso.get("stats", 0.0) + iso.get("front", 0.0)
        names1 = {"stats": "k_stats" if front_mode == "0" else "k_stats+" + kname["front"], "poly": kname["poly"],
                  "fft_fwd": "k_fft_fwd"}
        one_engine = {names1[k]: {"ms_per_288_series": round(iso1[k], 4), "achieved": alg1[k] / (iso1[k] * 1e-3) / 1e9,
                                  "frac": alg1[k] / (iso1[k] * 1e-3) / 1e9 / peak}
                      for k in alg1 if iso1.get(k)}
    roofline = {"bound": "hbm", "kernel": kname.get(dom, "k_" + dom), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_detail": traffic_detail, "one_engine": one_engine,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_step": dom_samples * 8,
                "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items() if v and not k.startswith("reserved")},
                "kernel_ms_per_288_series_one_engine": iso,
                "whole_step_frac": n_samples * 8 / (dt / args.steps) / 1e9 / peak,
                "note": "achieved / frac / kernel_ms_per_step: CUDA events on each engine's stream inside the timed region "
                        "(waves of several engines overlap, so these durations include contention and add up to more "
                        "than the step); one_engine: the same kernels on a 288-series call with a single engine, i.e. each "
                        "launch alone on the GPU, which is the figure to hold against the kernel's own roofline (k_sfold / k_front "
                        "are reported under 'front'; 'poly' covers k_plan + k_poly1s + k_poly); the Polynomial step is bound by the "
                        "schedulers' issue ports, not by HBM: 19 FP64 instructions per sample, each holding its port for two "
                        "cycles (issue_model below, tools/ubench/p1arith.cu), the stats pass runs nearer the HBM line (DESIGN.md "
                        "section 4); the HBM line is the task's stated denominator",
                "issue_model": {
                    "kernel": "k_poly1s",
                    "cycles_per_4_samples_per_warp": 194,
                    "derivation": "118 instructions per trip of 4 samples, 76 of them FP64 (2 issue cycles each on B200): 2 * 76 + 42",
                    "samples_per_cycle_per_sm_at_full_issue": 4 * 32 * 4 / 194.0,
                    "ubench_samples_per_cycle_per_sm": 2.6,
                    "hbm_equivalent_gbs_at_full_issue": 4 * 32 * 4 / 194.0 * 148 * 1.965 * 8,
                    "source": "tools/ubench/p1arith.cu on B200 (arithmetic from registers only), profiles/r2_p1_variants.md"}}

    # ---- decompression of the fleet just produced (device-resident output)
    frames_in = []
    oo = 0
    for i in range(len(lens)):
        o = out[i]
        frames_in.append((o.compressor, int(lens[i]), int(o.payload_off), int(o.payload_len), oo))
        oo += int(lens[i])
    frames_in = ctx.frames_in(frames_in)
    # the payload stays page-locked (its H2D is inside the timed region): a second pinned buffer, because
    # pbuf is reused by later compress calls
    qptr = L.atsc_gpu_host_alloc(max(len(payload), 1))
    qbuf = np.ctypeslib.as_array(C.cast(qptr, C.POINTER(C.c_uint8)), shape=(max(len(payload), 1),))
    qbuf[:len(payload)] = payload
    payload = qbuf[:len(payload)]
    dout = dev  # the input fleet is no longer needed on the device: decode over it
    dec_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.decompress_frames(frames_in, payload, out_ptr=dout.data_ptr())
    ctx.kernel_ms(reset=True)
    dt_dec, _ = timed(lambda: ctx.decompress_frames(frames_in, payload, out_ptr=dout.data_ptr()), dec_steps, "decompress")
    dec_ms = ctx.kernel_ms(reset=True)["decode"] / dec_steps
    hout = host.reshape(-1)  # page-locked; the input fleet is no longer needed
    dt_dec_e2e, _ = timed(lambda: ctx.decompress_frames(frames_in, payload, out=hout), 3)
    dec = {"value": world * n_samples * 8 * dec_steps / dt_dec / 1e9, "unit": "GB/s (f64 out)",
           "kernel_gbs": n_samples * 8 / (dec_ms * 1e-3) / 1e9 if dec_ms else None,
           "kernel_frac_of_hbm_peak": (n_samples * 8 / (dec_ms * 1e-3) / 1e9 / peak) if dec_ms else None,
           "e2e_gbs": world * n_samples * 8 * 3 / dt_dec_e2e / 1e9,
           "payload_bytes": int(len(payload))}

    # ---- sanity: the timed result is a real compress (sizes, winners)
    names = atsc_b200.COMPRESSOR_NAMES
    hist = {names[c]: int((comps == c).sum()) for c in np.unique(comps)}

    line = None
    if rank == 0:
        threads = os.cpu_count() or 1
        cpu = None
        configs = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(threads)
        if world == 1 and not args.no_configs:
            L.atsc_gpu_host_free(hptr)
            hptr = None
            del dev2d, dev, dout
            torch.cuda.empty_cache()
            try:
                configs = extra_configs(ctx, torch, local, peak, threads, "")
            except Exception as ex:  # the headline must not die with an extra
                configs = {"error": repr(ex)}
        line = {
            "metric": "auto_compress_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": dt / args.steps * 1e3,
            "device_ms_per_step": dev_ms, "host_issue_ms_per_step": kms["host_issue"] / args.steps, "cpu_binding": cores,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32",
            "dtype_note": "f64 for stats / polynomial / rle / error metrics, f32 for the FFT (Complex<f32> in the reference)",
            "data": "synthetic (generated on the device; same formulas as tests/gen.py)", "config": workload_config(S, world),
            "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches, "near_tie_frames": near_ties, "frames_per_step": int(len(lens)),
            "roofline": roofline, "cpu_baseline": cpu, "decompress": dec,
            "clocks": clk.summary(), "winners": hist, "compressed_bytes": dec["payload_bytes"],
            "compression_ratio": n_samples * 8 / max(1, dec["payload_bytes"]), "configs": configs,
        }
        print(json.dumps(line))
    if hptr is not None:
        L.atsc_gpu_host_free(hptr)
    L.atsc_gpu_host_free(pptr)
    L.atsc_gpu_host_free(qptr)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
