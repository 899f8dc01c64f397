#!/usr/bin/env python
"""bench.py -- throughput of the ATSC hot path on B200 (contract: see the task statement).

Workload (config.workload): BASELINE.json config 4 ("auto selection at -c 0 ... mixed
constant/periodic/noisy") scaled to one GPU: S series x 1,000,000 samples per GPU, one third
constant, one third periodic (SURVEY.md C2 formula, sigma 0.05), one third noisy (C3 class b
utilisation gauge), `atsc --compressor auto -e 5 -c 0`; the 100k x 1M fleet of the config is
800 GB and does not fit, so each GPU processes a stated subsample per step (weak scaling: the
per-GPU share is fixed as N grows, frames never communicate, no collective).

One "step" = one pass of the hot path over the GPU's whole batch:
  value   : Msamples/s, inputs resident in HBM when the clock starts, payloads + frame
            records copied back to the host inside the timed region.
  e2e     : same call with the samples in pinned HOST memory (H2D inside the timed region).
  roofline: dominant kernel (largest summed CUDA-event time) -- algorithmic bytes = 8 B x samples of
            the frames it reads / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline: the oracle port (C restatement of the reference) on the host cores, bounded sample.
  decompress: GB/s of f64 output for the BRO fleet produced by the compress step.
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


def make_series(cls, seed):
    import gen
    if cls == 0:
        return gen.constant(SERIES_LEN, seed)
    if cls == 1:
        return gen.periodic(SERIES_LEN, seed, sigma=0.05)
    return gen.utilisation(SERIES_LEN, seed)


def make_fleet(n_series, seed0, out):
    """Fills out[n_series, SERIES_LEN] with the mixed fleet (class = series index mod 3)."""
    for s in range(n_series):
        out[s, :] = make_series(s % 3, seed0 + s)


def frame_table(n_series):
    import atsc_b200
    cs = atsc_b200.chunk_sizes(SERIES_LEN)
    offs, lens = [], []
    for s in range(n_series):
        o = s * SERIES_LEN
        for c in cs:
            offs.append(o)
            lens.append(c)
            o += c
    return np.array(offs, dtype=np.uint64), np.array(lens, dtype=np.uint32)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, a few hundred Hz; falls back
    to polling nvidia-smi when the NVML binding is unavailable)."""

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
            time.sleep(0.002)

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


def cpu_baseline(threads, budget_s=20.0):
    """Oracle port timed on the host cores on a bounded sample of the same workload."""
    import oracle_lib as O
    O.lib()
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
                      f"{threads} threads, one series per thread at a time (reference is single-threaded per series)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Rust
    reference cannot be built in this image) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib as O
    O.lib()
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
    sample = f"{n} series x {SERIES_LEN} samples per step ({per_class} per class), {threads} threads"
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--series", type=int, default=288, help="series per GPU (multiple of 3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--as-rank", type=int, default=None, help="diagnostic: generate the fleet rank R of a multi-GPU run would get")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import atsc_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
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

    # ---- inputs: pinned host fleet (for e2e) + device copy (for value)
    nbytes = S * SERIES_LEN * 8
    hptr = L.atsc_gpu_host_alloc(nbytes)
    import ctypes as C
    host = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(S, SERIES_LEN))
    make_fleet(S, 5000 + (rank if args.as_rank is None else args.as_rank) * S, host)
    dev = torch.empty(S * SERIES_LEN, dtype=torch.float64, device="cuda")
    dev.copy_(torch.from_numpy(host.reshape(-1)))
    torch.cuda.synchronize()
    offs, lens = frame_table(S)
    n_samples = S * SERIES_LEN

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # page-locked result buffer: payload bytes land in it straight from the device
    pcap = 64 << 20
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
    for _ in range(2):
        step_host()
    dt_e2e, _ = timed(step_host, args.steps, "compress_e2e")
    e2e_v = world * n_samples * args.steps / dt_e2e / 1e6
    d2h = int(len(payload)) + len(lens) * 168

    # ---- roofline of the dominant kernel
    comps = np.array([out[i].compressor for i in range(len(lens))])
    nonconst = comps != atsc_b200.CONSTANT
    fft_samples = int(lens[nonconst].astype(np.int64).sum())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom = max(("stats", "poly", "rle", "fft_fwd", "fft_small", "fft", "select", "emit"), key=lambda k: kms[k])  # host_issue is not a kernel
    dom_ms = kms[dom] / args.steps
    dom_samples = n_samples if dom in ("stats", "select") else fft_samples
    achieved = dom_samples * 8 / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    # the same kernels timed without other waves sharing the GPU (one engine): explains how much of
    # the in-pipeline duration above is contention
    iso = None
    try:
        os.environ["ATSC_ENGINES"] = "1"
        ctx1 = atsc_b200.Context([local])
        for _ in range(2):
            ctx1.compress_frames(None, offs, lens, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                 samples_ptr=dev.data_ptr(), payload_out=pbuf)
        ctx1.kernel_ms(reset=True)
        for _ in range(3):
            ctx1.compress_frames(None, offs, lens, atsc_b200.AUTO, ERROR_PCT / 100.0, SPEED, True,
                                 samples_ptr=dev.data_ptr(), payload_out=pbuf)
        k1 = ctx1.kernel_ms(reset=True)
        ctx1.close()
        iso = {k: v / 3 for k, v in k1.items() if v}
    except Exception as ex:  # the figure is explanatory only
        iso = {"error": str(ex)}
    finally:
        os.environ.pop("ATSC_ENGINES", None)
    # dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed
    # ncu --set full capture (profiles/traffic.json): that launch is one 96-series wave, whose
    # algorithmic bytes are stated beside it
    traffic, traffic_detail = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic_detail = tj.get("k_" + dom)
        if traffic_detail:
            traffic = traffic_detail["bytes_per_launch"]
            traffic_detail = dict(traffic_detail, algorithmic_bytes_of_that_launch=(96 if dom == "stats" else 64) * SERIES_LEN * 8)
    except Exception:
        pass
    alg = {"stats": n_samples * 8, "poly": fft_samples * 8, "fft_fwd": fft_samples * 8}
    one_engine = None
    if iso and "error" not in iso:
        one_engine = {"k_" + k: {"ms_per_step": round(iso[k], 4), "achieved": alg[k] / (iso[k] * 1e-3) / 1e9,
                                 "frac": alg[k] / (iso[k] * 1e-3) / 1e9 / peak}
                      for k in alg if iso.get(k)}
    roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_detail": traffic_detail, "one_engine": one_engine,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_step": dom_samples * 8,
                "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items() if v and not k.startswith("reserved")},
                "kernel_ms_per_step_one_engine": iso,
                "whole_step_frac": n_samples * 8 / (dt / args.steps) / 1e9 / peak,
                "note": "achieved / frac / kernel_ms_per_step: CUDA events on each engine's stream inside the timed region "
                        "(waves of several engines overlap, so these durations include contention and add up to more "
                        "than the step); one_engine: the same kernels with a single engine, i.e. each launch alone on "
                        "the GPU, which is the figure to hold against the kernel's own roofline; k_poly and k_fft_fwd wait on "
                        "loads (ncu long-scoreboard stalls), k_stats runs near the HBM line (DESIGN.md section 4); the HBM "
                        "line is the task's stated denominator"}

    # ---- decompression of the fleet just produced (device-resident output)
    frames_in, po = [], 0
    oo = 0
    for i in range(len(lens)):
        o = out[i]
        frames_in.append((o.compressor, int(lens[i]), int(o.payload_off), int(o.payload_len), oo))
        oo += int(lens[i])
    frames_in = ctx.frames_in(frames_in)
    dout = torch.empty(n_samples, dtype=torch.float64, device="cuda")
    for _ in range(2):
        ctx.decompress_frames(frames_in, payload, out_ptr=dout.data_ptr())
    ctx.kernel_ms(reset=True)
    dt_dec, _ = timed(lambda: ctx.decompress_frames(frames_in, payload, out_ptr=dout.data_ptr()), args.steps, "decompress")
    dec_ms = ctx.kernel_ms(reset=True)["decode"] / args.steps
    hout = host.reshape(-1)  # page-locked; the input fleet is no longer needed
    dt_dec_e2e, _ = timed(lambda: ctx.decompress_frames(frames_in, payload, out=hout), max(1, args.steps // 2))
    dec = {"value": world * n_samples * 8 * args.steps / dt_dec / 1e9, "unit": "GB/s (f64 out)",
           "kernel_gbs": n_samples * 8 / (dec_ms * 1e-3) / 1e9 if dec_ms else None,
           "kernel_frac_of_hbm_peak": (n_samples * 8 / (dec_ms * 1e-3) / 1e9 / peak) if dec_ms else None,
           "e2e_gbs": world * n_samples * 8 * max(1, args.steps // 2) / dt_dec_e2e / 1e9,
           "payload_bytes": int(len(payload))}

    # ---- sanity: the timed result is a real compress (sizes, winners)
    names = atsc_b200.COMPRESSOR_NAMES
    hist = {names[c]: int((comps == c).sum()) for c in np.unique(comps)}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(os.cpu_count() or 1)
        line = {
            "metric": "auto_compress_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": dt / args.steps * 1e3,
            "device_ms_per_step": dev_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32",
            "dtype_note": "f64 for stats / polynomial / rle / error metrics, f32 for the FFT (Complex<f32> in the reference)",
            "data": "synthetic", "config": workload_config(S, world),
            "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "decompress": dec,
            "clocks": clk.summary(), "winners": hist, "compressed_bytes": int(len(payload)),
            "compression_ratio": n_samples * 8 / max(1, len(payload)),
        }
        print(json.dumps(line))
    L.atsc_gpu_host_free(hptr)
    L.atsc_gpu_host_free(pptr)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
