"""Regenerates BASELINE.md section 4 from the committed bench lines (profiles/r2_bench_n1.json, r2_bench_n8.json)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
j = json.loads(open(os.path.join(ROOT, 'profiles/r2_bench_n1.json')).readline())
n8 = json.loads(open(os.path.join(ROOT, 'profiles/r2_bench_n8.json')).readline())
c = j['configs']
f = lambda x, d=1: f"{x:,.{d}f}"
rows = []
for e in (1, 5, 10):
    k = c[f'C2_fft_1M_e{e}']
    rows.append((f"C2 FFT, one 1 M series, e = {e} %", "periodic + noise σ=0.5", "1", f(k['msamples_per_s']), "—",
                 f"{100*k['frac_of_hbm_line']:.2f} (latency: 11 frames, {k['ms_per_call']:.2f} ms per call)",
                 "`k_fft` fp32 FMA pipe 10 %, fp64 5 % (`r2_k_fft_noisy_full.md`)", f"{k['cpu_port_msamples_per_s']:.1f} (1)"))
k = c['C2_fft_48x1M_e5']
rows.append(("C2 FFT, 48 × 1 M series, e = 5 %", "periodic + noise σ=0.5", "1", f(k['msamples_per_s']), "—", f"{100*k['frac_of_hbm_line']:.2f}",
             "same kernel", f"{k['cpu_port_msamples_per_s']:.1f} (16)"))
k = c['C3_polynomial_3072x64k']
rows.append(("C3 Polynomial, 3072 × 64 k (subsample of 10 k × 64 k)", "monitoring mix", "1", f(k['msamples_per_s']), "—", f"{100*k['frac_of_hbm_line']:.1f}",
             "`k_poly` fp64 pipe 38 % (`r2_final_k_stats_k_poly_k_fft_fwd_full.md`)", f"{k['cpu_port_msamples_per_s']:.1f} (16)"))
k = c['C3_idw_96x64k']
rows.append(("C3 IDW, 96 × 64 k (O(N·K) per frame)", "monitoring mix", "1", f(k['msamples_per_s'], 2), "—", "≈ 0 (FP64 compute bound)",
             "FP64 (per-term divide)", f"{k['cpu_port_msamples_per_s']:.2f} (12)"))
r = j['roofline']
rows.append(("C4 auto `-c 0`, 1152 × 1 M per GPU (subsample of 100 k × 1 M)", "const / periodic / noisy gauge", "1", f(j['value']),
             f(j['decompress']['value']) + " (decompress of the same fleet)",
             f"{100*r['whole_step_frac']:.1f} whole step; dominant slot `{r['kernel']}` {100*r['frac']:.1f} in-pipeline",
             "`k_poly1s` fp64 pipe 63 % = issue port 87 % busy, `k_sfold` fp64 17 % (`r2_final_k_sfold_k_plan_k_poly1s_k_poly_full.md`, `r2_p1_variants.md`)",
             f"{j['cpu_baseline']['value']:.1f} ({j['cpu_baseline']['cores']})"))
rows.append(("C4 auto `-c 0`, 8 × 1152 × 1 M (build before `k_probe` / `k_poly1s`)", "same", "8", f(n8['value']), f(n8['decompress']['value']),
             f"{100*n8['value']*8e6/(8*6529.1e9):.1f} of 8 × the HBM line", "same kernels", "—"))
k = c['C4_auto_c6_288x1M']
rows.append(("C4 auto `-c 6`, 288 × 1 M", "same", "1", f(k['msamples_per_s']), "—", f"{100*k['frac_of_hbm_line']:.1f}",
             f"`k_fft` ({k['winners'].get('FFT', 0)} of {sum(k['winners'].values())} frames are sent to FFT by the 128-sample probe)", f"{k['cpu_port_msamples_per_s']:.1f} (12)"))
k = c['noisy_auto_c0_24x1M']
rows.append(("all-noise auto `-c 0`, 24 × 1 M", "positive white noise", "1", f(k['msamples_per_s']), "—", f"{100*k['frac_of_hbm_line']:.2f}",
             "`k_fft` 24 % issue-active, 23 iterations per frame; `k_rle` full sort", f"{k['cpu_port_msamples_per_s']:.1f} (12)"))
k = c['C5_decompress_C3_polynomial']
rows.append(("C5 decompress the C3 Polynomial fleet (3072 × 64 k)", "monitoring mix", "1", "—",
             f(k['gb_per_s_f64_out']) + f" wall, {f(k['kernel_gb_per_s'])} `k_decode` (CUDA events inside the pipelined call)",
             f"{100*k['frac_of_hbm_line']:.1f}", "`k_decode` fp64 pipe 29 % (`r2_final_k_decode_full.md`)", f"{k['cpu_port_gb_per_s']:.1f} GB/s (16)"))
d = j['decompress']
rows.append(("C5 decompress the C4 fleet", "const / periodic / noisy gauge", "1 / 8", "—", f(d['value']) + " / " + f(n8['decompress']['value']),
             f"{100*d['value']/6529.1:.0f} (wall, pipelined); kernel {100*d['kernel_frac_of_hbm_peak']:.0f} in-pipeline, 48 alone", "same", "—"))
hdr = ("| Config | Class | GPUs | Msamples/s (compress) | GB/s f64 out (decompress) | % HBM roofline (6529.1 GB/s measured) | "
       "FP64/FP32 pipe % of the dominant kernel (ncu) | CPU restated, Msamples/s (threads) |\n|---|---|---|---|---|---|---|---|\n")
tab = hdr + "".join("| " + " | ".join(r) + " |\n" for r in rows)
p = os.path.join(ROOT, 'BASELINE.md')
s = open(p).read()
i = s.index("## 4. Results table")
nt = c['noisy_auto_c0_24x1M']['near_tie_frames']
s = s[:i] + ("## 4. Results table (round 2 final build; `profiles/r2_bench_n1.json`, `profiles/r2_bench_n8.json`; regenerate with `python tools/baseline_table.py`)\n\n"
             "One B200 (and 8 for the headline), SM clock 1965 MHz, no throttle reasons.  Compress figures are device-resident (`value` of the bench line); "
             f"the end-to-end figure with host buffers is PCIe-bound at {j['e2e']['value']/1e3:.1f} Gsamples/s per GPU ({n8['e2e']['value']/1e3:.1f} on 8) for every compress shape.  "
             "The CPU column is the `-O3 -march=native` restatement (`oracle/atsc_oracle.c`) on the GPU box's host, one series per thread.  "
             f"`near_tie` frames: {j['near_tie_frames']} of {j['frames_per_step']:,} on the headline fleet, {nt} of 264 on the all-noise fleet.\n\n") + tab
open(p, 'w').write(s)
print(tab)
