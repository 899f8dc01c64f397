"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + top stall lines by source.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
for v in vals:
    print("## kernel:", v[hdr.index("Kernel Name")][:60])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {v[i]} | {units[i]} |")
    # stall breakdown
    st = [(float(v[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[i]]
    st.sort(reverse=True)
    print("\nstalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={x:.2f}" for x, h in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
h = rows[hi]
si = h.index("Source")
ci = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
ii = next((i for i, c in enumerate(h) if c.startswith("Instructions Executed")), None)
agg = collections.Counter(); ins = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= max(ci, ii or 0): continue  # rows of the next kernel's header block are shorter
    try:
        agg[r[si].strip()] += float(r[ci] or 0)
        if ii is not None: ins[r[si].strip()] += float(r[ii] or 0)
    except ValueError:
        pass
tot = sum(agg.values()) or 1
print(f"\n## warp-stall samples by source line (top {top} of {int(tot)} samples)\n| samples | share | inst | source |\n|---|---|---|---|")
for s, c in agg.most_common(top):
    print(f"| {int(c)} | {100*c/tot:.1f}% | {int(ins[s])} | `{s[:110]}` |")
