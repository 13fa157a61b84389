#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics (raw page) and per-source-line instruction / stall shares (source page).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18; SORT = 2 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("----")
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"{k:90s} {r[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; h = None; files = collections.defaultdict(dict)
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": h = r; continue
    if r[0] == "Function Name": continue
    if h and cur and r[0] != "":
        try:
            ln = int(r[0]); si = h.index("# Samples"); ii = h.index("Instructions Executed")
            d = files[cur].setdefault(ln, [r[1], 0, 0]); d[1] += int(r[si]); d[2] += int(r[ii])
        except Exception: pass
tot_i = sum(v[2] for f in files.values() for v in f.values()) or 1
tot_s = sum(v[1] for f in files.values() for v in f.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for f, lines in files.items():
    fi = sum(v[2] for v in lines.values()); fs = sum(v[1] for v in lines.values())
    if fi / tot_i < 0.01 and fs / tot_s < 0.01: continue
    print(f"== {f}: inst {100*fi/tot_i:.1f}%  samples {100*fs/tot_s:.1f}%")
    for ln, (srcl, s, i) in sorted(lines.items(), key=lambda kv: -kv[1][SORT])[:topn]:
        print(f"  {ln:5d} inst={100*i/tot_i:5.1f}% samp={100*s/tot_s:5.1f}%  {srcl.strip()[:100]}")
