#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + per-source-line hot spots.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in KEYS:
    if k in m:
        print("%-85s %s %s" % (k, m[k][0], m[k][1]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
fname, h, ci, data = None, None, {}, []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h, ci = r, {}
        for i, n in enumerate(r):
            ci.setdefault(n, i)
        continue
    if h and len(r) == len(h) and r[0] != "":
        def gi(n):
            try:
                return int(r[ci[n]])
            except Exception:
                return 0
        data.append((fname, int(r[0]), r[1].strip()[:95], gi("# Samples"), gi("Instructions Executed"), gi("Thread Instructions Executed")))
ts, ti = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
print("\n-- top source lines by stall samples (share of samples | share of warp instructions | avg active threads)")
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%5.1f%% smp %5.1f%% inst thr %4.1f  %s:%-4d %s" % (100 * d[3] / ts, 100 * d[4] / ti, d[5] / max(d[4], 1), d[0], d[1], d[2]))
print("\n-- top source lines by warp instructions")
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print("%5.1f%% inst %5.1f%% smp thr %4.1f  %s:%-4d %s" % (100 * d[4] / ti, 100 * d[3] / ts, d[5] / max(d[4], 1), d[0], d[1], d[2]))
