#!/usr/bin/env python
"""Every source line of an .ncu-rep with its share of warp instructions, in source order (read here, no GPU needed).
Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_share_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
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
        data.append((fname, int(r[0]), r[1].strip()[:110], gi("# Samples"), gi("Instructions Executed"), gi("Thread Instructions Executed")))
ts, ti = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
print("total warp instructions %d, samples %d" % (ti, ts))
cum = 0.0
for d in data:
    share = 100.0 * d[4] / ti
    if share < min_share:
        continue
    cum += share
    print("%5.2f%% inst %5.2f%% smp thr %4.1f  %s:%-4d %s" % (share, 100.0 * d[3] / ts, d[5] / max(d[4], 1), d[0], d[1], d[2]))
print("listed %.1f%%" % cum)
