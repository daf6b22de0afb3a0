#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (ncu -i X.ncu-rep --page source --print-source cuda,sass --csv) by the
function of sim_core.cuh each source line belongs to.  Usage: ncu_by_function.py src.csv [warp_iterations]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
iters = float(sys.argv[2]) if len(sys.argv) > 2 else 2048 * 1450
h = None; data = []
for r in rows:
    if r and r[0] == "Line No":
        h = r; ci = {}
        for i, n in enumerate(r): ci.setdefault(n, i)
        continue
    if h and len(r) == len(h) and r[0] != "":
        def gi(n):
            try: return int(r[ci[n]])
            except Exception: return 0
        data.append((int(r[0]), gi('# Samples'), gi('Instructions Executed'), gi('Thread Instructions Executed')))
src = open('librabft_simulator_b200/csrc/sim_core.cuh').read().split('\n')
func_at = {}; cur = '?'
for i, l in enumerate(src, 1):
    m = re.match(r'\s*(?:LBFT_HD|LBFT_COLD)\s+(?:static\s+)?[\w:<>\*& ]+?\s+(\w+)\s*\(', l)
    if m: cur = m.group(1)
    func_at[i] = cur
agg = {}
for ln, s, i, t in data:
    f = func_at.get(ln, '?') if ln <= len(src) else 'api'
    a = agg.setdefault(f, [0, 0, 0]); a[0] += s; a[1] += i; a[2] += t
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("warp instructions per warp-iteration ~ %.0f (source-attributed)" % (ti / iters))
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print("%-24s inst %5.1f%%  (%7.1f/iter)  stall samples %5.1f%%  active threads %5.1f" % (f, 100 * a[1] / ti, a[1] / iters, 100 * a[0] / ts, a[2] / max(a[1], 1)))
