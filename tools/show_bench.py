#!/usr/bin/env python
"""One-screen digest of a bench.py JSON line: python tools/show_bench.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
p = d.get("parity", {})
print("value %.1fM rounds/s  e2e %.1fM  %.3f ms/step  alg frac %.3f  parity %s/%s ok=%s  cpu %.3fM (%s cores)" % (
    d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"], p.get("checked"), p.get("of"), p.get("ok"),
    d.get("cpu_baseline", {}).get("value", 0) / 1e6, d.get("cpu_baseline", {}).get("cores")))
print("clocks", d.get("clocks"), "flagged", d.get("flagged_instances"))
for k, c in sorted(d.get("configs", {}).items()):
    print("config %s %-44s %9.2f ms  value %8.3fM  e2e %8.3fM  ev/s %6.2fG  alg frac %.3f  parity %d/%d ok=%s  cpu %.4fM  e2e/cpu %.1f" % (
        k, c["kernel"], c["kernel_ms"], c["value"] / 1e6, c["e2e"] / 1e6, c["events_per_s"] / 1e9, c["roofline"]["frac"],
        c["parity"]["checked"], c["parity"]["of"], c["parity"]["ok"], c["cpu_baseline"]["value"] / 1e6, c["e2e_over_cpu"]))
