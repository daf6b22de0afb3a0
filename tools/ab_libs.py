#!/usr/bin/env python
"""A/B of alternatively built copies of the library (LBFT_LIB_PATH) and of selection overrides on the BASELINE configurations,
one GPU: python tools/ab_libs.py <label>=<lib.so>[:ENV=VAL[:ENV=VAL...]] ... -- <config>:<instances> ...
Each (variant, workload) pair runs tools/profile_one.py in its own process; prints the best of three launches and the result
digest (rounds, events) so that variants can be seen to compute the same thing."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
split = args.index("--")
variants, work = args[:split], [w.split(":") for w in args[split + 1:]]
for cid, inst in work:
    for v in variants:
        label, spec = v.split("=", 1)
        parts = spec.split(":")
        env = dict(os.environ)
        for k in ("LBFT_FORCE_KERNEL", "LBFT_THREAD_TILE", "LBFT_WIDE_GROUP", "LBFT_NO_FIXED_SHAPES", "LBFT_LIB_PATH"):
            env.pop(k, None)
        if parts[0]:
            env["LBFT_LIB_PATH"] = os.path.join(ROOT, parts[0])
        for kv in parts[1:]:
            k, val = kv.split("=")
            env[k] = val
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_one.py"), cid, inst, "auto", "3"], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        ms = [float(x) for x in re.findall(r"launch \d+: ([0-9.]+) ms", p.stdout)]
        tail = re.findall(r"rounds (\d+) events (\d+)", p.stdout)
        print("config %s x %6s | %-22s | %-58s | best %9.3f ms | %s %s" % (
            cid, inst, label, p.stdout.splitlines()[0] if p.stdout else "?", min(ms) if ms else -1, tail,
            "" if p.returncode == 0 else p.stdout[-300:]), flush=True)
