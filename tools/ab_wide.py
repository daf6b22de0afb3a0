#!/usr/bin/env python
"""A/B of wide-kernel variants on one GPU: for each (library build, lanes per instance, shared-memory state) combination run
tools/profile_one.py in a subprocess and collect the best launch time.  Usage: python tools/ab_wide.py [config:instances ...]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "librabft_simulator_b200", "csrc")
work = [w.split(":") for w in sys.argv[1:]] or [("2", "1024"), ("5", "16384"), ("4", "2048"), ("1", "1")]
libs = [("r128", "liblbft_b200.so")]
for cid, inst in work:
    for lname, lib in libs:
        if not os.path.exists(os.path.join(CSRC, lib)):
            continue
        for group in ("8", "32"):
            for smem in ("0", "1"):
                env = dict(os.environ, LBFT_LIB_PATH=os.path.join(CSRC, lib), LBFT_WIDE_GROUP=group, LBFT_WIDE_SMEM=smem, LBFT_FORCE_KERNEL="wide")
                p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_one.py"), cid, inst, "wide", "3"], env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                ms = [float(x) for x in re.findall(r"launch \d+: ([0-9.]+) ms", p.stdout)]
                name = p.stdout.splitlines()[0] if p.stdout else "?"
                if smem == "1" and "true" not in name:
                    continue  # the host declined shared-memory state for this shape: same as smem=0
                print("config %s x %6s | %-5s group %2s smem %s | %-34s | best %9.3f ms | %s" % (
                    cid, inst, lname, group, smem, name, min(ms) if ms else -1, "" if p.returncode == 0 else p.stdout[-300:]), flush=True)
