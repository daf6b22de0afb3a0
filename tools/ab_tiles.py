#!/usr/bin/env python
"""Sparse warp tiles of the thread kernel (LBFT_THREAD_TILE = 8 / 4) against full tiles and the wide kernel, one GPU.
Usage: python tools/ab_tiles.py [config:instances ...]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = [w.split(":") for w in sys.argv[1:]] or [("5", "16384"), ("4", "8192")]
for cid, inst in work:
    for label, env in (("thread tile 32", {"LBFT_FORCE_KERNEL": "thread"}), ("thread tile 8", {"LBFT_FORCE_KERNEL": "thread", "LBFT_THREAD_TILE": "8"}),
                       ("thread tile 16", {"LBFT_FORCE_KERNEL": "thread", "LBFT_THREAD_TILE": "16"}), ("host's choice", {})):
        if cid == "4" and label == "thread tile 32" and int(inst) > 2048:
            continue  # 4.4 s per launch: known
        e = dict(os.environ, **env)
        for k in ("LBFT_FORCE_KERNEL", "LBFT_THREAD_TILE"):
            if k not in env:
                e.pop(k, None)
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_one.py"), cid, inst, "auto", "3"], env=e,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        ms = [float(x) for x in re.findall(r"launch \d+: ([0-9.]+) ms", p.stdout)]
        tail = re.findall(r"rounds (\d+) events (\d+)", p.stdout)
        print("config %s x %6s | %-14s | %-52s | best %9.3f ms | %s %s" % (cid, inst, label, p.stdout.splitlines()[0] if p.stdout else "?",
                                                                           min(ms) if ms else -1, tail, "" if p.returncode == 0 else p.stdout[-300:]), flush=True)
