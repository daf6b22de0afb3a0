#!/usr/bin/env python
"""Run one BASELINE configuration a few times (for ncu): python tools/profile_one.py <config> <instances> [family] [launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cid, inst = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3] in ("wide", "thread"):
    os.environ["LBFT_FORCE_KERNEL"] = sys.argv[3]
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 2
from bench import CONFIGS, make_sim, step_seeds  # noqa: E402
cfg = CONFIGS[cid]
sim = make_sim(step_seeds(cfg, 0, 0, inst), cfg["nodes"], **cfg["kw"]).create(cfg["max_clock"])
print(sim.kernel_info())
for i in range(launches):
    sim.upload()
    sim.run_device()
    print("launch %d: %.3f ms" % (i, sim.timing.sim_ms))
res = sim.download(strict=False)
print("rounds", int(res.active_rounds.sum()), "events", int(res.counters[:, :4].sum()))
