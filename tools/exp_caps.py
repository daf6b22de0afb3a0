#!/usr/bin/env python
"""Config 4 / 5 with explicit capacities (instance footprint A/B): python tools/exp_caps.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_sim, step_seeds  # noqa: E402


def run(cid, inst, label, **extra):
    cfg = CONFIGS[cid]
    kw = dict(cfg["kw"])
    kw.update(extra)
    sim = make_sim(step_seeds(cfg, 0, 0, inst), cfg["nodes"], **kw).create(cfg["max_clock"])
    ms = []
    for i in range(3):
        sim.upload()
        sim.run_device()
        ms.append(sim.timing.sim_ms)
    res = sim.download(strict=False)
    print("config %d x %5d | %-28s | %-50s | best %9.3f ms | rounds %d events %d maxq %d maxpay %d bad %d" % (
        cid, inst, label, sim.kernel_info(), min(ms), int(res.active_rounds.sum()), int(res.counters[:, :4].sum()),
        int(res.counters[:, 8].max()), int(res.counters[:, 10].max()), int((res.status != 1).sum())), flush=True)
    sim.close()


if __name__ == "__main__":
    run(4, 8192, "default caps")
    run(4, 8192, "queue 8192 payload 128", queue_cap=8192, payload_cap=128)
    run(4, 8192, "queue 8192 payload 128 rc 64", queue_cap=8192, payload_cap=128, round_cap=64)
    run(5, 16384, "default caps")
    run(5, 16384, "queue 256 payload 32", queue_cap=256, payload_cap=32)
