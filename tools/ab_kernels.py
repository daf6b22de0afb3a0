#!/usr/bin/env python
"""Thread-per-instance vs warp-per-instance kernel on the BASELINE configurations and on a sweep of batch sizes
(LBFT_FORCE_KERNEL A/B on one GPU): kernel ms (CUDA events, seeds resident), best of `reps` launches, and a parity bit
(both families must return identical results).  Usage: python tools/ab_kernels.py [quick]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_sim, step_seeds  # noqa: E402


def time_one(cfg, instances, family, reps=3):
    os.environ["LBFT_FORCE_KERNEL"] = family
    seeds = step_seeds(cfg, 0, 0, instances)
    sim = make_sim(seeds, cfg["nodes"], **cfg["kw"]).create(cfg["max_clock"])
    name = sim.kernel_info()
    best, res = 1e30, None
    for r in range(reps):
        sim.upload()
        sim.run_device()
        best = min(best, float(sim.timing.sim_ms))
    res = sim.download(strict=False)
    out = (best, name, res.commit_counts.copy(), res.last_committed_states.copy(), res.counters.copy(), np.unique(res.status).tolist())
    sim.close()
    return out


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    plan = [(2, 1024), (5, 16384), (4, 8192 if not quick else 1024), (1, 1)]
    plan += [(3, n) for n in ((1024, 8192, 16384, 32768, 65536) if not quick else (1024, 8192))]
    plan += [(5, n) for n in ((2048, 65536) if not quick else ())]
    for cid, inst in plan:
        cfg = CONFIGS[cid]
        row = {}
        for fam in ("thread", "wide"):
            if cid == 4 and fam == "thread" and inst > 2048:
                t = time_one(cfg, inst, fam, reps=1)
            else:
                t = time_one(cfg, inst, fam)
            row[fam] = t
        same = all((row["thread"][i] == row["wide"][i]).all() for i in (2, 3)) and (row["thread"][4][:, :8] == row["wide"][4][:, :8]).all()
        rounds = float(row["wide"][4][:, 6].sum())
        print("config %d x %6d inst | thread %9.3f ms (%s) | wide %9.3f ms (%s) | wide/thread %.2f | %.2f Mrounds/s best | same=%s status %s/%s maxq %d" % (
            cid, inst, row["thread"][0], row["thread"][1], row["wide"][0], row["wide"][1], row["wide"][0] / row["thread"][0],
            rounds / min(row["thread"][0], row["wide"][0]) / 1e3, same, row["thread"][5], row["wide"][5], row["wide"][4][:, 8].max()), flush=True)


if __name__ == "__main__":
    main()
