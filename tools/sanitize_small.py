#!/usr/bin/env python
"""Small runs of every queue mode for compute-sanitizer (memcheck / racecheck): plain kernels, then the recording
(LBFT_FLAG_ROUND_SWITCHES) and resumable (LBFT_FLAG_RESUMABLE) instantiations with staged runs and a snapshot round
trip — the modes that write to the regions appended after the payload pool (switch table, save area, spilled queue)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librabft_simulator_b200 import BatchSimulator, RandomDelay
SEEDS = np.arange(1, 41, dtype=np.uint64)   # 40 instances: one full tile and a ragged one
DELAY = RandomDelay.new(10.0, 4.0)
MODES = (("smem scan queue", 4, 600, {}), ("hbm scan queue", 4, 600, {"queue_cap": 128}),
         ("calendar queue", 7, 400, {"partition_windows": 2, "partition_max_len": 50}),
         ("calendar queue n=33", 33, 200, {}), ("heap", 6, 4200, {}))
for name, n, max_clock, kw in MODES:
    res = BatchSimulator(SEEDS, n, DELAY, **kw).loop_until(max_clock)
    print(name, "ok", int(res.commit_counts.sum()), np.unique(res.status))
# shared-memory queue in the new modes needs a committee whose queue fits 64 entries without timer elision
NEW = (("smem scan queue", 3, 600, {"queue_cap": 64}),) + MODES[1:]
for name, n, max_clock, kw in NEW:
    with BatchSimulator(SEEDS, n, DELAY, record_round_switches=True, **kw) as sim:
        res = sim.loop_until(max_clock)
        sw = sum(len(sim.round_switches(i)) for i in (0, 31, 32, 39))
        print("recording:", name, "ok", int(res.commit_counts.sum()), np.unique(res.status), "switches", sw)   # results are lazy
    with BatchSimulator(SEEDS, n, DELAY, resumable=True, record_round_switches=True, **kw) as sim, \
            BatchSimulator(SEEDS, n, DELAY, resumable=True, record_round_switches=True, **kw) as twin:
        sim.create(max_clock)
        twin.create(max_clock)
        sim.run_until(max_clock // 3)
        twin.restore(sim.snapshot())
        sim.run_until(2 * max_clock // 3)
        a = sim.run_until(max_clock)
        twin.run_until(2 * max_clock // 3)
        b = twin.run_until(max_clock)
        same = bool((a.last_committed_states == b.last_committed_states).all()) and sim.round_switches(39) == twin.round_switches(39)
        print("resumable + recording, 3 stages + snapshot twin:", name, "ok", int(a.commit_counts.sum()), np.unique(a.status),
              "twin equal:", same)
        assert same
