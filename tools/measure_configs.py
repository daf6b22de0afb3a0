#!/usr/bin/env python
"""Time BASELINE.json configs 2, 4 and 5 (parity-test cases, not bench lines) on one GPU and check a sample of
each against the oracle.  Usage: python tools/measure_configs.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librabft_simulator_b200 import BatchSimulator, RandomDelay  # noqa: E402
from tests.support import Oracle  # noqa: E402

W64 = [1 + (i % 3) for i in range(64)]
SILENT64 = [1 if i % 3 == 0 and i <= 60 else 0 for i in range(64)]
CONFIGS = [
    ("config2 1024x4 LogNormal(10,4)", 1024, 4, RandomDelay.new(10.0, 4.0), {}, {}),
    ("config2 1024x4 uniform[5,15]", 1024, 4, RandomDelay.uniform(5, 15), {}, {"delay_kind": 1, "delay_lo": 5, "delay_hi": 15}),
    ("config3 65536x4 LogNormal(10,4)", 65536, 4, RandomDelay.new(10.0, 4.0), {}, {}),
    ("config3 65536x4 + round-switch recording (flag)", 65536, 4, RandomDelay.new(10.0, 4.0), {"record_round_switches": True}, {}),
    ("config5 16384x7 partitions(4 windows <=150ms)", 16384, 7, RandomDelay.new(10.0, 4.0),
     {"partition_windows": 4, "partition_max_len": 150}, {"partition_windows": 4, "partition_max_len": 150}),
    ("config4 8192x64 weighted, 21 silent", 8192, 64, RandomDelay.new(10.0, 4.0),
     {"voting_rights": W64, "silent": SILENT64}, {"voting_rights": W64, "silent": SILENT64}),
]
oracle = Oracle()
for name, I, N, delay, kw, okw in CONFIGS:
    seeds = np.arange(52, 52 + I, dtype=np.uint64)
    sim = BatchSimulator(seeds, N, delay, **kw).create(1000)
    dev_bytes, words = sim.memory_info()
    sim.run()
    t0 = time.perf_counter()
    res = sim.run()
    wall = time.perf_counter() - t0
    ms = sim.timing.sim_ms
    rounds, events = float(res.active_rounds.sum()), float(res.events_processed.sum())
    sample = [0, I // 2, I - 1]
    ref = oracle.run(seeds[sample], N, 1000, **okw)
    ok = (ref.last_states == res.last_committed_states[sample]).all() and (ref.counters[:, :8] == res.counters[sample, :8]).all()
    if kw.get("record_round_switches"):  # the recorded switches of the sampled instances, too
        ok = ok and all(sim.round_switches(i) == oracle.round_switches(seeds[[i]], N, 0, 1000, **okw) for i in sample)
    print("%-48s kernel %9.2f ms  e2e %9.2f ms  %8.2f Mrounds/s  %7.3f Gev/s  state %6.1f KB/inst (%5.2f GB)  status %s  "
          "maxq %d maxpay %d  oracle-sample %s" % (name, ms, wall * 1e3, rounds / ms / 1e3, events / ms / 1e6, words * 4 / 1024,
                                                   dev_bytes / 1e9, np.unique(res.status).tolist(), res.counters[:, 8].max(),
                                                   res.counters[:, 10].max(), "OK" if ok else "MISMATCH"))
    sim.close()
