#!/usr/bin/env python
"""Small runs of every queue mode for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librabft_simulator_b200 import BatchSimulator, RandomDelay
for name, n, max_clock, kw in (("smem scan queue", 4, 600, {}), ("hbm scan queue", 4, 600, {"queue_cap": 128}),
                               ("calendar queue", 7, 400, {"partition_windows": 2, "partition_max_len": 50}),
                               ("calendar queue n=33", 33, 200, {}), ("heap", 6, 4200, {})):
    res = BatchSimulator(np.arange(1, 41, dtype=np.uint64), n, RandomDelay.new(10.0, 4.0), **kw).loop_until(max_clock)
    print(name, "ok", int(res.commit_counts.sum()), np.unique(res.status))
