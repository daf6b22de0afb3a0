#!/usr/bin/env python
"""Cost of the resumable mode on the bench workload (65 536 instances x 4 authors, horizon 1000): one-shot plain run,
one-shot run of a resumable handle, and a 3-stage lbft_run_until sequence; plus snapshot size / save / load times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librabft_simulator_b200 import BatchSimulator, RandomDelay  # noqa: E402

seeds = np.arange(52, 52 + 65536, dtype=np.uint64)
delay = RandomDelay.new(10.0, 4.0)
with BatchSimulator(seeds, 4, delay) as plain:
    plain.create(1000); plain.run(); plain.run()
    print("plain one-shot                     kernel %7.2f ms" % plain.timing.sim_ms)
with BatchSimulator(seeds, 4, delay, resumable=True) as sim:
    sim.create(1000); sim.run(); sim.run()
    dev_bytes, words = sim.memory_info()
    print("resumable handle, one-shot lbft_run kernel %7.2f ms   state %.1f KB/inst" % (sim.timing.sim_ms, words * 4 / 1024))
    for rep in range(2):
        sim.set_seeds(seeds)
        parts = []
        for stop in (300, 650, 1000):
            sim.run_until(stop)
            parts.append(sim.timing.sim_ms)
    print("resumable, run_until 300/650/1000  kernel %s = %7.2f ms" % (" + ".join("%.2f" % p for p in parts), sum(parts)))
    sim.set_seeds(seeds); sim.run_until(500)
    t0 = time.perf_counter(); snap = sim.snapshot(); t1 = time.perf_counter(); sim.restore(snap); t2 = time.perf_counter()
    print("snapshot %.1f MB   save %.1f ms   load %.1f ms (pageable host memory)" % (snap.nbytes / 1e6, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
