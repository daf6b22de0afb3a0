#!/usr/bin/env python
"""Generates tests/golden/commit_logs.json with the CPU oracle (oracle/, pinned by the reference's own golden
values: librabft-v2/tests/simulated_run.rs:45-94).  The reference itself is Rust and cannot be run in this image,
so these fixtures are oracle outputs: full commit logs (proposer, index, time), state keys and event counters for
the two reference golden runs and for a few small batches.  Re-run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.support import Oracle  # noqa: E402

CASES = [
    {"name": "reference golden: seed 52, 3 nodes (simulated_run.rs:45-66)", "seed0": 52, "count": 1, "nodes": 3, "max_clock": 1000, "kw": {}},
    {"name": "reference golden: seed 48, 8 nodes (simulated_run.rs:68-94)", "seed0": 48, "count": 1, "nodes": 8, "max_clock": 1000, "kw": {}},
    {"name": "config 3 shape: 8 instances x 4 nodes", "seed0": 52, "count": 8, "nodes": 4, "max_clock": 1000, "kw": {}},
    {"name": "config 1: fixed 10 ms delay, 3 nodes, > 100 rounds", "seed0": 52, "count": 1, "nodes": 3, "max_clock": 3000,
     "kw": {"delay_variance": 0.0}},
    {"name": "config 5 shape: 4 instances x 7 nodes, partitions", "seed0": 1000, "count": 4, "nodes": 7, "max_clock": 1000,
     "kw": {"partition_windows": 4, "partition_max_len": 150}},
    {"name": "uniform delay [5,15], 4 nodes", "seed0": 7, "count": 4, "nodes": 4, "max_clock": 1000,
     "kw": {"delay_kind": 1, "delay_lo": 5, "delay_hi": 15}},
]


def main():
    oracle = Oracle()
    out = []
    for c in CASES:
        seeds = np.arange(c["seed0"], c["seed0"] + c["count"], dtype=np.uint64)
        r = oracle.run(seeds, c["nodes"], c["max_clock"], **c["kw"])
        logs = [[oracle.commit_log(seeds, c["nodes"], i, n, c["max_clock"], **c["kw"]) for n in range(c["nodes"])]
                for i in range(c["count"])]
        out.append(dict(c, commit_counts=r.commit_counts.tolist(), last_states=[[str(x) for x in row] for row in r.last_states.tolist()],
                        counters=r.counters[:, :8].tolist(), logs=logs))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "commit_logs.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
