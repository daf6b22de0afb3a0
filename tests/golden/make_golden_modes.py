#!/usr/bin/env python
"""Generates tests/golden/modes.json with the CPU oracle: DataWriter round switches (data_writer.rs:34-50) and staged
runs (loop_until called once per stop, simulator.rs:380) for the two reference golden seeds and two small batches.
ORACLE OUTPUTS, NOT REFERENCE OUTPUTS: the reference has no golden for either (no test uses DataWriter or calls
loop_until twice) and cannot be run in this image; the fixture exists so that the oracle and the device cannot drift
together unnoticed.  Re-run:  python tests/golden/make_golden_modes.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.support import Oracle  # noqa: E402

CASES = [
    {"name": "seed 52, 3 nodes (simulated_run.rs:45-66)", "seed0": 52, "count": 1, "nodes": 3, "max_clock": 1000, "kw": {}},
    {"name": "seed 48, 8 nodes (simulated_run.rs:68-94)", "seed0": 48, "count": 1, "nodes": 8, "max_clock": 1000, "kw": {}},
    {"name": "3 instances x 4 nodes", "seed0": 52, "count": 3, "nodes": 4, "max_clock": 1000, "kw": {}},
    {"name": "2 instances x 7 nodes, partitions", "seed0": 1000, "count": 2, "nodes": 7, "max_clock": 1000,
     "kw": {"partition_windows": 4, "partition_max_len": 150}},
]
SCHEDULES = [[500, 1000], [300, 650, 1000], [17, 400, 399, 1000]]


def main():
    o = Oracle()
    out = []
    for c in CASES:
        seeds = np.arange(c["seed0"], c["seed0"] + c["count"], dtype=np.uint64)
        N, mc, kw = c["nodes"], c["max_clock"], c["kw"]
        one = o.run(seeds, N, mc, **kw)
        entry = dict(c, message_count=[int(x) for x in one.counters[:, :3].sum(axis=1)],
                     round_switches=[o.round_switches(seeds, N, i, mc, **kw) for i in range(c["count"])], staged=[])
        for stops in SCHEDULES:
            r = o.run_staged(seeds, N, stops, mc, **kw)
            entry["staged"].append({"stops": stops, "commit_counts": r.commit_counts.tolist(),
                                    "last_states": [[str(x) for x in row] for row in r.last_states.tolist()],
                                    "counters": r.counters[:, :8].tolist(),
                                    "round_switches": [o.round_switches_staged(seeds, N, i, stops, mc, **kw) for i in range(c["count"])]})
        out.append(entry)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "modes.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
