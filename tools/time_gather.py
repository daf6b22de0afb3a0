#!/usr/bin/env python
"""Where does a multi-GPU e2e step spend its time?  torchrun --nproc-per-node 2 tools/time_gather.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_sim, step_seeds  # noqa: E402
from librabft_simulator_b200 import ShardedBatchSimulator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = CONFIGS[3]
total = cfg["instances"] * world
sh = ShardedBatchSimulator(step_seeds(cfg, 0, 0, total), cfg["nodes"], rank=rank, world=world, dist=dist, device=local,
                           make_local=lambda s: make_sim(s, cfg["nodes"], device=local, **cfg["kw"]))
sh.create(cfg["max_clock"])
for w in range(3):
    sh.run(strict=False)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
T = {"set_seeds": 0, "local_run": 0, "gather_enqueue": 0, "gather_sync": 0}
for s in range(5):
    t0 = time.perf_counter(); sh.set_seeds(step_seeds(cfg, s, 0, total)); t1 = time.perf_counter()
    res = sh.local.run(strict=False); t2 = time.perf_counter()
    g = sh.gather(res); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
        T[k] += v * 1e3 / 5
print("rank %d ms per step: %s  (kernel %.2f ms)" % (rank, {k: round(v, 3) for k, v in T.items()}, sh.local.timing.sim_ms), flush=True)
dist.barrier(); dist.destroy_process_group()
