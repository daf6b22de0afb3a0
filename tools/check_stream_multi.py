#!/usr/bin/env python
"""Under torchrun (one rank per GPU, NCCL): ShardedBatchSimulator.run_stream over a few seed batches, every gathered result
compared with the oracle on every rank (the all-gather of a finished run reads the device buffers before the next run is
launched).  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_stream_multi.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from librabft_simulator_b200 import RandomDelay, ShardedBatchSimulator  # noqa: E402
from tests.support import Oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = 512 * world
batches = [np.arange(b, b + total, dtype=np.uint64) for b in (1, 100000, 7777777, 31337)]
sim = ShardedBatchSimulator(batches[0], 4, RandomDelay.new(10.0, 4.0), rank=rank, world=world, dist=dist, device=local).create(1000)
oracle = Oracle()
n = 0
for res, b in zip(sim.run_stream(batches), batches):
    ref = oracle.run(b, 4, 1000)
    assert (res.commit_counts == ref.commit_counts).all(), "commit counts differ (batch %d, rank %d)" % (n, rank)
    assert (res.last_committed_states == ref.last_states).all(), "state keys differ (batch %d, rank %d)" % (n, rank)
    assert (res.active_rounds == ref.counters[:, 6]).all(), "rounds differ (batch %d, rank %d)" % (n, rank)
    n += 1
assert n == len(batches)
one = sim.run()                       # and the plain path after a stream
assert (one.commit_counts == oracle.run(batches[-1], 4, 1000).commit_counts).all()
dist.barrier()
if rank == 0:
    print("run_stream over %d ranks: %d batches x %d instances bit-exact vs the oracle on every rank" % (world, n, total))
dist.destroy_process_group()
