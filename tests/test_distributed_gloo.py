"""The N>1 path on CPU: world_size-2 gloo processes shard the batch, run their shard (the host-compiled device
core stands in for the GPU kernel — test infrastructure), all-gather the summaries; every rank must end up with
exactly the single-process result."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from librabft_simulator_b200.distributed import run_sharded, shard_bounds
from tests.support import HostCore
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
hc = HostCore()
seeds = np.arange(300, 364, dtype=np.uint64)
def run_local(s):
    r = hc.run(s, 4, 1000)
    return r.commit_counts, r.last_states
counts, states = run_sharded(seeds, 4, 1000, rank, world, run_local, dist=dist)
np.save(os.path.join(%(out)r, "counts_%%d.npy" %% rank), counts)
np.save(os.path.join(%(out)r, "states_%%d.npy" %% rank), states)
dist.barrier()
dist.destroy_process_group()
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_the_batch():
    from librabft_simulator_b200.distributed import shard_bounds
    for total in (1, 7, 64, 65536):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_two_rank_gloo_matches_single_process(tmp_path, oracle):
    port = free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out.decode()[-2000:]
    seeds = np.arange(300, 364, dtype=np.uint64)
    ref = oracle.run(seeds, 4, 1000)
    for rank in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / ("counts_%d.npy" % rank)), ref.commit_counts)
        np.testing.assert_array_equal(np.load(tmp_path / ("states_%d.npy" % rank)), ref.last_states)
