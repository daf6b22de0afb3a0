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
from librabft_simulator_b200.distributed import ShardedBatchSimulator
from tests.support import HostCore
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
hc = HostCore()

class HostLocal:
    # stand-in for this rank's BatchSimulator (same create / set_seeds / run surface): the host-compiled device core
    def __init__(self, shard): self.shard = shard
    def create(self, max_clock): self.max_clock = max_clock; return self
    def close(self): pass
    def set_seeds(self, shard): self.shard = shard
    def run(self, strict=True):
        r = hc.run(self.shard, 4, self.max_clock)
        r.last_committed_states, r.active_rounds = r.last_states, np.ascontiguousarray(r.counters[:, 6])
        return r

seeds = np.arange(300, 364, dtype=np.uint64)
sim = ShardedBatchSimulator(seeds, 4, rank=rank, world=world, dist=dist, make_local=HostLocal)
res = sim.loop_until(1000)
assert (res.lo, res.hi) == (32 * rank, 32 * rank + 32)
np.save(os.path.join(%(out)r, "counts_%%d.npy" %% rank), res.commit_counts)
np.save(os.path.join(%(out)r, "states_%%d.npy" %% rank), res.last_committed_states)
np.save(os.path.join(%(out)r, "rounds_%%d.npy" %% rank), res.active_rounds)
sim.set_seeds(seeds + 1000)          # re-seed the whole job: every rank keeps its own contiguous shard
res2 = sim.run()
np.save(os.path.join(%(out)r, "counts2_%%d.npy" %% rank), res2.commit_counts)
streamed = list(sim.run_stream([seeds, seeds + 1000]))   # (this stand-in has no run_async: one run after the other)
assert len(streamed) == 2
assert (streamed[0].commit_counts == res.commit_counts).all() and (streamed[1].commit_counts == res2.commit_counts).all()
assert (streamed[1].last_committed_states == res2.last_committed_states).all()
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
    ref2 = oracle.run(seeds + 1000, 4, 1000)
    for rank in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / ("counts_%d.npy" % rank)), ref.commit_counts)
        np.testing.assert_array_equal(np.load(tmp_path / ("states_%d.npy" % rank)), ref.last_states)
        np.testing.assert_array_equal(np.load(tmp_path / ("rounds_%d.npy" % rank)), ref.counters[:, 6])
        np.testing.assert_array_equal(np.load(tmp_path / ("counts2_%d.npy" % rank)), ref2.commit_counts)


def test_functional_form_and_uneven_batches():
    from librabft_simulator_b200.distributed import ShardedBatchSimulator, run_sharded
    import pytest
    with pytest.raises(ValueError, match="multiple of the world size"):
        ShardedBatchSimulator(np.arange(7), 4, rank=0, world=2, make_local=lambda s: None)
    seeds = np.arange(10, 18, dtype=np.uint64)
    counts, states = run_sharded(seeds, 3, 1000, 0, 1, lambda s: (np.tile(s[:, None], (1, 3)), np.tile(s[:, None] * 2, (1, 3))))
    assert counts.shape == (8, 3) and (counts[:, 0] == seeds).all() and (states[:, 2] == seeds * 2).all()


class _AsyncHostLocal:
    """A local runner with the asynchronous surface of BatchSimulator (set_seeds / run_async / wait(relaunch, before_relaunch)),
    the host-compiled core doing the work at wait(): checks the ORDER ShardedBatchSimulator.run_stream drives it in."""

    def __init__(self, hc, log):
        self.hc, self.log, self.staged, self.inflight = hc, log, None, None

    def create(self, max_clock):
        self.max_clock = max_clock
        return self

    def close(self):
        pass

    def set_seeds(self, shard):
        self.log.append("stage")
        self.staged = np.array(shard)

    def run_async(self):
        assert self.inflight is None, "two runs in flight"
        self.log.append("launch")
        self.inflight = self.staged

    def wait(self, strict=True, relaunch=False, before_relaunch=None):
        self.log.append("wait")
        r = self.hc.run(self.inflight, 4, self.max_clock)
        self.inflight = None
        r.last_committed_states, r.active_rounds = r.last_states, np.ascontiguousarray(r.counters[:, 6])
        if before_relaunch is not None:
            before_relaunch()
        if relaunch:
            self.run_async()
        return r

    def drain(self):
        self.log.append("drain")
        self.inflight = None


def test_run_stream_stages_the_next_batch_before_waiting(oracle, hostcore):
    from librabft_simulator_b200.distributed import ShardedBatchSimulator
    log = []
    batches = [np.arange(b, b + 16, dtype=np.uint64) for b in (10, 900, 4242)]
    sim = ShardedBatchSimulator(batches[0], 4, make_local=lambda shard: _AsyncHostLocal(hostcore, log)).create(500)
    outs = list(sim.run_stream(iter(batches)))
    assert log == ["stage", "launch", "stage", "wait", "launch", "stage", "wait", "launch", "wait"]
    for res, b in zip(outs, batches):
        ref = oracle.run(b, 4, 500)
        np.testing.assert_array_equal(res.commit_counts, ref.commit_counts)
        np.testing.assert_array_equal(res.last_committed_states, ref.last_states)
    assert list(sim.run_stream([])) == []
    # a consumer that stops early leaves no run in flight
    del log[:]
    stream = sim.run_stream(iter(batches))
    next(stream)
    stream.close()
    assert log == ["stage", "launch", "stage", "wait", "launch", "drain"] and sim.local.inflight is None
