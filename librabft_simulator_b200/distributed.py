"""Instance-level data parallelism: simulator instances are independent (own seed, own RNG,
own nodes — simulator.rs:200-250), so the batch is sharded contiguously over ranks with NO data-path
collective; one all-gather of the per-instance summaries ``{commit_count[N], state_key[N], rounds}`` happens at
the end of a run.  One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch on the B200 box, gloo in
the CPU tests).

``ShardedBatchSimulator`` is the product's multi-GPU entry: it owns the sharding (``shard_bounds``), this rank's
``BatchSimulator`` handle and the all-gather.  With NCCL the gather reads the library's device-resident result
buffers in place (``lbft_device_buffer``): no host round trip.  (One host thread driving several GPUs without
``torch.distributed`` is the other route: ``lbft_run_async`` / ``lbft_wait`` on one handle per device, see
``tests/cabi/multi_handle.c`` and ``bft-lib-gpu``.)"""
import numpy as np


def shard_bounds(total, world, rank):
    """GPU g gets instances [g*I/G, (g+1)*I/G)."""
    return (rank * total) // world, ((rank + 1) * total) // world


def all_gather_rows(local, world, dist=None):
    """All-gather equally-shaped per-rank tensors along dim 0 (device tensors for NCCL, CPU for gloo)."""
    import torch
    if world == 1 or dist is None:
        return local
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class _DeviceView:
    """``__cuda_array_interface__`` over a library-owned device buffer (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedResult:
    """Whole-job summaries after the all-gather, plus this rank's own ``BatchResult`` (``.local``)."""

    def __init__(self, local, commit_counts, last_committed_states, active_rounds, lo, hi):
        self.local, self.lo, self.hi = local, lo, hi
        self._cc, self._ls, self._ar = commit_counts, last_committed_states, active_rounds

    @staticmethod
    def _host(t, dtype):
        return t.detach().cpu().numpy().view(dtype) if hasattr(t, "detach") else np.asarray(t).view(dtype)

    @property
    def commit_counts(self):
        """[all instances, node] ``committed_history().len()``"""
        return self._host(self._cc, np.uint32)

    @property
    def last_committed_states(self):
        return self._host(self._ls, np.uint64)

    @property
    def active_rounds(self):
        return self._host(self._ar, np.uint32)


class ShardedBatchSimulator:
    """The batch of ``seeds`` (the WHOLE job's, identical on every rank) spread over ``world`` ranks.

    ``make_local(seeds_shard)`` builds this rank's simulator (default: a ``BatchSimulator`` on ``device``); the CPU
    tests inject a host stand-in with the same ``create / set_seeds / run`` surface, so the sharding and gathering
    logic is covered without a GPU."""

    def __init__(self, seeds, num_nodes, *args, rank=0, world=1, dist=None, device=0, make_local=None, **kw):
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(-1))
        if len(seeds) % world != 0:
            raise ValueError("number of instances (%d) must be a multiple of the world size (%d): the summaries travel in one "
                             "all_gather_into_tensor" % (len(seeds), world))
        self.total, self.num_nodes, self.rank, self.world, self.dist, self.device = len(seeds), int(num_nodes), rank, world, dist, device
        self.lo, self.hi = shard_bounds(self.total, world, rank)
        if make_local is None:
            from .simulator import BatchSimulator

            def make_local(shard):
                return BatchSimulator(shard, num_nodes, *args, device=device, **kw)
        self.local = make_local(seeds[self.lo:self.hi])
        self._views = None

    def create(self, max_clock):
        self.local.create(int(max_clock))
        self._views = None
        return self

    def close(self):
        self.local.close()

    def set_seeds(self, seeds):
        """Re-seed the whole job; this rank keeps its contiguous shard."""
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).reshape(-1))
        if len(seeds) != self.total:
            raise ValueError("expected %d seeds" % self.total)
        self.local.set_seeds(seeds[self.lo:self.hi])

    # -- the one collective of the path ---------------------------------------------------------------
    def _device_views(self):
        """Torch views of the library's device result buffers (NCCL gathers them in place)."""
        import torch
        if self._views is None:
            I, N = self.hi - self.lo, self.num_nodes
            dev = "cuda:%d" % self.device
            cc, _ = self.local.device_buffer(0)
            ls, _ = self.local.device_buffer(1)
            ar, _ = self.local.device_buffer(4)
            self._views = (torch.as_tensor(_DeviceView(cc, (I, N), "<i4"), device=dev),
                           torch.as_tensor(_DeviceView(ls, (I, N), "<i8"), device=dev),
                           torch.as_tensor(_DeviceView(ar, (I,), "<i4"), device=dev))
        return self._views

    def gather(self, local_result=None):
        """All-gather ``{commit_count[N], state_key[N], rounds}`` of every rank's shard.  On a CUDA backend the
        sources are the device-resident result buffers; otherwise the host arrays of ``local_result``."""
        import torch
        on_device = self.dist is not None and self.world > 1 and self.dist.get_backend() == "nccl"
        if on_device:
            cc, ls, ar = self._device_views()
        else:
            cc = torch.as_tensor(np.ascontiguousarray(local_result.commit_counts).view(np.int32))
            ls = torch.as_tensor(np.ascontiguousarray(local_result.last_committed_states).view(np.int64))
            ar = torch.as_tensor(np.ascontiguousarray(local_result.active_rounds).view(np.int32))
        d = self.dist if self.world > 1 else None
        return all_gather_rows(cc, self.world, d), all_gather_rows(ls, self.world, d), all_gather_rows(ar, self.world, d)

    def run(self, strict=True):
        """``lbft_run`` on this rank's shard (host seeds in, host summaries out) + the all-gather."""
        res = self.local.run(strict=strict)
        cc, ls, ar = self.gather(res)
        return ShardedResult(res, cc, ls, ar, self.lo, self.hi)

    def loop_until(self, max_clock, strict=True):
        """``Simulator::new`` + ``loop_until(max_clock)`` for the whole job (simulator.rs:200-250, 380-475)."""
        self.create(max_clock)
        return self.run(strict=strict)


def run_sharded(seeds, num_nodes, max_clock, rank, world, run_local, dist=None, device=None):
    """Functional form kept for callers that bring their own local runner: ``run_local(seeds_shard) -> (commit_counts[I_r,N],
    last_states[I_r,N])``; returns the whole job's two arrays on every rank."""
    class _Local:
        def __init__(self, shard):
            self.shard = shard

        def create(self, max_clock):
            return self

        def close(self):
            pass

        def set_seeds(self, shard):
            self.shard = shard

        def run(self, strict=True):
            counts, states = run_local(self.shard)
            return type("R", (), {"commit_counts": np.asarray(counts, dtype=np.uint32), "last_committed_states": np.asarray(states, dtype=np.uint64),
                                  "active_rounds": np.zeros(len(self.shard), dtype=np.uint32)})()

    sim = ShardedBatchSimulator(seeds, num_nodes, rank=rank, world=world, dist=dist, make_local=_Local)
    res = sim.loop_until(max_clock)
    total = len(np.asarray(seeds).reshape(-1))
    return res.commit_counts.reshape(total, num_nodes), res.last_committed_states.reshape(total, num_nodes)
