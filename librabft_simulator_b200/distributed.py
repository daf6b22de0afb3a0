"""Instance-level data parallelism: simulator instances are independent (own seed, own RNG,
own nodes — simulator.rs:200-250), so the batch is sharded contiguously over ranks with NO data-path
collective; one all-gather of the per-instance summaries (commit counts, state keys) happens at the end.
One process per GPU, ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in CPU tests)."""
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


def run_sharded(seeds, num_nodes, max_clock, rank, world, run_local, dist=None, device=None):
    """Run this rank's contiguous shard with ``run_local(seeds_shard) -> (commit_counts[I_r,N], last_states[I_r,N])``
    and all-gather the summaries so that every rank holds the whole batch's results.

    Shards must be equal-sized for the single ``all_gather_into_tensor`` (pad the batch to a multiple of the
    world size otherwise)."""
    import torch
    seeds = np.asarray(seeds, dtype=np.uint64)
    total = len(seeds)
    if total % world != 0:
        raise ValueError("number of instances (%d) must be a multiple of the world size (%d)" % (total, world))
    lo, hi = shard_bounds(total, world, rank)
    counts, states = run_local(seeds[lo:hi])
    counts_t = torch.as_tensor(np.ascontiguousarray(counts).astype(np.int32))
    states_t = torch.as_tensor(np.ascontiguousarray(states).view(np.int64))
    if device is not None:
        counts_t, states_t = counts_t.to(device), states_t.to(device)
    counts_all = all_gather_rows(counts_t, world, dist)
    states_all = all_gather_rows(states_t, world, dist)
    return (counts_all.cpu().numpy().astype(np.uint32).reshape(total, num_nodes),
            states_all.cpu().numpy().view(np.uint64).reshape(total, num_nodes))
