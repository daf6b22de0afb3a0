"""Epoch changes (node.rs:329-348) on the GPU through the C ABI, both kernel families: bit-exact against the oracle incl. the
[10, 9, 10, 9]-type outcomes of the reference's stall, the commit logs across the epoch boundary (per-node reader and bulk
export) and the advisory status bit.  CPU-side matrix: tests/test_epochs.py."""
import numpy as np
import pytest

from tests.support import assert_same
from tests.test_gpu_parity import gpu_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cpe", [5, 10, 30])
@pytest.mark.parametrize("nodes", [3, 4, 7])
def test_gpu_epoch_changes_match_oracle(oracle, kernel_choice, nodes, cpe):
    seeds = np.arange(600, 664, dtype=np.uint64)
    o = oracle.run(seeds, nodes, 1000, commands_per_epoch=cpe)
    sim, g = gpu_run(seeds, nodes, 1000, commands_per_epoch=cpe)          # strict: an epoch change is not an error any more
    assert ((g.status & ~np.uint32(64 | 32)) == 1).all(), np.unique(g.status)
    np.testing.assert_array_equal(o.status & 32, g.status & 32)
    assert_same(o, g, "N=%d commands_per_epoch=%d (%s kernel)" % (nodes, cpe, kernel_choice))
    rows, lens = sim.commit_logs()
    np.testing.assert_array_equal(lens, o.commit_counts)
    for inst in (0, 31, 63):
        for node in range(nodes):
            want = oracle.commit_log(seeds, nodes, inst, node, 1000, commands_per_epoch=cpe)
            assert sim.commit_log(inst, node) == want
            assert [(int(r["proposer"]), int(r["index"]), int(r["time"])) for r in rows[inst, :lens[inst, node]]] == want


def test_known_outcome(oracle):
    seeds = np.arange(1, 33, dtype=np.uint64)
    sim, g = gpu_run(seeds, 4, 1000, commands_per_epoch=10)
    assert_same(oracle.run(seeds, 4, 1000, commands_per_epoch=10), g)
    assert g.commit_counts.max() <= 12 and (g.status & 32).all()          # was [34, 34, 34, 34] behind an error flag in round 1


def test_epochs_with_modes_are_refused_on_the_device():
    """The kernels with the epoch machinery are built for plain runs only (recording / resumable x epochs is covered on the
    host-compiled core, tests/test_epochs.py)."""
    from librabft_simulator_b200 import _lib
    from tests.test_gpu_parity import make_sim
    with pytest.raises(_lib.LbftError) as e:
        make_sim([1, 2], 4, commands_per_epoch=10, resumable=True).create(1000)
    assert e.value.code == -1 and "epoch" in str(e.value)
