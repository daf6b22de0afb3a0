"""LBFT_FLAG_TRUE_DATA_SYNC on the GPU through the C ABI (see tests/test_true_data_sync.py for what the variant is): the
kernel instantiations with the request / response payloads against the oracle running the same variant, all four queue modes."""
import numpy as np
import pytest

from tests.support import assert_same
from tests.test_gpu_parity import gpu_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nodes,count,max_clock,kw", [
    (4, 512, 1000, {}), (3, 96, 1000, {}), (5, 64, 1000, {}), (7, 64, 1000, {"partition_windows": 3, "partition_max_len": 150}),
    (16, 8, 600, {}), (6, 8, 4200, {}), (4, 40, 2000, {"silent": [0, 0, 0, 1]}),
])
def test_gpu_true_data_sync_matches_the_oracle_variant(oracle, nodes, count, max_clock, kw):
    seeds = np.arange(900, 900 + count, dtype=np.uint64)
    sim, g = gpu_run(seeds, nodes, max_clock, true_data_sync=True, **dict(kw))
    assert sim.kernel_info().endswith(",true,32>") and sim.kernel_info().startswith("lbft_event_loop_kernel")   # <.., TDS, TILE>
    assert ((g.status & ~np.uint32(64)) == 1).all(), np.unique(g.status)
    assert_same(oracle.run(seeds, nodes, max_clock, flags=4, **kw), g, "true data-sync N=%d" % nodes)
    for inst in (0, count - 1):
        assert sim.commit_log(inst, 0) == oracle.commit_log(seeds, nodes, inst, 0, max_clock, flags=4, **kw)


def test_default_is_unchanged_and_flag_combinations_are_refused(oracle):
    from librabft_simulator_b200 import _lib
    seeds = np.arange(900, 964, dtype=np.uint64)
    _, plain = gpu_run(seeds, 4, 1000)
    _, tds = gpu_run(seeds, 4, 1000, true_data_sync=True)
    assert_same(oracle.run(seeds, 4, 1000), plain, "default dispatch")
    assert (plain.last_states != tds.last_states).any()
    with pytest.raises(_lib.LbftError):
        gpu_run(seeds, 4, 1000, true_data_sync=True, resumable=True)
