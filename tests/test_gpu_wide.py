"""The warp-per-instance kernel (lbft_wide_kernel, sim_core.cuh G = 32) at the sizes it exists for — BASELINE configs[1]
(1 024 x 4), configs[3] (64 authors, weighted, silent) and configs[4] (7 authors, partitions) — bit-exact against the
oracle, plus the host's automatic choice between the two kernel families.  (tests/test_gpu_parity.py and test_gpu_fuzz.py
run their whole case matrices once per family through the `kernel_choice` fixture.)"""
import numpy as np
import pytest

from tests.support import assert_same
from tests.test_gpu_parity import SILENT64, W64, gpu_run

pytestmark = pytest.mark.gpu


def test_automatic_choice(monkeypatch):
    monkeypatch.delenv("LBFT_FORCE_KERNEL", raising=False)
    from tests.test_gpu_parity import make_sim
    picks = {}
    for name, count, nodes, kw in (("config2", 1024, 4, {}), ("config3", 65536, 4, {}), ("config5", 16384, 7, {}),
                                   ("config5p", 16384, 7, {"partition_windows": 4, "partition_max_len": 150}),
                                   ("mid7", 8192, 7, {}), ("full7", 65536, 7, {}), ("big64", 8192, 64, {}),
                                   ("small64", 64, 64, {}), ("resumable", 64, 4, {"resumable": True}),
                                   ("recording", 64, 7, {"record_round_switches": True}), ("mid4", 8192, 4, {}),
                                   ("long7", 64, 7, {"max_clock": 20000})):
        max_clock = kw.pop("max_clock", 1000)
        sim = make_sim(np.arange(1, count + 1, dtype=np.uint64), nodes, **kw).create(max_clock)
        picks[name] = sim.kernel_info()
        sim.close()
    # <NMAX, QMODE, SMEM, lanes per instance, EP, FX> / <NMAX, QMODE, FX, REC, RES, EP, TDS, instances per warp tile>
    assert picks["config2"] == "lbft_wide_kernel<16,2,true,32,false,0>"    # a warp per instance, whole instance in shared memory
    assert picks["config3"] == picks["mid4"] == "lbft_event_loop_kernel<16,2,1,false,false,false,false,32>"
    assert picks["config5"] == "lbft_event_loop_kernel<16,3,0,false,false,false,false,8>"  # one wave of 8-instance warp tiles
    assert picks["config5p"] == "lbft_event_loop_kernel<16,3,2,false,false,false,false,8>"  # ... BASELINE configs[4]: its compile-time shape
    assert picks["mid7"] == "lbft_wide_kernel<16,2,false,8,false,0>"       # 8 lanes per instance, four instances per warp
    assert picks["full7"] == "lbft_event_loop_kernel<16,3,0,false,false,false,false,32>"
    assert picks["small64"] == "lbft_wide_kernel<64,3,false,32,false,0>"
    assert picks["big64"] == "lbft_wide_kernel<64,3,false,8,false,3>"      # BASELINE configs[3]: compile-time shape, 8 lanes per instance
    assert picks["long7"] == "lbft_wide_kernel<16,0,false,32,false,0>"     # beyond the 14-bit times of the compact queue
    assert picks["resumable"].startswith("lbft_event_loop_kernel") and picks["recording"].startswith("lbft_event_loop_kernel")


@pytest.mark.parametrize("count,nodes,max_clock,kw", [
    (1024, 4, 1000, {}),                                                   # BASELINE configs[1], LogNormal leg
    (1024, 4, 1000, {"delay_kind": 1, "delay_lo": 5, "delay_hi": 15}),     # ... uniform-delay leg
    (2048, 7, 1000, {"partition_windows": 4, "partition_max_len": 150}),   # BASELINE configs[4] (an eighth of the batch)
    (16, 64, 1000, {"voting_rights": W64, "silent": SILENT64}),            # BASELINE configs[3]: full horizon, 16 instances
    (33, 20, 600, {}),
    (5, 40, 400, {"voting_rights": [1 + (i % 4) for i in range(40)]}),
    (96, 5, 2500, {}),                                                      # QMODE 1: 64-bit keys in HBM, scanned by the warp
    (8, 6, 4200, {}),                                                       # QMODE 0: binary heap
])
def test_wide_kernel_matches_oracle(oracle, monkeypatch, count, nodes, max_clock, kw):
    monkeypatch.setenv("LBFT_FORCE_KERNEL", "wide")
    seeds = np.arange(77000, 77000 + count, dtype=np.uint64)
    sim, g = gpu_run(seeds, nodes, max_clock, **dict(kw))
    assert sim.kernel_info().startswith("lbft_wide_kernel")
    assert ((g.status & ~np.uint32(64)) == 1).all(), np.unique(g.status)
    assert_same(oracle.run(seeds, nodes, max_clock, **kw), g, "wide kernel N=%d" % nodes)
    for inst in sorted({0, count - 1}):
        assert sim.commit_log(inst, nodes - 1) == oracle.commit_log(seeds, nodes, inst, nodes - 1, max_clock, **kw)
    rows, lens = sim.commit_logs()
    np.testing.assert_array_equal(lens, g.commit_counts)


def test_both_families_agree_on_a_large_committee_batch(monkeypatch):
    """Same seeds through both kernel families: identical results including every counter the oracle defines."""
    seeds = np.arange(5, 5 + 64, dtype=np.uint64)
    res = {}
    for fam in ("thread", "wide"):
        monkeypatch.setenv("LBFT_FORCE_KERNEL", fam)
        sim, g = gpu_run(seeds, 24, 500)
        res[fam] = g
    assert_same(res["thread"], res["wide"], "thread vs wide")
