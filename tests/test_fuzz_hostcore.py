"""Fuzz parity (CPU): random committee sizes, delay models, pacemaker parameters, voting rights, silent nodes and
partition plans — the device core compiled for the host must match the oracle bit for bit on every instance that
stays within the automatically chosen table capacities, and on ALL instances once the capacities are raised."""
import numpy as np

from tests.fuzz_configs import BIG_CAPS, CAPACITY_BITS, random_config
from tests.support import assert_same


def check(oracle, runner, n, max_clock, seed0, kw, count):
    seeds = np.arange(seed0, seed0 + count, dtype=np.uint64)
    o = oracle.run(seeds, n, max_clock, **kw)
    g = runner(seeds, n, max_clock, **kw)
    flagged = (g.status & CAPACITY_BITS) != 0
    if flagged.any():
        big = dict(kw, **BIG_CAPS)
        o = oracle.run(seeds, n, max_clock, **big)
        g = runner(seeds, n, max_clock, **big)
        assert ((g.status & CAPACITY_BITS) == 0).all(), (n, max_clock, kw, np.unique(g.status))
    ok = (o.status & 32) == 0  # an epoch end is flagged on both sides and not modelled by the device
    assert ((g.status[ok] & ~np.uint32(64)) == 1).all(), (n, max_clock, kw, np.unique(g.status))
    np.testing.assert_array_equal(o.commit_counts[ok], g.commit_counts[ok], err_msg=str((n, max_clock, seed0, kw)))
    np.testing.assert_array_equal(o.last_states[ok], g.last_states[ok], err_msg=str((n, max_clock, seed0, kw)))
    np.testing.assert_array_equal(o.counters[ok][:, :8], g.counters[ok][:, :8], err_msg=str((n, max_clock, seed0, kw)))
    return int(flagged.sum())


def test_fuzz_hostcore_vs_oracle(oracle, hostcore):
    rng = np.random.default_rng(20260922)
    for _ in range(250):
        n, max_clock, seed0, kw = random_config(rng)
        check(oracle, hostcore.run, n, max_clock, seed0, kw, count=6)
