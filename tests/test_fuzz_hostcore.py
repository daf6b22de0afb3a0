"""Fuzz parity (CPU): random committee sizes, delay models, pacemaker parameters, voting rights, silent nodes and
partition plans — the device core compiled for the host must match the oracle bit for bit on every instance that
stays within the automatically chosen table capacities, and on ALL instances once the capacities are raised."""
import numpy as np

from tests.fuzz_configs import BIG_CAPS, CAPACITY_BITS, random_config, random_modes


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
    ok = np.ones(len(seeds), dtype=bool)  # (epoch changes are simulated: bit 32 is advisory on both sides)
    assert ((g.status[ok] & ~np.uint32(64 | 32)) == 1).all(), (n, max_clock, kw, np.unique(g.status))
    np.testing.assert_array_equal(o.commit_counts[ok], g.commit_counts[ok], err_msg=str((n, max_clock, seed0, kw)))
    np.testing.assert_array_equal(o.last_states[ok], g.last_states[ok], err_msg=str((n, max_clock, seed0, kw)))
    np.testing.assert_array_equal(o.counters[ok][:, :8], g.counters[ok][:, :8], err_msg=str((n, max_clock, seed0, kw)))
    return int(flagged.sum())


def test_fuzz_hostcore_vs_oracle(oracle, hostcore):
    rng = np.random.default_rng(20260922)
    for _ in range(250):
        n, max_clock, seed0, kw = random_config(rng)
        check(oracle, hostcore.run, n, max_clock, seed0, kw, count=6)


def check_modes(oracle, run, switches, n, max_clock, seed0, kw, flags, stops, count):
    """`run(seeds, n, max_clock, flags, stops, **kw)` and `switches(seeds, n, max_clock, flags, stops, instance, **kw)` are the
    implementation under test (host-compiled core or GPU); the oracle runs the same stops / records the same way."""
    seeds = np.arange(seed0, seed0 + count, dtype=np.uint64)
    what = str((n, max_clock, seed0, kw, flags, stops))

    def ref(k):
        return oracle.run(seeds, n, max_clock, **k) if stops is None else oracle.run_staged(seeds, n, stops, max_clock, **k)

    g = run(seeds, n, max_clock, flags, stops, **kw)
    if ((g.status & CAPACITY_BITS) != 0).any():
        kw = dict(kw, **BIG_CAPS)
        g = run(seeds, n, max_clock, flags, stops, **kw)
        assert ((g.status & CAPACITY_BITS) == 0).all(), what
    o = ref(kw)
    ok = (o.status & 32) == 0
    np.testing.assert_array_equal(o.commit_counts[ok], g.commit_counts[ok], err_msg=what)
    np.testing.assert_array_equal(o.last_states[ok], g.last_states[ok], err_msg=what)
    np.testing.assert_array_equal(o.counters[ok][:, :8], g.counters[ok][:, :8], err_msg=what)
    assert (g.counters[:, 11] == 0).all(), what          # nothing elided in either mode
    if flags & 1:
        for i in (0, count - 1):
            if ok[i]:
                want = (oracle.round_switches(seeds, n, i, max_clock, **kw) if stops is None
                        else oracle.round_switches_staged(seeds, n, i, stops, max_clock, **kw))
                assert switches(seeds, n, max_clock, flags, stops, i, **kw) == want, what


def test_fuzz_modes_hostcore_vs_oracle(oracle, hostcore):
    def run(seeds, n, max_clock, flags, stops, **kw):
        if stops is None:
            return hostcore.run(seeds, n, max_clock, flags=flags, **kw)
        return hostcore.run_staged(seeds, n, stops, max_clock, flags=flags, **kw)

    def switches(seeds, n, max_clock, flags, stops, i, **kw):
        if stops is None:
            return hostcore.round_switches(seeds, n, i, max_clock, flags=flags, **kw)
        return hostcore.round_switches_staged(seeds, n, i, stops, max_clock, flags=flags, **kw)

    rng = np.random.default_rng(20260923)
    for _ in range(150):
        n, max_clock, seed0, kw = random_config(rng)
        flags, stops = random_modes(rng, max_clock)
        check_modes(oracle, run, switches, n, max_clock, seed0, kw, flags, stops, count=5)
