"""BASELINE.json configs 2, 4 and 5 use features the reference does not have: uniform
delay, weighted voting rights + silent (crashed) nodes, random network partitions.  Their semantics are
fixed in the oracle first; parity for them is device-core-vs-oracle plus reference-derived invariants
(prefix-consistent logs, one-step commits, quorum arithmetic).  With every extension switched off the run
must equal the reference goldens (checked in test_hostcore_parity.py).  CPU only (host-compiled device core)."""
import numpy as np

from tests.support import assert_same

W64 = [1 + (i % 3) for i in range(64)]                    # total 127, quorum 85
SILENT64 = [1 if i % 3 == 0 and i <= 60 else 0 for i in range(64)]  # the 21 weight-1 nodes 0,3,..,60


def both(oracle, hostcore, seeds, nodes, max_clock, **kw):
    o = oracle.run(seeds, nodes, max_clock, **kw)
    h = hostcore.run(seeds, nodes, max_clock, **kw)
    assert (o.status == 1).all(), o.status
    assert ((h.status & ~np.uint32(64)) == 1).all(), h.status
    assert_same(o, h)
    return o


def test_uniform_delay_config2(oracle, hostcore):
    seeds = np.arange(1, 65, dtype=np.uint64)
    o = both(oracle, hostcore, seeds, 4, 1000, delay_kind=1, delay_lo=5, delay_hi=15)
    assert o.commit_counts.min() > 10


def test_weighted_voting_rights(oracle, hostcore):
    seeds = np.arange(7, 23, dtype=np.uint64)
    both(oracle, hostcore, seeds, 5, 1000, voting_rights=[1, 2, 5, 1, 3])
    both(oracle, hostcore, seeds, 4, 1000, voting_rights=[3, 1, 1, 1])


def test_silent_nodes(oracle, hostcore):
    seeds = np.arange(100, 116, dtype=np.uint64)
    o = both(oracle, hostcore, seeds, 4, 2000, silent=[0, 0, 0, 1])     # f = 1 of 4: still live
    assert o.commit_counts[:, :3].min() > 5 and (o.commit_counts[:, 3] == 0).all()
    o = both(oracle, hostcore, seeds, 7, 2000, silent=[1, 0, 0, 1, 0, 0, 0])
    assert (o.commit_counts[:, 0] == 0).all()
    o = both(oracle, hostcore, seeds, 4, 1000, silent=[0, 1, 1, 0])     # f = 2 of 4: no quorum, no commits
    assert (o.commit_counts == 0).all()


def test_config4_64_authors_weighted_silent(oracle, hostcore):
    seeds = np.arange(1, 3, dtype=np.uint64)
    o = both(oracle, hostcore, seeds, 64, 400, voting_rights=W64, silent=SILENT64)
    assert o.counters[:, 6].min() >= 3


def test_64_authors_plain(oracle, hostcore):
    both(oracle, hostcore, np.arange(5, 6, dtype=np.uint64), 64, 300)
    both(oracle, hostcore, np.arange(5, 7, dtype=np.uint64), 33, 300)
    both(oracle, hostcore, np.arange(5, 7, dtype=np.uint64), 20, 400)


def test_partition_fuzzing_config5(oracle, hostcore):
    for base in (1, 2, 3):
        seeds = np.arange(base * 1000, base * 1000 + 24, dtype=np.uint64)
        o = both(oracle, hostcore, seeds, 7, 1000, partition_windows=4, partition_max_len=150)
        # safety under partitions: logs stay prefix-consistent
        for i in (0, 11, 23):
            logs = [oracle.commit_log(seeds, 7, i, n, 1000, partition_windows=4, partition_max_len=150) for n in range(7)]
            longest = max(logs, key=len)
            for lg in logs:
                assert lg == longest[: len(lg)]


def test_partitions_change_results_but_not_the_main_stream_seeding(oracle):
    seeds = np.arange(50, 58, dtype=np.uint64)
    a = oracle.run(seeds, 7, 1000)
    b = oracle.run(seeds, 7, 1000, partition_windows=6, partition_max_len=300)
    assert (a.last_states != b.last_states).any()
    # the plan comes from a separate stream: the nodes' start-up delays (first N draws) are unchanged, so the
    # first timer events fire identically -> identical results when the windows are empty
    c = oracle.run(seeds, 7, 1000, partition_windows=0, partition_max_len=300)
    np.testing.assert_array_equal(a.last_states, c.last_states)


def test_silent_elision_is_exact_in_every_combination(oracle, hostcore):
    """Events addressed to silent nodes are counted at send time and never queued in plain one-shot runs (sim_core.cuh
    enqueue_network_event); the oracle queues and drops them.  Every counter must still agree — with partitions (a dropped
    send is not counted), uniform and LogNormal delays, weights, each queue kind — and the modes that keep every pop
    (recording, resumable, true data-sync) must give the same final results as the plain run."""
    from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES
    seeds = np.arange(4000, 4012, dtype=np.uint64)
    cases = [
        (5, 1500, dict(silent=[0, 1, 0, 0, 0], partition_windows=3, partition_max_len=200)),
        (7, 1000, dict(silent=[0, 0, 1, 0, 0, 1, 0], delay_kind=1, delay_lo=0, delay_hi=12)),        # zero delays: same-time events
        (4, 3000, dict(silent=[1, 0, 0, 0], voting_rights=[1, 2, 2, 2])),
        (9, 800, dict(silent=[0, 0, 0, 1, 0, 0, 0, 1, 0], partition_windows=2, partition_max_len=100)),
        (20, 400, dict(silent=[1 if i % 4 == 1 else 0 for i in range(20)])),
        (4, 20000, dict(silent=[0, 0, 1, 0])),                                                      # long horizon: heap queue
    ]
    for nodes, max_clock, kw in cases:
        plain = both(oracle, hostcore, seeds, nodes, max_clock, **kw)
        for flags in (FLAG_ROUND_SWITCHES, FLAG_RESUMABLE):
            h = hostcore.run(seeds, nodes, max_clock, flags=flags, **kw)
            np.testing.assert_array_equal(h.commit_counts, plain.commit_counts)
            np.testing.assert_array_equal(h.last_states, plain.last_states)
            np.testing.assert_array_equal(h.counters[:, list(range(8)) + [9]], plain.counters[:, list(range(8)) + [9]])
