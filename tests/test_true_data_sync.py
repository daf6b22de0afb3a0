"""SURVEY §8(f).4, second half — LBFT_FLAG_TRUE_DATA_SYNC: the opt-in NON-PARITY variant in which a data-sync request is
answered by the node it was sent to (handle_request on `sender`, from its records: data_sync.rs:183-207,
record_store.rs:766-831) and the requester inserts the records of the response (data_sync.rs:209-240) — instead of the
reference simulator's dispatch to the requester itself (simulator.rs:446), which stays the default because
the reference's golden runs pin it.
What is checked: the device core (compiled for the host) against the oracle running the same variant — the oracle's
handle_request / handle_response / unknown_records / known_quorum_certificate_rounds are the function-by-function
restatements of the reference's, only the dispatch differs — bit-exact over committee sizes, queue modes and the extension
features; that the variant really changes outcomes; and that the flag is refused where it is not built."""
import numpy as np
import pytest

from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES, FLAG_TRUE_DATA_SYNC, assert_same

CASES = [
    (2, 1000, {}, 2), (3, 1000, {}, 2), (4, 1000, {}, 2),
    (4, 1000, dict(queue_cap=128), 1), (5, 1000, {}, 1), (5, 1000, dict(voting_rights=[1, 2, 3, 1, 1]), 1),
    (7, 1000, {}, 3), (7, 1000, dict(partition_windows=3, partition_max_len=150), 3), (12, 1000, {}, 3),
    (4, 2000, dict(silent=[0, 0, 0, 1]), 1), (6, 4200, {}, 0),
    (4, 600, dict(delay_kind=1, delay_lo=0, delay_hi=6, round_cap=256), 1),
    (9, 700, dict(delay_mean=25.0, delay_variance=200.0), 3),
]


@pytest.mark.parametrize("nodes,max_clock,kw,qmode", CASES)
def test_true_data_sync_matches_the_oracle_variant(oracle, hostcore, nodes, max_clock, kw, qmode):
    seeds = np.arange(300, 348, dtype=np.uint64)
    assert hostcore.setup_info(nodes, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw)["queue_scan"] == qmode
    o = oracle.run(seeds, nodes, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw)
    h = hostcore.run(seeds, nodes, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw)
    assert ((h.status & ~np.uint32(64)) == 1).all(), np.unique(h.status)
    assert_same(o, h, "true data-sync N=%d" % nodes)


def test_the_variant_changes_outcomes_and_the_default_does_not_move(oracle, hostcore):
    seeds = np.arange(300, 364, dtype=np.uint64)
    plain, tds = hostcore.run(seeds, 4, 1000), hostcore.run(seeds, 4, 1000, flags=FLAG_TRUE_DATA_SYNC)
    assert (plain.last_states != tds.last_states).any(axis=1).mean() > 0.5      # responses now carry records that get inserted
    assert_same(oracle.run(seeds, 4, 1000), plain, "default dispatch (simulator.rs:446)")
    golden = hostcore.run([52], 3, 1000)                                          # simulated_run.rs:45-66 is the default's
    assert golden.commit_counts.tolist() == [[27, 27, 27]] and golden.last_states.tolist() == [[11134312813757838303] * 3]


def test_the_flag_is_refused_where_it_is_not_built(hostcore):
    for bad in (dict(flags=FLAG_TRUE_DATA_SYNC | FLAG_RESUMABLE), dict(flags=FLAG_TRUE_DATA_SYNC | FLAG_ROUND_SWITCHES),
                dict(flags=FLAG_TRUE_DATA_SYNC, commands_per_epoch=10)):
        with pytest.raises(RuntimeError, match="TRUE_DATA_SYNC"):
            hostcore.setup_info(4, 1000, **bad)


def test_fuzz_true_data_sync(oracle, hostcore):
    from tests.fuzz_configs import BIG_CAPS, CAPACITY_BITS, random_config
    rng = np.random.default_rng(20260924)
    done = 0
    for _ in range(80):
        n, max_clock, seed0, kw = random_config(rng)
        kw.pop("commands_per_epoch", None)
        seeds = np.arange(seed0, seed0 + 6, dtype=np.uint64)
        h = hostcore.run(seeds, n, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw)
        if (h.status & CAPACITY_BITS).any():
            kw = dict(kw, **BIG_CAPS)
            h = hostcore.run(seeds, n, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw)
            if (h.status & CAPACITY_BITS).any():
                continue   # (zero-delay networks make far more progress once data-sync works: beyond the largest tables)
        assert_same(oracle.run(seeds, n, max_clock, flags=FLAG_TRUE_DATA_SYNC, **kw), h, str((n, max_clock, seed0, kw)))
        done += 1
    assert done >= 60
