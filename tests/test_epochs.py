"""SURVEY §8(f).2 — epoch changes (librabft-v2/src/node.rs:329-348, data_sync.rs:123-146, simulated_context.rs:199-207,
pacemaker.rs:158, node.rs:372-376; quirk B.9(iii) base_types.rs:31-37): when a node's commit count reaches a multiple of
commands_per_epoch it swaps in a fresh record store, resets its voting constraints and stops delivering commits; records
of another epoch are dropped, a sender that is ahead triggers a data-sync request.  The device core (compiled for the
host) against the oracle, bit-exact, over committee sizes x commands_per_epoch, with the extension features mixed in.
PARITY: the oracle implements process_commits faithfully; the reference has no test that crosses an epoch boundary
(its CLI default of 30 000 commands per epoch is never reached), so this is oracle-only parity."""
import numpy as np
import pytest

from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES, assert_same

CASES = [(2, {}), (3, {}), (4, {}), (7, {}), (12, {}),
         (5, dict(voting_rights=[1, 2, 3, 1, 1])), (4, dict(silent=[0, 0, 0, 1])),
         (7, dict(partition_windows=3, partition_max_len=150)), (4, dict(delay_kind=1, delay_lo=0, delay_hi=4)),
         (4, dict(queue_cap=64)), (6, dict(round_cap=160))]


@pytest.mark.parametrize("nodes,kw", CASES)
@pytest.mark.parametrize("cpe", [1, 2, 5, 10, 30])
def test_epoch_changes_match_oracle(oracle, hostcore, nodes, kw, cpe):
    seeds = np.arange(100, 124, dtype=np.uint64)
    o = oracle.run(seeds, nodes, 1000, commands_per_epoch=cpe, **kw)
    h = hostcore.run(seeds, nodes, 1000, commands_per_epoch=cpe, **kw)
    assert ((h.status & ~np.uint32(64 | 32)) == 1).all(), np.unique(h.status)      # the epoch bit (32) is advisory now
    np.testing.assert_array_equal(o.status & 32, h.status & 32)                    # ... and set exactly where the oracle crosses one
    assert_same(o, h, "N=%d commands_per_epoch=%d" % (nodes, cpe))
    if cpe <= 10 and nodes >= 3 and "silent" not in kw and "partition_windows" not in kw:
        assert ((o.status & 32) != 0).mean() > 0.8 and o.commit_counts.max() <= 3 * cpe + 3   # the reference semantics stall


def test_known_outcome_four_nodes_ten_commands(oracle, hostcore):
    # the example VERDICT r1 quotes: the oracle ends at [10, 9, 10, 9]-type counts where the round-1 device carried on to 34
    seeds = np.arange(1, 33, dtype=np.uint64)
    o, h = oracle.run(seeds, 4, 1000, commands_per_epoch=10), hostcore.run(seeds, 4, 1000, commands_per_epoch=10)
    assert_same(o, h)
    assert set(np.unique(h.commit_counts).tolist()) <= {9, 10, 11, 12} and h.commit_counts.max(axis=1).min() >= 10


def test_single_epoch_layout_is_unchanged(hostcore):
    # commands_per_epoch that cannot be reached keeps the one-epoch layout (every BASELINE configuration)
    a = hostcore.setup_info(4, 1000)
    b = hostcore.setup_info(4, 1000, commands_per_epoch=128)
    c = hostcore.setup_info(4, 1000, commands_per_epoch=10)
    assert a == b and c["round_cap"] > a["round_cap"] and c["words"] > a["words"]


def test_more_epochs_than_the_tables_hold_is_flagged(oracle, hostcore):
    # a lone node runs through an epoch every few milliseconds: beyond 32 epochs the instance is flagged, never silently wrong
    h = hostcore.run([5, 6], 1, 1000, commands_per_epoch=2, round_cap=2048)
    assert (h.status & 2).all()


@pytest.mark.parametrize("nodes,cpe", [(3, 5), (4, 10), (7, 30)])
def test_epochs_with_recording_and_staged_runs(oracle, hostcore, nodes, cpe):
    seeds = list(range(40, 52))
    kw = dict(commands_per_epoch=cpe)
    for i in (0, 5, 11):   # DataWriter: after an epoch change the active round restarts at 1; only new maxima are recorded
        assert hostcore.round_switches(seeds, nodes, i, 1000, **kw) == oracle.round_switches(seeds, nodes, i, 1000, **kw)
    stops = [300, 650, 1000]
    assert_same(oracle.run_staged(seeds, nodes, stops, 1000, **kw), hostcore.run_staged(seeds, nodes, stops, 1000, **kw), "staged")
    assert (hostcore.round_switches_staged(seeds, nodes, 3, stops, 1000, **kw) ==
            oracle.round_switches_staged(seeds, nodes, 3, stops, 1000, **kw))
