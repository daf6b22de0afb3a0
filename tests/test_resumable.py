"""SURVEY 8(f).4 — resumable loop_until (simulator.rs:380-475 called again with a larger clock) on the device state
machine compiled for the host, against the oracle's loop_until called once per stop on the same Simulator.

The reference's exit behaviour is part of the contract: each call pops the first event beyond its clock and drops it
(simulator.rs:383-391), so a staged run is a DIFFERENT simulation from a one-shot run — the test asserts that too.
PARITY: the oracle's loop_until is the restatement pinned by the commit-log goldens; calling it repeatedly has no
golden of its own in the reference (no test resumes a simulator), so staged parity is oracle-only."""
import pytest

from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES, assert_same

SCHEDULES = [
    [300, 650, 1000],
    [0, 1, 2, 500, 500, 501, 1000],   # stops before any startup timer drop those timers, as in the reference
    [17, 400, 399, 1000],             # a smaller clock than before: one more event is dropped, nothing else happens
    [999, 1000],
]
# (num_nodes, horizon, instances, extra config, queue mode expected)
CASES = [
    (2, 1000, 16, dict(queue_cap=64), 2),          # shared-memory queue: spilled to / restored from the save area
    (3, 1000, 24, dict(queue_cap=64), 2),
    (3, 1000, 16, {}, 1),
    (4, 1000, 24, {}, 1),
    (4, 1000, 8, dict(silent=[0, 1, 0, 0]), 1),
    (5, 1000, 12, dict(voting_rights=[1, 2, 3, 4, 5]), 1),
    (7, 1000, 8, dict(partition_windows=3, partition_max_len=200), 3),
    (9, 1000, 8, dict(delay_kind=1, delay_lo=0, delay_hi=3, round_cap=256), 3),
    (20, 1000, 2, {}, 3),
    (40, 300, 2, {}, 3),
    (4, 6000, 3, {}, 1),                              # long horizon, small committee: still the scan queue
    (7, 4500, 2, {}, 0),                              # N > 5 beyond the calendar queue's horizon: binary heap
]


def scaled(schedule, horizon):
    return [t * horizon // 1000 for t in schedule]


@pytest.mark.parametrize("N,horizon,count,kw,qmode", CASES)
def test_staged_hostcore_matches_staged_oracle(oracle, hostcore, N, horizon, count, kw, qmode):
    seeds = list(range(9100 + 7 * N, 9100 + 7 * N + count))
    info = hostcore.setup_info(N, horizon, flags=FLAG_RESUMABLE, **kw)
    assert info["queue_scan"] == qmode
    one_shot = oracle.run(seeds, N, horizon, **kw)
    differs = 0
    for schedule in SCHEDULES:
        stops = scaled(schedule, horizon)
        for k in sorted({1, len(stops) // 2, len(stops)}):   # the state after the first k calls
            ref = oracle.run_staged(seeds, N, stops[:k], horizon, **kw)
            got = hostcore.run_staged(seeds, N, stops[:k], horizon, **kw)
            assert not (got.status & 0xFFFFFFFE).any(), sorted(set(got.status.tolist()))
            assert_same(ref, got, "after %s of %s" % (stops[:k], stops))
            assert (got.counters[:, 11] == 0).all(), "no timer may be elided in a resumable run"
        differs += int((ref.last_states != one_shot.last_states).any())
    if N <= 5:  # in a large committee one lost message out of ~3 N^2 per round rarely changes what is committed
        assert differs >= 2, "staged runs should differ from the one-shot run: the dropped events are not being dropped"


@pytest.mark.parametrize("N,kw", [(3, {}), (4, {}), (8, {}), (3, dict(queue_cap=64))])
def test_staged_round_switches(oracle, hostcore, N, kw):
    seeds = list(range(9300, 9306))
    for stops in ([300, 650, 1000], [17, 400, 399, 1000]):
        for i in range(len(seeds)):
            want = oracle.round_switches_staged(seeds, N, i, stops, **kw)
            got = hostcore.round_switches_staged(seeds, N, i, stops, flags=FLAG_RESUMABLE | FLAG_ROUND_SWITCHES, **kw)
            assert got == want and want


def test_one_stop_at_the_horizon_is_the_one_shot_run(oracle, hostcore):
    seeds = list(range(9400, 9432))
    assert_same(oracle.run(seeds, 4), hostcore.run_staged(seeds, 4, [1000]))
    # and a resumable configuration run in one go through the plain entry point
    assert_same(oracle.run(seeds, 4), hostcore.run(seeds, 4, flags=FLAG_RESUMABLE))


def test_staged_needs_the_flag_and_a_stop_inside_the_horizon(hostcore):
    with pytest.raises(RuntimeError, match="LBFT_FLAG_RESUMABLE"):
        hostcore.run_staged([1], 4, [500, 1000], flags=0)
    with pytest.raises(RuntimeError, match="stop_clock"):
        hostcore.run_staged([1], 4, [500, 1001])
    with pytest.raises(RuntimeError, match="flags"):
        hostcore.setup_info(4, flags=8)


def test_save_area_is_appended_after_everything_else(hostcore):
    plain = hostcore.setup_info(3, queue_cap=64)
    res = hostcore.setup_info(3, queue_cap=64, flags=FLAG_RESUMABLE)
    both = hostcore.setup_info(3, queue_cap=64, flags=FLAG_RESUMABLE | FLAG_ROUND_SWITCHES)
    rc, qc = plain["round_cap"], plain["queue_cap"]
    assert plain["queue_scan"] == res["queue_scan"] == 2
    even = lambda w: (w + 1) & ~1   # an instance is a whole number of 8-byte units (sim_params.h make_layout)
    assert res["words"] == even(plain["words"] + 40 + rc + qc + (qc + 1) // 2)
    assert both["words"] == even(plain["words"] + 40 + rc + qc + (qc + 1) // 2 + 3 * (rc + 1))


def test_many_stops_do_not_leak_notification_slots(oracle, hostcore):
    """the event dropped at a stop (simulator.rs:389-391) may be a DataSyncNotifyEvent; its reference to
    the shared notification snapshot has to be released, or every stop loses a payload slot for good (100 stops over a
    1000 ms horizon used to end in LBFT_ST_PAYLOAD_OVERFLOW with the default 32-slot pool)."""
    N, horizon, seeds = 4, 3000, list(range(7700, 7732))
    stops = list(range(20, horizon + 1, 20))  # 150 stops
    one = hostcore.run(seeds, N, horizon, flags=FLAG_RESUMABLE)
    got = hostcore.run_staged(seeds, N, stops, horizon)
    assert not (got.status & 0xFFFFFFFE).any(), sorted(set(got.status.tolist()))
    assert_same(oracle.run_staged(seeds, N, stops, horizon), got, "150 stops")
    # high-water mark of in-flight snapshots (lbft_instance_counters.max_payloads): same order as a one-shot run
    assert got.counters[:, 10].max() <= one.counters[:, 10].max() + 4, (got.counters[:, 10].max(), one.counters[:, 10].max())
