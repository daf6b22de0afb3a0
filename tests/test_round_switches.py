"""SURVEY 8(f).3 — DataWriter round switches (bft-lib/src/data_writer.rs:34-50, sampled at simulator.rs:393-394).

CPU side: the oracle's restatement against the device state machine compiled for the host, and the text of
round_switches.txt.  PARITY UNPINNED for this output: the reference holds no test, golden or fixture for DataWriter
and cannot be run here (no Rust toolchain), so the oracle is a reading of the source; the event sequence it samples
is the one the commit-log goldens pin (tests/test_oracle_golden.py)."""
import csv
import io

import pytest

from librabft_simulator_b200.simulator import format_round_switches_csv
from tests.support import FLAG_ROUND_SWITCHES, assert_same


# (num_nodes, max_clock, instances, extra config) — queue modes 1 (HBM scan), 3 (calendar), 0 (heap); both mask widths.
# (mode 2, the shared-memory queue, needs an explicit queue_cap=64 while recording: see QMODES below)
CASES = [
    (1, 1000, 6, dict(round_cap=1056)),
    (2, 1000, 12, {}),
    (3, 1000, 40, {}),
    (4, 1000, 40, {}),
    (4, 1000, 12, dict(delay_kind=1, delay_lo=1, delay_hi=30)),
    (4, 400, 12, dict(delay_kind=1, delay_lo=0, delay_hi=3, round_cap=256)),
    (4, 1000, 12, dict(silent=[0, 1, 0, 0])),
    (4, 1000, 12, dict(voting_rights=[1, 2, 3, 4])),
    (5, 1000, 12, {}),
    (7, 1000, 8, dict(partition_windows=3, partition_max_len=200)),
    (8, 1000, 8, {}),
    (9, 700, 8, dict(delay_kind=1, delay_lo=0, delay_hi=3, round_cap=256)),
    (20, 600, 2, {}),
    (40, 300, 2, {}),
    (4, 6000, 3, {}),  # long horizon, small committee: still the HBM scan queue
    (7, 4500, 2, {}),  # N > 5 beyond the calendar queue's horizon: binary heap
]


@pytest.mark.parametrize("N,max_clock,count,kw", CASES)
def test_hostcore_round_switches_match_oracle(oracle, hostcore, N, max_clock, count, kw):
    seeds = list(range(7000 + 13 * N, 7000 + 13 * N + count))
    ref = oracle.run(seeds, N, max_clock, **kw)
    rec = hostcore.run(seeds, N, max_clock, flags=FLAG_ROUND_SWITCHES, **kw)
    assert not (rec.status & 0xFFFFFFFE).any(), "capacity/invariant flags raised while recording: %s" % sorted(set(rec.status.tolist()))
    assert_same(ref, rec, "(recording must not change the simulation)")
    assert (rec.counters[:, 11] == 0).all(), "no timer may be elided while recording: each pop is a sampling point"
    for i in range(count):
        want = oracle.round_switches(seeds, N, i, max_clock, **kw)
        assert hostcore.round_switches(seeds, N, i, max_clock, **kw) == want, "instance %d" % i
        assert want, "nothing recorded"


# the queue mode each kind of case is meant to exercise, asserted so that a comment cannot drift from the selection logic
QMODES = [(3, 1000, dict(queue_cap=64), 2), (4, 1000, {}, 1), (4, 6000, {}, 1), (8, 1000, {}, 3), (7, 4500, {}, 0)]


@pytest.mark.parametrize("N,max_clock,kw,qmode", QMODES)
def test_queue_mode_coverage(oracle, hostcore, N, max_clock, kw, qmode):
    assert hostcore.setup_info(N, max_clock, flags=FLAG_ROUND_SWITCHES, **kw)["queue_scan"] == qmode
    seeds = [811, 812, 813]
    for i in range(len(seeds)):
        assert hostcore.round_switches(seeds, N, i, max_clock, **kw) == oracle.round_switches(seeds, N, i, max_clock, **kw)


@pytest.mark.parametrize("seed,N", [(52, 3), (48, 8)])  # the reference's golden runs (simulated_run.rs:46-93)
def test_oracle_round_switch_shape(oracle, seed, N):
    sw = oracle.round_switches([seed], N, 0)
    res = oracle.run([seed], N)
    per_node = {}
    for node, rnd, time in sw:
        per_node.setdefault(node, []).append((rnd, time))
    assert sorted(per_node) == list(range(N))
    for node, entries in per_node.items():
        rounds = [r for r, _ in entries]
        times = [t for _, t in entries]
        assert rounds == sorted(set(rounds)) and rounds[0] >= 1, "rounds are recorded once, ascending"
        assert times == sorted(times) and 1 <= times[0] and times[-1] <= 1000
    # the last switch of a run is only seen if another event is popped afterwards
    assert max(r for _, r, _ in sw) <= int(res.counters[0, 6]) <= max(r for _, r, _ in sw) + 1


def test_unknown_flag_bits_are_rejected(hostcore):
    with pytest.raises(RuntimeError, match="flags"):
        hostcore.setup_info(4, flags=8)
    with pytest.raises(RuntimeError, match="LBFT_FLAG_ROUND_SWITCHES"):
        hostcore.round_switches([1], 4, 0, flags=0)


def test_recording_keeps_the_bench_layout_untouched(hostcore):
    plain, rec = hostcore.setup_info(4), hostcore.setup_info(4, flags=FLAG_ROUND_SWITCHES)
    assert plain["queue_scan"] == 2 and plain["queue_cap"] == 64 and plain["words"] == 852
    assert rec["words"] == plain["words"] - 64 * 2 + 128 * 2 + 4 * (rec["round_cap"] + 1) and rec["queue_cap"] == 128


def test_csv_text():
    # node 0: rounds 1, 2, 4 (3 skipped); node 1: rounds 1, 3; max round 4 -> rows for rounds 0..3 only
    sw = [(0, 1, 10), (0, 2, 39), (0, 4, 90), (1, 1, 11), (1, 3, 70)]
    assert format_round_switches_csv(2, sw) == "node 0,node 1\n,\n10,11\n39,\n,70\n"
    assert format_round_switches_csv(3, []) == "node 0,node 1,node 2\n"
    # one column: a record that is a single empty field is written as "" (csv crate), never as an empty line
    assert format_round_switches_csv(1, [(0, 1, 5), (0, 3, 9)]) == 'node 0\n""\n5\n""\n'


@pytest.mark.parametrize("seed,N", [(52, 3), (48, 8), (7, 1)])
def test_csv_reads_back_through_the_reference_plotter_steps(oracle, seed, N):
    kw = dict(round_cap=1056) if N == 1 else {}
    sw = oracle.round_switches([seed], N, 0, **kw)
    text = format_round_switches_csv(N, sw)
    rows = list(csv.reader(io.StringIO(text)))  # round_plotter.py:11-14
    assert rows[0] == ["node %d" % n for n in range(N)]
    rows = rows[1:]  # round_plotter.py:53
    max_round = max(r for _, r, _ in sw)
    assert len(rows) == max_round and all(len(r) == N for r in rows)
    back = [(n, i, int(rows[i][n])) for n in range(N) for i in range(len(rows)) if rows[i][n]]  # row index = round (:27-33)
    assert back == [e for e in sw if e[1] < max_round]
