"""SURVEY 8(f).3 on the device: lbft_round_switches / loop_until(max_clock, csv_path) through the C ABI against the
oracle's DataWriter restatement (data_writer.rs:34-96, simulator.rs:380-381, 393-395, 470-472).
PARITY UNPINNED for this output (see tests/test_round_switches.py): checked bit-exactly against the oracle only."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (num_nodes, max_clock, instances, delay, extra config, oracle kwargs)
CASES = [
    (3, 1000, 96, ("lognormal", 10.0, 4.0), {}, {}),                       # HBM scan queue, three tiles
    (4, 1000, 70, ("lognormal", 10.0, 4.0), {}, {}),                       # the bench committee size, ragged last tile
    (4, 1000, 33, ("uniform", 1, 30), {}, dict(delay_kind=1, delay_lo=1, delay_hi=30)),
    (8, 1000, 40, ("lognormal", 10.0, 4.0), {}, {}),                       # calendar queue
    (9, 700, 33, ("uniform", 0, 3), dict(round_cap=256), dict(delay_kind=1, delay_lo=0, delay_hi=3, round_cap=256)),
    (40, 300, 3, ("lognormal", 10.0, 4.0), {}, {}),                        # two-word author masks
    (4, 6000, 34, ("lognormal", 10.0, 4.0), {}, {}),                       # long horizon, small committee: HBM scan queue
    (7, 4500, 33, ("lognormal", 10.0, 4.0), {}, {}),                       # N > 5 beyond the calendar horizon: binary heap
    (3, 1000, 40, ("lognormal", 10.0, 4.0), dict(queue_cap=64), dict(queue_cap=64)),   # shared-memory queue while recording
]


def delay_of(spec):
    from librabft_simulator_b200 import RandomDelay
    return RandomDelay.new(spec[1], spec[2]) if spec[0] == "lognormal" else RandomDelay.uniform(spec[1], spec[2])


@pytest.mark.parametrize("N,max_clock,count,delay,kw,okw", CASES)
def test_round_switches_match_oracle(oracle, N, max_clock, count, delay, kw, okw):
    from librabft_simulator_b200 import BatchSimulator
    seeds = np.arange(3000 + 17 * N, 3000 + 17 * N + count, dtype=np.uint64)
    ref = oracle.run(seeds, N, max_clock, **okw)
    with BatchSimulator(seeds, N, delay_of(delay), record_round_switches=True, **kw) as sim:
        res = sim.loop_until(max_clock)
        np.testing.assert_array_equal(ref.last_states, res.last_committed_states)   # recording changes nothing
        np.testing.assert_array_equal(ref.commit_counts, res.commit_counts)
        np.testing.assert_array_equal(ref.counters[:, :8], res.counters[:, :8])
        assert (res.counters[:, 11] == 0).all(), "timers were elided while recording"
        picks = sorted({0, 1, 31, 32, count // 2, count - 1} & set(range(count)))      # tile / lane boundaries
        for i in picks:
            assert sim.round_switches(i) == oracle.round_switches(seeds, N, i, max_clock, **okw), "instance %d" % i


def test_round_switches_need_the_flag(oracle):
    from librabft_simulator_b200 import BatchSimulator, RandomDelay, _lib
    seeds = np.arange(1, 33, dtype=np.uint64)
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0)) as sim:
        sim.loop_until(1000)
        with pytest.raises(_lib.LbftError) as e:
            sim.round_switches(0)
        assert e.value.code == -3 and "LBFT_FLAG_ROUND_SWITCHES" in str(e.value)
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0), record_round_switches=True) as sim:
        sim.create(1000)
        with pytest.raises(_lib.LbftError) as e:
            sim.round_switches(0)          # before lbft_run
        assert e.value.code == -3
        sim.run()
        with pytest.raises(_lib.LbftError):
            sim.round_switches(32)         # out of range


def test_plain_and_recording_runs_agree_on_the_bench_committee():
    """The compile-time-layout kernel (no recording) and the generic kernel (recording) simulate the same thing."""
    from librabft_simulator_b200 import BatchSimulator, RandomDelay
    seeds = np.arange(52, 52 + 256, dtype=np.uint64)
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0)) as a, \
            BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0), record_round_switches=True) as b:
        ra, rb = a.loop_until(1000), b.loop_until(1000)
        np.testing.assert_array_equal(ra.last_committed_states, rb.last_committed_states)
        np.testing.assert_array_equal(ra.counters[:, :3], rb.counters[:, :3])
        # popped timers: the recording run pops the duplicates the plain run accounts for without queueing them
        np.testing.assert_array_equal(ra.counters[:, 3], rb.counters[:, 3])
        np.testing.assert_array_equal(ra.counters[:, 4], rb.counters[:, 4])


@pytest.mark.parametrize("seed,N", [(52, 3), (48, 8)])  # the reference's golden runs, simulated_run.rs:46-93
def test_loop_until_with_csv_path_writes_the_data_files(oracle, tmp_path, seed, N):
    from librabft_simulator_b200 import GlobalTime, RandomDelay, Simulator, format_round_switches_csv
    out = str(tmp_path / "results")
    contexts = Simulator.new(seed, N, RandomDelay.new(10.0, 4.0), None).loop_until(GlobalTime(1000).value, out)
    ref = oracle.run([seed], N)
    assert [len(c.committed_history()) for c in contexts] == ref.commit_counts[0].tolist()
    assert sorted(os.listdir(out)) == ["number_of_messages.txt", "round_switches.txt"]
    want = format_round_switches_csv(N, oracle.round_switches([seed], N, 0))
    assert open(os.path.join(out, "round_switches.txt")).read() == want
    assert open(os.path.join(out, "number_of_messages.txt")).read() == "%d\n" % int(ref.counters[0, :3].sum())


def test_csv_path_on_a_batch_is_refused():
    from librabft_simulator_b200 import BatchSimulator, RandomDelay
    with BatchSimulator([1, 2], 4, RandomDelay.new(10.0, 4.0)) as sim:
        with pytest.raises(ValueError, match="write_data_files"):
            sim.loop_until(1000, "/tmp/never_created")
