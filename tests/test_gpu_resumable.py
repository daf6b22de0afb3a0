"""SURVEY 8(f).4 on the device: lbft_run_until / lbft_snapshot_* through the C ABI against the oracle's loop_until
called once per stop on the same Simulator (simulator.rs:380-475), the dropped-event exit (:383-391) included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (num_nodes, horizon, instances, extra config, oracle kwargs)
CASES = [
    (3, 1000, 70, dict(queue_cap=64), dict(queue_cap=64)),        # shared-memory queue, spilled between launches
    (4, 1000, 96, {}, {}),                                         # HBM scan queue, three tiles
    (8, 1000, 40, {}, {}),                                         # calendar queue
    (7, 4500, 33, {}, {}),                                         # binary heap
    (40, 300, 3, {}, {}),                                          # two-word author masks
]
SCHEDULES = [[300, 650, 1000], [17, 400, 399, 1000], [0, 1, 2, 500, 500, 501, 1000]]


def check(ref, res, what):
    np.testing.assert_array_equal(ref.last_states, res.last_committed_states, err_msg=what)
    np.testing.assert_array_equal(ref.commit_counts, res.commit_counts, err_msg=what)
    np.testing.assert_array_equal(ref.counters[:, :8], res.counters[:, :8], err_msg=what)


@pytest.mark.parametrize("N,horizon,count,kw,okw", CASES)
def test_run_until_matches_the_staged_oracle_at_every_stop(oracle, N, horizon, count, kw, okw):
    from librabft_simulator_b200 import BatchSimulator, RandomDelay
    seeds = np.arange(6000 + 11 * N, 6000 + 11 * N + count, dtype=np.uint64)
    with BatchSimulator(seeds, N, RandomDelay.new(10.0, 4.0), resumable=True, **kw) as sim:
        sim.create(horizon)
        for schedule in SCHEDULES:
            stops = [t * horizon // 1000 for t in schedule]
            sim.set_seeds(seeds)                                   # start over: the next run_until is Simulator::new
            for k, stop in enumerate(stops, 1):
                res = sim.run_until(stop)
                check(oracle.run_staged(seeds, N, stops[:k], horizon, **okw), res, "after %s" % stops[:k])
                assert (res.counters[:, 11] == 0).all()


def test_snapshot_round_trip_into_a_second_handle(oracle):
    from librabft_simulator_b200 import BatchSimulator, RandomDelay, _lib
    seeds = np.arange(7100, 7100 + 80, dtype=np.uint64)
    delay = RandomDelay.new(10.0, 4.0)
    with BatchSimulator(seeds, 4, delay, resumable=True) as a, BatchSimulator(seeds + np.uint64(1), 4, delay, resumable=True) as b, \
            BatchSimulator(seeds, 5, delay, resumable=True) as other:
        a.create(1000)
        b.create(1000)           # same configuration, different seeds: the seeds are not needed to continue
        other.create(1000)
        with pytest.raises(_lib.LbftError):
            a.snapshot()         # nothing has run yet
        a.run_until(300)
        snap = a.snapshot()
        assert snap.nbytes > 80 * 4 * 200
        a.run_until(650)         # a moves on ...
        b.restore(snap)          # ... b continues from the checkpoint
        with pytest.raises(_lib.LbftError):
            other.restore(snap)  # differently configured
        with pytest.raises(_lib.LbftError):
            b.restore(snap[:-8])
        res_b = b.run_until(650)
        check(oracle.run_staged(seeds, 4, [300, 650]), res_b, "restored handle after [300, 650]")
        res_b = b.run_until(1000)
        res_a = a.run_until(1000)
        ref = oracle.run_staged(seeds, 4, [300, 650, 1000])
        check(ref, res_a, "original handle")
        check(ref, res_b, "restored handle")
        assert b.commit_log(5, 2) == a.commit_log(5, 2) == oracle.commit_log_staged(seeds, 4, 5, 2, [300, 650, 1000])


def test_run_until_errors_and_fresh_runs(oracle):
    from librabft_simulator_b200 import BatchSimulator, RandomDelay, _lib
    seeds = np.arange(1, 41, dtype=np.uint64)
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0)) as plain:
        plain.create(1000)
        for call in (lambda: plain.run_until(500), plain.snapshot):
            with pytest.raises(_lib.LbftError) as e:
                call()
            assert e.value.code == -3 and "LBFT_FLAG_RESUMABLE" in str(e.value)
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0), resumable=True) as sim:
        sim.create(1000)
        for bad in (-1, 1001):
            with pytest.raises(_lib.LbftError) as e:
                sim.run_until(bad)
            assert e.value.code == -1
        sim.run_until(400)
        one_shot = sim.run()     # lbft_run on a resumable handle: a fresh Simulator::new + loop_until(max_clock)
        check(oracle.run(seeds, 4), one_shot, "lbft_run after a staged run starts over")
        # ... and leaves a simulator that has run to max_clock: run_until continues IT (loop_until called again with a
        # smaller clock drops one more event and returns, simulator.rs:383-391)
        check(oracle.run_staged(seeds, 4, [1000, 400]), sim.run_until(400), "run_until after lbft_run continues that simulator")
        sim.set_seeds(seeds)     # a new staged run starts with fresh seeds
        check(oracle.run_staged(seeds, 4, [400]), sim.run_until(400), "run_until after lbft_set_seeds is Simulator::new")


def test_resumable_with_round_switch_recording(oracle):
    from librabft_simulator_b200 import BatchSimulator, RandomDelay
    seeds = np.arange(7300, 7300 + 40, dtype=np.uint64)
    stops = [300, 650, 1000]
    with BatchSimulator(seeds, 4, RandomDelay.new(10.0, 4.0), resumable=True, record_round_switches=True) as sim:
        sim.create(1000)
        for stop in stops:
            res = sim.run_until(stop)
        check(oracle.run_staged(seeds, 4, stops), res, "recording + resumable")
        for i in (0, 31, 32, 39):
            assert sim.round_switches(i) == oracle.round_switches_staged(seeds, 4, i, stops)


def test_simulator_loop_until_can_be_called_again_with_a_horizon(oracle):
    from librabft_simulator_b200 import GlobalTime, RandomDelay, Simulator
    sim = Simulator.new(52, 3, RandomDelay.new(10.0, 4.0), None, horizon=1000)
    first = [len(c.committed_history()) for c in sim.loop_until(GlobalTime(500).value)]
    second = sim.loop_until(GlobalTime(1000).value)
    ref1, ref2 = oracle.run_staged([52], 3, [500]), oracle.run_staged([52], 3, [500, 1000])
    assert first == ref1.commit_counts[0].tolist()
    assert [len(c.committed_history()) for c in second] == ref2.commit_counts[0].tolist()
    assert [c.last_committed_state() for c in second] == ref2.last_states[0].tolist()
    one_shot = Simulator.new(52, 3, RandomDelay.new(10.0, 4.0), None)
    assert [len(c.committed_history()) for c in one_shot.loop_until(1000)] == [27, 27, 27]   # the reference golden
    with pytest.raises(RuntimeError, match="horizon"):
        one_shot.loop_until(1200)
