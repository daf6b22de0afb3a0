"""Committed fixtures for the opt-in modes (tests/golden/modes.json, made by tests/golden/make_golden_modes.py):
round switches and staged runs.  These are ORACLE outputs — the reference holds no golden for DataWriter or for a
resumed simulator — kept so that the oracle and the device cannot drift together: the oracle (CPU), the host-compiled
core (CPU) and the kernel through the C ABI (GPU) must all keep reproducing them."""
import json
import os

import numpy as np
import pytest

from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "modes.json")))
IDS = [c["name"] for c in CASES]


def seeds_of(c):
    return np.arange(c["seed0"], c["seed0"] + c["count"], dtype=np.uint64)


def as_lists(switches):
    return [list(e) for e in switches]


def check_staged(st, commit_counts, last_states, counters):
    assert commit_counts.tolist() == st["commit_counts"]
    assert [[str(x) for x in row] for row in last_states.tolist()] == st["last_states"]
    assert counters[:, :8].tolist() == st["counters"]


def test_fixture_shape():
    assert len(CASES[0]["round_switches"][0]) == 111 and CASES[0]["round_switches"][0][0] == [0, 1, 10]   # seed 52: 37 rounds x 3 nodes
    assert all(len(c["staged"]) == 3 for c in CASES)


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_oracle_reproduces_mode_fixtures(oracle, c):
    seeds, N, mc, kw = seeds_of(c), c["nodes"], c["max_clock"], c["kw"]
    assert [int(x) for x in oracle.run(seeds, N, mc, **kw).counters[:, :3].sum(axis=1)] == c["message_count"]
    for i in range(c["count"]):
        assert as_lists(oracle.round_switches(seeds, N, i, mc, **kw)) == c["round_switches"][i]
    for st in c["staged"]:
        r = oracle.run_staged(seeds, N, st["stops"], mc, **kw)
        check_staged(st, r.commit_counts, r.last_states, r.counters)
        for i in range(c["count"]):
            assert as_lists(oracle.round_switches_staged(seeds, N, i, st["stops"], mc, **kw)) == st["round_switches"][i]


@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_hostcore_reproduces_mode_fixtures(hostcore, c):
    seeds, N, mc, kw = seeds_of(c), c["nodes"], c["max_clock"], c["kw"]
    for i in range(c["count"]):
        assert as_lists(hostcore.round_switches(seeds, N, i, mc, **kw)) == c["round_switches"][i]
    for st in c["staged"]:
        r = hostcore.run_staged(seeds, N, st["stops"], mc, flags=FLAG_RESUMABLE | FLAG_ROUND_SWITCHES, **kw)
        check_staged(st, r.commit_counts, r.last_states, r.counters)
        for i in range(c["count"]):
            assert as_lists(hostcore.round_switches_staged(seeds, N, i, st["stops"], mc, **kw)) == st["round_switches"][i]


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=IDS)
def test_gpu_reproduces_mode_fixtures(c):
    from tests.test_gpu_parity import make_sim
    seeds, N, mc, kw = seeds_of(c), c["nodes"], c["max_clock"], c["kw"]
    with make_sim(seeds, N, record_round_switches=True, **kw) as sim:
        res = sim.loop_until(mc)
        assert [int(x) for x in res.counters[:, :3].sum(axis=1)] == c["message_count"]
        for i in range(c["count"]):
            assert as_lists(sim.round_switches(i)) == c["round_switches"][i]
    with make_sim(seeds, N, record_round_switches=True, resumable=True, **kw) as sim:
        sim.create(mc)
        for st in c["staged"]:
            sim.set_seeds(seeds)
            for stop in st["stops"]:
                res = sim.run_until(stop)
            check_staged(st, res.commit_counts, res.last_committed_states, res.counters)
            for i in range(c["count"]):
                assert as_lists(sim.round_switches(i)) == st["round_switches"][i]
