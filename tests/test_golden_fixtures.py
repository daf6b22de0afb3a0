"""Committed golden fixtures (tests/golden/commit_logs.json, made by tests/golden/make_golden.py with the oracle):
the oracle must keep reproducing them (CPU), the host-compiled device core too (CPU), and the kernel through the
C ABI (GPU) — full commit logs, state keys and event counters."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "commit_logs.json")))


def seeds_of(c):
    return np.arange(c["seed0"], c["seed0"] + c["count"], dtype=np.uint64)


def check_summary(c, commit_counts, last_states, counters):
    assert commit_counts.tolist() == c["commit_counts"], c["name"]
    assert [[str(x) for x in row] for row in last_states.tolist()] == c["last_states"], c["name"]
    assert counters[:, :8].tolist() == c["counters"], c["name"]


def test_reference_values_are_in_the_fixture():
    assert CASES[0]["commit_counts"] == [[27, 27, 27]] and CASES[0]["last_states"] == [["11134312813757838303"] * 3]
    assert CASES[1]["commit_counts"] == [[28] * 7 + [30]]
    assert CASES[1]["last_states"] == [["12785928431398617538"] * 7 + ["4890275890002623733"]]


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_fixture(oracle, c):
    r = oracle.run(seeds_of(c), c["nodes"], c["max_clock"], **c["kw"])
    check_summary(c, r.commit_counts, r.last_states, r.counters)
    log = oracle.commit_log(seeds_of(c), c["nodes"], 0, 0, c["max_clock"], **c["kw"])
    assert [list(x) for x in log] == c["logs"][0][0]
    assert oracle.state_key(log) == int(c["last_states"][0][0])


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_hostcore_reproduces_fixture(hostcore, c):
    r = hostcore.run(seeds_of(c), c["nodes"], c["max_clock"], **c["kw"])
    check_summary(c, r.commit_counts, r.last_states, r.counters)


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_gpu_reproduces_fixture(c):
    from tests.test_gpu_parity import gpu_run
    sim, g = gpu_run(seeds_of(c), c["nodes"], c["max_clock"], **dict(c["kw"]))
    check_summary(c, g.commit_counts, g.last_states, g.counters)
    for i in range(c["count"]):
        for n in range(c["nodes"]):
            assert [list(x) for x in sim.commit_log(i, n)] == c["logs"][i][n], (c["name"], i, n)
