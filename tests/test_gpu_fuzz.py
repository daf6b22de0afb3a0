"""Fuzz parity on the GPU: the same random configurations as tests/test_fuzz_hostcore.py, through the C ABI."""
import numpy as np
import pytest

from tests.fuzz_configs import random_config, random_modes
from tests.test_fuzz_hostcore import check, check_modes
from tests.test_gpu_parity import GpuResult, gpu_run, make_sim

pytestmark = pytest.mark.gpu


def gpu_runner(seeds, n, max_clock, **kw):
    sim, res = gpu_run(seeds, n, max_clock, strict=False, **kw)
    return res


def test_fuzz_gpu_vs_oracle(oracle, kernel_choice):
    rng = np.random.default_rng(20260922)
    reran = 0
    for _ in range(120):
        n, max_clock, seed0, kw = random_config(rng)
        reran += 1 if check(oracle, gpu_runner, n, max_clock, seed0, kw, count=12) else 0
    assert reran < 60  # most random configurations fit the automatically chosen capacities


def test_fuzz_modes_gpu_vs_oracle(oracle):
    """Random flag combinations (round-switch recording, resumable) and stop schedules, as in the CPU fuzz."""
    live = {}

    def run(seeds, n, max_clock, flags, stops, **kw):
        for old in live.values():
            old.close()
        sim = make_sim(seeds, n, record_round_switches=bool(flags & 1), resumable=bool(flags & 2), **kw)
        live["sim"] = sim
        if stops is None:
            return GpuResult(sim.loop_until(max_clock, strict=False))
        sim.create(max_clock)
        for stop in stops:
            res = sim.run_until(stop, strict=False)
        return GpuResult(res)

    def switches(seeds, n, max_clock, flags, stops, i, **kw):
        return live["sim"].round_switches(i)   # of the run `run` has just made with these very arguments

    rng = np.random.default_rng(20260923)
    for _ in range(60):
        n, max_clock, seed0, kw = random_config(rng)
        kw.pop("commands_per_epoch", None)   # epochs x recording / staged runs: host-compiled core only (tests/test_epochs.py)
        flags, stops = random_modes(rng, max_clock)
        check_modes(oracle, run, switches, n, max_clock, seed0, kw, flags, stops, count=12)
    for old in live.values():
        old.close()
