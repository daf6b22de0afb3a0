"""Fuzz parity on the GPU: the same random configurations as tests/test_fuzz_hostcore.py, through the C ABI."""
import numpy as np
import pytest

from tests.fuzz_configs import random_config
from tests.test_fuzz_hostcore import check
from tests.test_gpu_parity import gpu_run

pytestmark = pytest.mark.gpu


def gpu_runner(seeds, n, max_clock, **kw):
    sim, res = gpu_run(seeds, n, max_clock, strict=False, **kw)
    return res


def test_fuzz_gpu_vs_oracle(oracle):
    rng = np.random.default_rng(20260922)
    reran = 0
    for _ in range(120):
        n, max_clock, seed0, kw = random_config(rng)
        reran += 1 if check(oracle, gpu_runner, n, max_clock, seed0, kw, count=12) else 0
    assert reran < 60  # most random configurations fit the automatically chosen capacities
