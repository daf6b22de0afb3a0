"""The product's C-ABI library loads and exports every symbol include/lbft.h declares; argument validation
works without a GPU; and there is NO CPU fallback (creating a simulator without a device fails loudly)."""
import ctypes
import os
import re

import numpy as np
import pytest

from librabft_simulator_b200 import _build, _lib
from tests.support import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build_product()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "lbft.h")).read()
    declared = set(re.findall(r"\b(lbft_[a-z_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_size(lib):
    assert lib.lbft_abi_version() == 1
    assert ctypes.sizeof(_lib.LbftConfig) == 152
    assert ctypes.sizeof(_lib.LbftCommit) == 16
    assert ctypes.sizeof(_lib.LbftRoundSwitch) == 16


def test_invalid_configs_are_rejected(lib):
    h = ctypes.c_void_p()
    cfg, keep = make_config([1, 2], 4)
    cfg.struct_size = 12
    assert lib.lbft_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"struct_size" in lib.lbft_last_error()
    for field, value in (("num_nodes", 0), ("num_nodes", 65), ("max_clock", -1), ("delay_kind", 7), ("flags", 8), ("flags", 0x80000001), ("flags", 4 | 2),
                         ("commands_per_epoch", 0), ("delay_mean", -1.0), ("delta", 0)):
        cfg, keep = make_config([1, 2], 4)
        setattr(cfg, field, value)
        assert lib.lbft_create(ctypes.byref(cfg), ctypes.byref(h)) == -1, field
        assert h.value is None
    assert lib.lbft_create(None, ctypes.byref(h)) == -1


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    cfg, keep = make_config([1, 2, 3], 4)
    rc = lib.lbft_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2, "without a GPU the product must fail with LBFT_ERR_CUDA, not fall back to the CPU"
    assert b"no CPU fallback" in lib.lbft_last_error()


def test_product_does_not_link_the_oracle():
    # the product sources never #include / import anything from oracle/ or tests/
    pkg = os.path.join(ROOT, "librabft_simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".cu", ".cuh", ".h", ".hpp", ".py")):
                continue
            for line in open(os.path.join(dirpath, f)):
                t = line.strip()
                if t.startswith("#include"):
                    assert "oracle" not in t and "tests/" not in t and "hostcore" not in t, (f, t)
                if f != "_build.py" and (t.startswith("import ") or t.startswith("from ")):
                    assert "oracle" not in t and "tests" not in t.split(), (f, t)
    # ... and the shared object does not carry oracle symbols
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _build.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "lbfo_" not in syms and "lbft_oracle" not in syms and "hostcore" not in syms


def test_results_read_after_close_fail_clearly(lib):
    """BatchResult fetches lazily from the handle: after close() that is a clear Python error, not a NULL handle in C."""
    from librabft_simulator_b200 import BatchSimulator, RandomDelay
    sim = BatchSimulator([1, 2], 4, RandomDelay.new(10.0, 4.0))
    with pytest.raises(RuntimeError, match="closed"):
        sim._fetch("lbft_commit_counts", np.uint32, (2, 4))


def test_every_export_is_placed_against_the_reference_in_integration_md():
    """INTEGRATION.md says, for each C entry point, what it replaces in the reference (or that it is new)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [name for name in _lib.EXPORTS if name not in doc] == []


def test_python_mirror_maps_its_options_onto_config_flags(lib):
    """Host logic of the mirror, no device needed: record_round_switches / resumable / horizon -> lbft_config.flags."""
    from librabft_simulator_b200 import BatchSimulator, RandomDelay, Simulator
    delay = RandomDelay.new(10.0, 4.0)
    want = {(False, False): 0, (True, False): _lib.FLAG_ROUND_SWITCHES, (False, True): _lib.FLAG_RESUMABLE,
            (True, True): _lib.FLAG_ROUND_SWITCHES | _lib.FLAG_RESUMABLE}
    for (rec, res), flags in want.items():
        cfg = BatchSimulator([1, 2, 3], 4, delay, record_round_switches=rec, resumable=res).make_config(1000)
        assert cfg.flags == flags and cfg.max_clock == 1000 and cfg.num_instances == 3 and cfg.struct_size == 152
    assert Simulator.new(52, 3, delay, None)._batch.resumable is False
    sim = Simulator.new(52, 3, delay, None, horizon=1000)
    assert sim._batch.resumable is True and sim._batch.make_config(sim._horizon).flags == _lib.FLAG_RESUMABLE
    with pytest.raises(ValueError, match="write_data_files"):       # csv_path names one simulator's directory
        BatchSimulator([1, 2], 4, delay).loop_until(1000, "/nonexistent/never_created")
