"""A plain C program (tests/cabi/golden_run.c) calling the C ABI reproduces the reference's golden runs — the
boundary works without Python or torch in the process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_caller(tmp_path):
    from librabft_simulator_b200 import _build
    _build.build_product()
    exe = str(tmp_path / "golden_run")
    libdir = os.path.dirname(_build.LIB_PATH)
    subprocess.run(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cabi", "golden_run.c"), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-llbft_b200", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_c_caller_compiles_and_links(tmp_path):
    # CPU box: the C program builds against include/lbft.h and links the product library
    exe = build_c_caller(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_caller_reproduces_goldens(tmp_path):
    exe = build_c_caller(tmp_path)
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0, p.stdout
    assert "golden runs reproduced" in p.stdout
