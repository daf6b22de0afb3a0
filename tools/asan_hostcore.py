#!/usr/bin/env python
"""Builds the device state machine for the host (tests/hostcore) with AddressSanitizer + UBSan and drives it through
the case matrices of tests/test_round_switches.py and tests/test_resumable.py (recording, resumable, both, every stop
schedule, ragged tiles) plus plain runs up to 64 authors, comparing with the oracle as the tests do.  The state vector
is allocated exactly, so a write past an instance's appended regions (switch table, save area, spilled queue) on the
last tile is reported.  CPU only.  Usage: python tools/asan_hostcore.py   (re-executes itself under LD_PRELOAD)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "asan")          # scratch (git-ignored)
LIB = os.path.join(OUT, "libhostcore_asan.so")


def gxx_file(name):
    return subprocess.run(["g++", "-print-file-name=" + name], capture_output=True, text=True, check=True).stdout.strip()


if os.environ.get("LBFT_ASAN_CHILD") != "1":
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    "-DLBFT_CHECK_C1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", LIB,
                    os.path.join(ROOT, "tests", "hostcore", "hostcore.cpp")], check=True)
    env = dict(os.environ, LBFT_ASAN_CHILD="1", LD_PRELOAD=gxx_file("libasan.so") + ":" + gxx_file("libubsan.so"),
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1")
    sys.exit(subprocess.run([sys.executable, os.path.abspath(__file__)], env=env).returncode)

sys.path.insert(0, ROOT)
from librabft_simulator_b200 import _build  # noqa: E402

_build.build_hostcore = lambda force=False: LIB
import tests.test_resumable as tr  # noqa: E402
import tests.test_round_switches as ts  # noqa: E402
from tests.support import FLAG_RESUMABLE, FLAG_ROUND_SWITCHES, HostCore, Oracle, assert_same  # noqa: E402

o, h = Oracle(), HostCore()
n = 0
for N, mc, count, kw in ts.CASES:                       # recording
    seeds = list(range(100, 100 + min(count, 33)))      # 33: a ragged second tile
    assert_same(o.run(seeds, N, mc, **kw), h.run(seeds, N, mc, flags=FLAG_ROUND_SWITCHES, **kw))
    for i in (0, len(seeds) - 1):
        assert h.round_switches(seeds, N, i, mc, **kw) == o.round_switches(seeds, N, i, mc, **kw)
    n += 1
for N, horizon, count, kw, qmode in tr.CASES:            # resumable, and resumable + recording
    seeds = list(range(200, 200 + min(count, 33)))
    for sched in tr.SCHEDULES:
        stops = tr.scaled(sched, horizon)
        ref = o.run_staged(seeds, N, stops, horizon, **kw)
        assert_same(ref, h.run_staged(seeds, N, stops, horizon, **kw))
        assert_same(ref, h.run_staged(seeds, N, stops, horizon, flags=FLAG_RESUMABLE | FLAG_ROUND_SWITCHES, **kw))
    i = len(seeds) - 1
    assert h.round_switches_staged(seeds, N, i, stops, horizon, **kw) == o.round_switches_staged(seeds, N, i, stops, horizon, **kw)
    n += 1
for N in (3, 4, 8, 33, 64):                              # plain path
    seeds = list(range(1, 34 if N <= 8 else 3))
    mc = 1000 if N <= 8 else 150
    assert_same(o.run(seeds, N, mc), h.run(seeds, N, mc))
    n += 1
print("ASan + UBSan: clean over %d configurations" % n)
