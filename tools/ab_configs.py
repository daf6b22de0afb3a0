#!/usr/bin/env python
"""Isolated A/B of the generic (non-bench) kernels: time a few BASELINE parity configs against each library given on
the command line, alternating, through the plain C ABI (works with older builds of the library too).
Usage: python tools/ab_configs.py [--rounds R] A.so B.so [C.so ...]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from librabft_simulator_b200._lib import LbftTiming  # noqa: E402
from tests.support import make_config  # noqa: E402

CONFIGS = [
    ("config2 1024x4 uniform[5,15]  <16,2,0>", 1024, 4, {"delay_kind": 1, "delay_lo": 5, "delay_hi": 15}),
    ("16384x5 LogNormal(10,4)       <16,1,0>", 16384, 5, {}),
    ("config5 16384x7 partitions    <16,3,0>", 16384, 7, {"partition_windows": 4, "partition_max_len": 150}),
    ("1024x20 LogNormal(10,4)       <32,3,0>", 1024, 20, {}),
]


def time_one(lib, I, N, kw, reps=3):
    cfg, keep = make_config(np.arange(52, 52 + I, dtype=np.uint64), N, 1000, **kw)
    h = ctypes.c_void_p()
    assert lib.lbft_create(ctypes.byref(cfg), ctypes.byref(h)) == 0, lib.lbft_last_error()
    best = 1e30
    for _ in range(reps + 1):
        assert lib.lbft_run(h) == 0, lib.lbft_last_error()
        t = LbftTiming()
        lib.lbft_timing_info(h, ctypes.byref(t))
        best = min(best, t.sim_ms)
    states = np.zeros((I, N), np.uint64)
    lib.lbft_last_states(h, ctypes.c_void_p(states.ctypes.data))
    lib.lbft_destroy(h)
    return best, int(np.bitwise_xor.reduce(states.reshape(-1)))


args = sys.argv[1:]
rounds = 2
if args and args[0] == "--rounds":
    rounds, args = int(args[1]), args[2:]
libs = []
for path in args:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.lbft_last_error.restype = ctypes.c_char_p
    lib.lbft_create.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    for fn in ("lbft_run", "lbft_destroy"):
        getattr(lib, fn).argtypes = [ctypes.c_void_p]
    lib.lbft_timing_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(LbftTiming)]
    lib.lbft_last_states.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    libs.append((os.path.basename(path), lib))
for name, I, N, kw in CONFIGS:
    res = {n: [] for n, _ in libs}
    digests = set()
    for _ in range(rounds):
        for n, lib in libs:
            ms, dig = time_one(lib, I, N, kw)
            res[n].append(ms)
            digests.add(dig)
    print("%s   (same results everywhere: %s)" % (name, len(digests) == 1))
    for n, _ in libs:
        print("    %-28s %s ms" % (n, " / ".join("%.2f" % v for v in res[n])))
