"""Random-configuration generator shared by the CPU (host-compiled core) and GPU fuzz parity tests."""
import numpy as np


def random_config(rng):
    n = int(rng.integers(2, 13))
    kw = {}
    if rng.random() < 0.3:
        lo = int(rng.integers(0, 8))
        kw.update(delay_kind=1, delay_lo=lo, delay_hi=lo + int(rng.integers(0, 12)))
    else:
        mean = float(rng.uniform(4.0, 30.0))
        kw.update(delay_mean=mean, delay_variance=float(rng.uniform(0.0, mean * mean * 0.3)))
    kw["delta"] = int(rng.integers(5, 60))
    kw["gamma"] = float(rng.choice([1.0, 1.5, 2.0, 2.5]))
    kw["lambda_"] = float(rng.choice([0.25, 0.5, 1.0]))
    kw["target_commit_interval"] = int(rng.choice([150, 400, 100000]))
    if rng.random() < 0.3:
        kw["voting_rights"] = [int(x) for x in rng.integers(1, 5, size=n)]
    if rng.random() < 0.3 and n >= 4:
        silent = np.zeros(n, dtype=np.uint8)
        silent[rng.choice(n, size=int(rng.integers(1, max(2, n // 3))), replace=False)] = 1
        kw["silent"] = [int(x) for x in silent]
    if rng.random() < 0.3:
        kw["partition_windows"] = int(rng.integers(1, 5))
        kw["partition_max_len"] = int(rng.integers(10, 200))
    max_clock = int(rng.choice([300, 600, 1000, 1500]))
    seed0 = int(rng.integers(1, 1 << 40))
    if seed0 % 6 == 0:   # epoch changes (node.rs:329-348); derived from seed0 so that the generator's stream is unchanged
        kw["commands_per_epoch"] = [3, 7, 15, 40][(seed0 // 6) % 4]
    return n, max_clock, seed0, kw


def random_modes(rng, max_clock):
    """A random combination of the opt-in modes: (flags, stops).  flags: 1 record round switches, 2 resumable; stops is
    None for a one-shot run, else the clocks of successive loop_until calls (now and then repeated or decreasing,
    sometimes ending before the horizon)."""
    flags = int(rng.choice([1, 2, 2, 3, 3]))
    if not flags & 2:
        return flags, None
    k = int(rng.integers(2, 6))
    stops = sorted(int(x) for x in rng.integers(0, max_clock + 1, size=k))
    if rng.random() < 0.3:
        i = int(rng.integers(1, k))
        stops[i] = max(0, stops[i - 1] - int(rng.integers(0, 3)))   # repeated / slightly smaller clock
    if rng.random() < 0.7:
        stops[-1] = max_clock
    return flags, stops


BIG_CAPS = {"round_cap": 1024, "queue_cap": 8192, "payload_cap": 2048}
CAPACITY_BITS = 2 | 4 | 8 | 128
