#!/usr/bin/env python
"""bench.py — simulated consensus rounds/sec of the batched LibraBFTv2 event loop (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: `Simulator::new` + `loop_until(max_clock)` for every
instance of the batch (65 536 instances x 4 authors per GPU, LogNormal(10, 4) delays, max_clock = 1000 —
BASELINE.json configs[2]/SURVEY §8d.3).  Each step uses fresh seeds.

  value     rounds/s with the seeds already resident in HBM (kernel only, CUDA events, max over ranks)
  e2e       rounds/s through the public API with HOST buffers: seeds host->device, kernel, summaries
            device->host (commit counts, state keys, counters, status), every step
  roofline  algorithmic bytes (SURVEY §8d formula over the run's own event counters) / kernel time vs the
            measured HBM copy peak
  cpu_baseline  the CPU oracle (oracle/, a C++ restatement pinned by the reference goldens — no Rust toolchain
            exists here) timed on the host cores on a bounded sample of the same workload

`--impl reference` times that CPU implementation alone (all host threads) on the same metric/config.
Multi-GPU: one process per GPU under torchrun, instances sharded with no data-path collective; one NCCL
all-gather of the per-instance commit counts at the end of every step ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INSTANCES_PER_GPU = 65536
NODES = 4
MAX_CLOCK = 1000
BASE_SEED = 52
METRIC = "simulated consensus rounds/sec (whole box)"
UNIT = "rounds/s"


def algorithmic_bytes(counters, n_nodes):
    """Bytes = P*(2*S_node + S_hdr) + P_notify*S_notif + Q*S_hdr + Q_notify*S_notif."""
    s_hdr, s_node, s_notif = 16, 208 + 12 * n_nodes, 40 + 10 * n_nodes
    c = counters.astype(np.float64)
    p = c[:, 0:4].sum()
    p_notify = c[:, 0].sum()
    q = c[:, 5].sum()
    q_notify = c[:, 9].sum()
    return p * (2 * s_node + s_hdr) + p_notify * s_notif + q * s_hdr + q_notify * s_notif


def effective_cores():
    """Host threads we can actually use: CPU affinity capped by the cgroup CPU quota (the GPU boxes expose 128
    logical CPUs but grant 16 CPUs of quota; oversubscribing only adds throttling)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def ncu_traffic_bytes(per_gpu):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the event-loop kernel, from the committed ncu
    --set full capture of this very command (profiles/traffic.json); None if the workload differs."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("instances") == per_gpu and t.get("nodes") == NODES and t.get("max_clock") == MAX_CLOCK:
            return t["dram_bytes_per_launch"], t.get("source")
    except Exception:
        pass
    return None, None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smmax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def step_seeds(step, rank, per_gpu):
    # instance i of the whole job uses seed BASE_SEED + i; every step moves on to a fresh block of seeds
    world_block = step * (1 << 24)
    first = BASE_SEED + world_block + rank * per_gpu
    return np.arange(first, first + per_gpu, dtype=np.uint64)


def run_cpu_sample(oracle, seeds, threads):
    t0 = time.perf_counter()
    res = oracle.run(seeds, NODES, MAX_CLOCK, threads=threads)
    dt = time.perf_counter() - t0
    rounds = float(res.counters[:, 6].sum())
    return rounds, dt, res


def cpu_baseline(per_gpu, target_seconds=12.0):
    """Time the CPU oracle on a bounded sample of the same workload, all host threads."""
    from tests.support import Oracle
    oracle = Oracle()
    threads = effective_cores()
    probe = step_seeds(0, 0, per_gpu)[: 64 * threads]
    r, dt, _ = run_cpu_sample(oracle, probe, threads)
    per_inst = dt / len(probe)
    count = int(max(len(probe), min(per_gpu, target_seconds / max(per_inst, 1e-9))))
    sample = step_seeds(0, 0, per_gpu)[:count]
    rounds, dt, res = run_cpu_sample(oracle, sample, threads)
    return {"value": rounds / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d of %d instances x %d authors, max_clock=%d, %.1f s wall, %d threads; C++ oracle restating the "
                      "Rust reference (no Rust toolchain), pinned by the reference goldens" % (count, per_gpu, NODES, MAX_CLOCK, dt, threads)}, res, sample


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path (here: the C++ oracle port, because the
    Rust reference cannot be built in this image) on all host threads; rank 0 only."""
    if rank != 0:
        return
    from tests.support import Oracle
    oracle = Oracle()
    threads = effective_cores()
    per_gpu = args.instances
    # bounded sample per step so that the whole run ends within minutes
    probe = step_seeds(0, 0, per_gpu)[: 32 * threads]
    _, dt, _ = run_cpu_sample(oracle, probe, threads)
    budget = 120.0 / max(1, args.steps + args.warmup)
    count = int(max(len(probe), min(per_gpu, min(budget, 20.0) / max(dt / len(probe), 1e-9))))
    for w in range(args.warmup):
        run_cpu_sample(oracle, step_seeds(1000 + w, 0, per_gpu)[:count], threads)
    rounds_total, t_total = 0.0, 0.0
    for s in range(args.steps):
        r, dt, _ = run_cpu_sample(oracle, step_seeds(s, 0, per_gpu)[:count], threads)
        rounds_total += r
        t_total += dt
    value = rounds_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "%d instances x %d authors per GPU, LogNormal(10,4) delay, max_clock=%d (BASELINE configs[2]); "
                               "reference arm steps over a bounded sample of %d instances" % (per_gpu, NODES, MAX_CLOCK, count),
                   "seeds": "base_seed %d + instance" % BASE_SEED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d instances per step, %d threads; C++ oracle port of the Rust reference "
                                   "(reference not buildable here: no cargo/rustc)" % (count, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print on stdout (e.g. the NCCL version banner) to stderr, so that the ONE JSON line
    the driver parses is the only thing on stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--instances", type=int, default=INSTANCES_PER_GPU, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from librabft_simulator_b200 import BatchSimulator, RandomDelay

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    per_gpu = args.instances
    sim = BatchSimulator(step_seeds(0, rank, per_gpu), NODES, RandomDelay.new(10.0, 4.0), device=local_rank)
    sim.create(MAX_CLOCK)
    dev_bytes, words_per_inst = sim.memory_info()

    # view of the device-resident commit counts for the end-of-step NCCL all-gather (no host round trip)
    class _Cai:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    ptr, nbytes = sim.device_buffer(0)
    counts_dev = torch.as_tensor(_Cai(ptr, nbytes // 4), device="cuda:%d" % local_rank)
    gathered = torch.empty(world * counts_dev.numel(), dtype=counts_dev.dtype, device=counts_dev.device) if distributed else None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_counts():
        if distributed:
            dist.all_gather_into_tensor(gathered, counts_dev)

    def step_device(step):
        """value leg: seeds already resident in HBM, kernel only"""
        sim.set_seeds(step_seeds(step, rank, per_gpu))
        sim.upload()            # outside the timed region
        return step

    # ---------------- warm-up (both legs) ----------------
    for w in range(args.warmup):
        sim.set_seeds(step_seeds(10000 + w, rank, per_gpu))
        sim.run(strict=False)
        gather_counts()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- e2e leg: host buffers in, host results out, every step ----------------
    e2e_rounds = 0.0
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        sim.set_seeds(step_seeds(s, rank, per_gpu))       # host array (pinned staging inside the library)
        res = sim.run(strict=False)                       # H2D seeds + kernel + D2H summaries
        gather_counts()
        e2e_rounds += float(res.active_rounds.sum())
    barrier()
    e2e_seconds = time.perf_counter() - t0
    h2d_bytes, d2h_bytes = int(sim.timing.h2d_bytes), int(sim.timing.d2h_bytes)

    # ---------------- value leg: inputs resident in HBM, device-timed kernel ----------------
    kernel_ms, rounds_dev, counters_last, flagged = [], 0.0, None, 0
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        sim.set_seeds(step_seeds(100 + s, rank, per_gpu))
        sim.upload()
        sim.run_device()                                  # CUDA events on the launching stream around the kernel
        kernel_ms.append(float(sim.timing.sim_ms))
        gather_counts()
        res = sim.download(strict=False)
        rounds_dev += float(res.active_rounds.sum())
        counters_last = res.counters
        flagged += int(((res.status & 0xBE) != 0).sum())  # any LBFT_ST_ERROR_MASK bit (capacity / invariant / epoch)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    dev_seconds = sum(kernel_ms) / 1e3
    stats = torch.tensor([dev_seconds, e2e_seconds], dtype=torch.float64, device="cuda")
    sums = torch.tensor([rounds_dev, e2e_rounds, float(counters_last[:, 0:4].sum())], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)     # max over ranks
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)      # whole-job totals
    dev_seconds, e2e_seconds = stats.tolist()
    rounds_dev, e2e_rounds, events_last = sums.tolist()

    if rank == 0:
        value = rounds_dev / dev_seconds
        peak, peak_src = measured_peak_gbs()
        bytes_launch = algorithmic_bytes(counters_last, NODES)
        k_ms = float(np.mean(kernel_ms))
        achieved = bytes_launch / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic_bytes(per_gpu)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_seconds / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "%d instances x %d authors per GPU, LogNormal(10,4) delay, max_clock=%d, delta=20 gamma=2 "
                                   "lambda=0.5 (BASELINE configs[2], SURVEY 8d.3)" % (per_gpu, NODES, MAX_CLOCK),
                       "seeds": "base_seed %d + instance, fresh block per step" % BASE_SEED,
                       "instances_total": per_gpu * world,
                       "l2": "state working set %.0f MB per GPU > 126 MB L2; every step re-initialises it" % (dev_bytes / 1e6),
                       "parallelism": "instances sharded over %d GPU(s); one NCCL all-gather of commit counts per step" % world},
            "e2e": {"value": e2e_rounds / e2e_seconds, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": 1e3 * e2e_seconds / args.steps},
            "gpu_launches": args.steps * 2,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "lbft_event_loop_kernel<16,2,true>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": bytes_launch,
                         "events_per_s": events_last / (k_ms * 1e-3)},
            "clocks": clocks,
            "state_bytes_per_instance": words_per_inst * 4,
            "flagged_instances": flagged,  # instances (rank 0, value leg) that hit a capacity/invariant flag: expected 0
        }
        if not args.no_cpu_baseline:
            base, cres, sample = cpu_baseline(per_gpu)
            line["cpu_baseline"] = base
        emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
