#!/usr/bin/env python
"""bench.py — resampled points/sec for BCa ci(np.mean) (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], "cfg5"): 1-D float64, N = 1e7 rows,
`ci(x, np.mean, method='bca', n_samples=1e5)`; a step is one whole ci() call =
n_samples x N resampled points.  With N > 1 GPUs (torchrun, one rank per GPU,
NCCL) the 1e5 resample ids are sharded over the ranks (strong scaling: the
total work is the config's, fixed); the only collectives are the integer
all-reduces of the z0 count and the radix-select histograms.

One JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline     K1 (fused draw+gather+reduce): algorithmic bytes = 8 B per resampled
               point, duration from CUDA events recorded by the library around the
               K1 launch on the launching stream; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline oracle port of the reference's compiled CPU path (numba prange over
               resample ids, bootstrap.py:639-664) on a bounded sample, rank 0, N=1.
  e2e          the same ci() call fed a pinned HOST tensor: H2D of the data and D2H
               of the results are inside the timed region.
`--impl reference` times the CPU port alone (all host threads) on the same config.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = 10**7
N_SAMPLES = 10**5
DATA_SEED = 5
BYTES_PER_POINT = 8            # one float64 operand per resampled point
METRIC = "resampled points/sec (n_samples x N), BCa ci(np.mean)"
UNIT = "points/s"


def make_data():
    return np.random.default_rng(DATA_SEED).standard_normal(N_ROWS) + 1.0


def workload_config(extra=None):
    cfg = {
        "workload": "cfg5: ci(x[1e7] f64, np.mean, method='bca', n_samples=1e5), resamples sharded over ranks",
        "n_rows": N_ROWS, "n_samples": N_SAMPLES, "alpha": 0.05,
        "l2": "L2 flushed between timed steps (256 MiB write)",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
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
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU port

def cpu_port_rate(x, budget_s=12.0, min_resamples=None):
    """Times the oracle port of the reference's CPU path for this config on a bounded
    sample.  numba available: the prange kernels (reference use_numba=True path,
    bootstrap.py:639-664), all threads.  Otherwise the numpy loop (bootstrap.py:429-432),
    one thread, method='pi' (the numpy BCa jackknife is O(N^2): infeasible at N=1e7)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bootstrap_oracle as orc
    kernels = orc.numba_kernels()
    n = len(x)
    if kernels is not None:
        import numba
        boot_mean, jack_mean = kernels
        threads = numba.get_num_threads()
        boot_mean(x[:1000], 2 * threads, 1)       # JIT warm-up
        jack_mean(x[:1000])
        batch = max(threads, min_resamples or 0)
        done, t0 = 0, time.perf_counter()
        stats = []
        while True:
            stats.append(boot_mean(x, batch, 1000 + done))
            done += batch
            if time.perf_counter() - t0 > budget_s:
                break
        js = jack_mean(x)
        stat = np.sort(np.concatenate(stats))
        acc = orc.acceleration(js)
        av = orc.bca_avals(stat, x.mean(), acc, np.array([0.025, 0.975]), len(stat))
        orc.order_indices(av, len(stat))
        dt = time.perf_counter() - t0
        return {"value": done * n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{done} of {N_SAMPLES} resamples at N={n}, numba prange port of "
                          f"bootstrap.py:639-664 incl. O(N) jackknife; {dt:.1f} s"}, dt, done
    done, t0 = 0, time.perf_counter()
    rng = np.random.default_rng(0)
    vals = []
    while True:
        idx = rng.integers(0, n, n)
        vals.append(x[idx].mean())
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done * n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{done} of {N_SAMPLES} resamples at N={n}, numpy loop port of "
                      f"bootstrap.py:429-432 (method='pi'; numpy BCa jackknife is O(N^2)); {dt:.1f} s"}, dt, done


def run_reference_arm(args):
    """--impl reference: the CPU implementation alone, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x = make_data()
    step_budget = 8.0
    for _ in range(args.warmup):
        cpu_port_rate(x, budget_s=0.5)
    rates, times = [], []
    info = None
    for _ in range(args.steps):
        info, dt, _ = cpu_port_rate(x, budget_s=step_budget)
        rates.append(info["value"])
        times.append(dt)
    value = float(np.mean(rates))
    info["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config({"cpu_step": "bounded sample, see cpu_baseline.sample"}),
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
# version banner on the first collective when NCCL_DEBUG is set on the box), so file
# descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# ----------------------------------------------------------------------------- GPU arm

def main():
    global N_ROWS, N_SAMPLES
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-samples", type=int, default=N_SAMPLES, help=argparse.SUPPRESS)
    ap.add_argument("--n-rows", type=int, default=N_ROWS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    N_ROWS, N_SAMPLES = args.n_rows, args.n_samples
    args.warmup = max(args.warmup, 3)      # timing rule: at least 3 warm-up steps (reported as run)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import scikits_bootstrap_b200 as boot
    from scikits_bootstrap_b200 import _abi, _device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        boot.distributed.enable()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = _abi.lib()
    x_host = torch.from_numpy(make_data()).pin_memory()
    x_dev = x_host.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    points = float(N_ROWS) * float(N_SAMPLES)

    def step(data, seed):
        return boot.ci(data, np.mean, alpha=0.05, n_samples=N_SAMPLES, method="bca", seed=seed)

    for w in range(args.warmup):
        step(x_dev, 1000 + w)
    barrier()

    # ---- device-resident timing (value) -----------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _abi.check(L.b200boot_profile_enable(1))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    k1_ms = []
    launches0 = L.b200boot_launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    result = None
    for i in range(args.steps):
        flush.zero_()                      # L2 flush, outside the event-timed region
        ev[i][0].record()
        result = step(x_dev, i)
        ev[i][1].record()
        cms = ctypes.c_float()
        _abi.check(L.b200boot_profile_last_k1_ms(ctypes.byref(cms)))
        k1_ms.append(cms.value)
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = L.b200boot_launch_count() - launches0
    _abi.check(L.b200boot_profile_enable(0))
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end timing from pinned host memory (e2e) -------------------------
    for w in range(2):                     # untimed: the host-input path has its own one-time costs
        step(x_host, 2000 + w)             # (copy stream, staging block of the caching allocator)
    _device.TRAFFIC["h2d"] = _device.TRAFFIC["d2h"] = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(x_host, i)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = _device.TRAFFIC["h2d"] // max(args.steps, 1)
    d2h = _device.TRAFFIC["d2h"] // max(args.steps, 1)

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3, float(np.mean(k1_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k1_mean_ms = (float(v) for v in t.tolist())

    if rank == 0:
        value = points * args.steps / (dev_ms * 1e-3)
        e2e_value = points * args.steps / (e2e_ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        local_points = points / world
        achieved = local_points * BYTES_PER_POINT / (k1_mean_ms * 1e-3) / 1e9
        traffic, ncu = None, None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            prof = json.load(open(tpath))
            traffic = prof.get("dram_bytes_per_launch")
            # the kernel serves rows from shared memory: what binds it, from the committed
            # ncu capture (profiles/), not re-measured here
            ncu = {k: prof[k] for k in ("issue_slots_active_pct", "smem_wavefront_pipe_pct", "alu_pipe_pct",
                                        "fp64_pipe_pct", "points_per_clk_per_sm", "source") if k in prof}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"parallelism": f"resample-shard x{world}", "result": [float(v) for v in result]}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": "k1_resample_kernel<kMomSum>",
                         "kernel_ms": k1_mean_ms, "peak_source": peak_src, "ncu_shared_memory_regime": ncu,
                         "note": "algorithmic 8 B per resampled point; rows are staged in shared memory, so "
                                 "DRAM traffic is far below the algorithmic bytes and frac may exceed 1"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": wall_s,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_port_rate(make_data(), budget_s=12.0)[0]
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
