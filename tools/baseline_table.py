"""All five BASELINE.json configs in one run: this package on the GPU next to the CPU
oracle port of the reference path on the box's host cores.

    python tools/baseline_table.py > profiles/rNN_configs_table.md

CPU legs are bounded samples (a few seconds each) of the same workload, extrapolated per
resample as BASELINE.md does; where the reference's numpy BCa jackknife is O(N^2) the
'pi' path is timed instead and that is said.  Development / reporting aid, not the bench.
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import torch  # noqa: E402

import bootstrap_oracle as orc  # noqa: E402
import scikits_bootstrap_b200 as boot  # noqa: E402

warnings.simplefilter("ignore")


def gpu_time(fn, reps=3):
    fn(0)
    best = 1e9
    for i in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn(i + 1)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


def cpu_rate_loop(tdata, stat, budget=4.0):
    """numpy path (bootstrap.py:429-432): resamples per second, single thread."""
    n = len(tdata[0])
    rng = np.random.default_rng(0)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        idx = rng.integers(0, n, n)
        stat(*(x[idx] for x in tdata))
        done += 1
    return done / (time.perf_counter() - t0), done


def main():
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    kernels = orc.numba_kernels()
    threads = 1
    if kernels:
        import numba
        threads = numba.get_num_threads()
        kernels[0](np.zeros(100), 2 * threads, 1)
    rows = []

    # cfg1
    x1 = np.random.default_rng(0).standard_normal(1000)
    x1d = dev(x1)
    g = gpu_time(lambda s: boot.ci(x1d, np.mean, method="bca", n_samples=10000, seed=s))
    t0 = time.perf_counter()
    ref = orc.ci(x1, np.mean, n_samples=10000, seed=0)
    c = time.perf_counter() - t0
    ours = boot.ci(x1d, np.mean, method="bca", n_samples=10000, seed=0)
    rows.append(("cfg1", "N=1e3 mean BCa n=1e4", 1e7, g, 1e7 / c, "numpy port, full run, 1 thread",
                 f"ours {np.round(ours, 5)} vs PCG64 {np.round(ref, 5)}"))

    # cfg2
    x2 = np.random.default_rng(1).standard_normal(100000)
    x2d = dev(x2)
    g = gpu_time(lambda s: boot.ci(x2d, np.average, method="pi", n_samples=100000, seed=s))
    rps, done = cpu_rate_loop((x2,), np.average)
    note = f"numpy port {done} resamples, 1 thread"
    cpu = rps * 1e5
    if kernels:
        t0 = time.perf_counter()
        kernels[0](x2, 4000, 7)
        cn = 4000 * 1e5 / (time.perf_counter() - t0)
        note += f"; numba port {cn:.2e} pts/s on {threads} threads"
        cpu = max(cpu, cn)
    rows.append(("cfg2", "N=1e5 average pi n=1e5", 1e10, g, cpu, note, ""))

    # cfg3
    a = np.random.default_rng(2).gamma(2, 1, 10**6)
    b = np.random.default_rng(3).gamma(3, 1, 10**6)
    ad, bd = dev(a), dev(b)
    g = gpu_time(lambda s: boot.ci((ad, bd), boot.ratio_of_means, method="bca", n_samples=100000, seed=s,
                                   multi="paired"))
    rps, done = cpu_rate_loop((a, b), lambda u, v: np.mean(u) / np.mean(v))
    rows.append(("cfg3", "paired N=1e6 ratio BCa n=1e5", 1e11, g, rps * 1e6,
                 f"numpy port 'pi' loop, {done} resamples, 1 thread; numpy BCa jackknife is O(N^2): infeasible",
                 ""))

    # cfg4
    x4 = np.random.default_rng(4).standard_normal((1000, 10000))
    x4d = dev(x4)
    g = gpu_time(lambda s: boot.ci(x4d, boot.median_axis0, method="bca", n_samples=10000, seed=s))
    gp = gpu_time(lambda s: boot.pval(x4d, boot.median_axis0, n_samples=10000, seed=s))
    rps, done = cpu_rate_loop((x4,), lambda t: np.median(t, axis=0), budget=6.0)
    rows.append(("cfg4", "1000 x 10000 median BCa n=1e4 (+pval)", 1e11, g, rps * 1e7,
                 f"numpy port 'pi' loop, {done} resamples, 1 thread; pval on GPU {gp * 1e3:.1f} ms", ""))

    # cfg5
    x5 = np.random.default_rng(5).standard_normal(10**7) + 1.0
    x5d = dev(x5)
    g = gpu_time(lambda s: boot.ci(x5d, np.mean, method="bca", n_samples=100000, seed=s), reps=2)
    if kernels:
        t0 = time.perf_counter()
        kernels[0](x5, 8 * threads, 3)
        kernels[1](x5)
        cpu = 8 * threads * 1e7 / (time.perf_counter() - t0)
        note = f"numba port incl. O(N) jackknife, {8 * threads} resamples on {threads} threads"
    else:
        rps, done = cpu_rate_loop((x5,), np.mean)
        cpu, note = rps * 1e7, f"numpy port 'pi' loop, {done} resamples, 1 thread"
    rows.append(("cfg5", "N=1e7 mean BCa n=1e5", 1e12, g, cpu, note, ""))

    print("| config | workload | GPU time (1 x B200, data resident) | GPU points/s | CPU port points/s | GPU / CPU | CPU leg | note |")
    print("|---|---|---|---|---|---|---|---|")
    for name, wl, pts, g, cpu, how, note in rows:
        print(f"| {name} | {wl} | {g * 1e3:.2f} ms | {pts / g:.3e} | {cpu:.3e} | {pts / g / cpu:.0f}x | {how} | {note} |")
    print(f"\nhost: os.cpu_count() = {os.cpu_count()}, numba threads = {threads}; "
          f"GPU: {torch.cuda.get_device_name(0)}; points = n_samples x N x columns")


if __name__ == "__main__":
    main()
