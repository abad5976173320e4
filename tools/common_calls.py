"""Latency of common small calls (development aid)."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
def t(name, f, reps=30):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(f"{name:60s} {1e6 * np.median(ts):9.1f} us")
for n in (100, 1000, 10000):
    x = rng.standard_normal(n); y = 0.5 * x + rng.standard_normal(n)
    t(f"ci mean   N={n} n=1e4", lambda: boot.ci(x, np.mean, n_samples=10000, seed=1))
    t(f"ci std    N={n} n=1e4", lambda: boot.ci(x, np.std, n_samples=10000, seed=1))
    t(f"ci median N={n} n=1e4", lambda: boot.ci(x, np.median, n_samples=10000, seed=1))
    t(f"ci median N={n} n=1e4 pi", lambda: boot.ci(x, np.median, n_samples=10000, seed=1, method="pi"))
    t(f"ci q90    N={n} n=1e4", lambda: boot.ci(x, boot.percentile(90), n_samples=10000, seed=1))
    t(f"ci pearson (x,y) N={n} n=1e4", lambda: boot.ci((x, y), boot.pearson_r, n_samples=10000, seed=1))
    t(f"ci linear_fit (x,y) N={n} n=1e4", lambda: boot.ci((x, y), boot.linear_fit, n_samples=10000, seed=1))
    t(f"pval mean N={n} n=1e4", lambda: boot.pval(x, np.mean, n_samples=10000, seed=1))
    t(f"ci diff_of_means independent N={n} n=1e4", lambda: boot.ci((x, y[: n // 2]), boot.diff_of_means, n_samples=10000, seed=1, multi="independent"))
