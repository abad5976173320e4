"""Row-space medians of long tables (K1mr): timing (development aid)."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
for (n, d, m) in [(30000, 8, 2000), (100000, 8, 2000), (100000, 32, 1000), (1000000, 4, 500)]:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    ts = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        boot.ci(x, boot.median_axis0, n_samples=m, seed=1, method="pi")
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n} D={d} n={m}: {1e3 * min(ts[1:]):.2f} ms  ({n * d * m / min(ts[1:]) / 1e9:.1f} G point-columns/s)", flush=True)
# column counts that are not a multiple of four: the rank rows are padded (128-bit loads all the same)
for (n, d, m) in [(100000, 3, 2000), (100000, 9, 1000)]:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    ts = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        boot.ci(x, boot.median_axis0, n_samples=m, seed=1, method="pi")
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n} D={d} n={m}: {1e3 * min(ts[1:]):.2f} ms  ({n * d * m / min(ts[1:]) / 1e9:.1f} G point-columns/s)", flush=True)
