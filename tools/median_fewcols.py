"""Resident medians of a few columns (development aid; B200BOOT_K1MG=0 takes K1m / the
column-by-column K1m1 loop instead of K1mg)."""
import os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
for n in (3000, 10000, 28000):
    for d in (2, 4, 8, 16, 32):
        x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
        f = lambda: boot.ci(x, boot.median_axis0, n_samples=10000, seed=1, method="pi")
        f(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"N={n} D={d}: {1e3 * np.median(ts):.2f} ms")
