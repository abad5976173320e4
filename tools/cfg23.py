"""cfg2 and cfg3 twice each after warm-up, for launch lists (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal(100_000) + 1).cuda()
a = torch.from_numpy(rng.gamma(2, 1, 1_000_000)).cuda()
b = torch.from_numpy(rng.gamma(3, 1, 1_000_000)).cuda()
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    boot.ci(x, np.average, method="pi", n_samples=100_000, seed=it)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    boot.ci((a, b), boot.ratio_of_means, method="bca", n_samples=100_000, multi="paired", seed=it)
    torch.cuda.synchronize(); print("cfg2", t1 - t, "cfg3", time.perf_counter() - t1)
