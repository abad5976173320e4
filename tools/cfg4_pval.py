"""cfg4 ci + pval once each after warm-up, for launch lists (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
x = torch.from_numpy(np.random.default_rng(4).standard_normal((1000, 10000))).cuda()
for it in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    o = boot.ci(x, boot.median_axis0, n_samples=10000, seed=it)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    p = boot.pval(x, boot.median_axis0, n_samples=10000, seed=it)
    torch.cuda.synchronize(); print("cfg4 ci", t1 - t, "pval", time.perf_counter() - t1)
