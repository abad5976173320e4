"""cProfile of repeated small ci() calls (cfg1 shape): where the host time goes (development aid)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import scikits_bootstrap_b200 as boot  # noqa: E402

x = np.random.default_rng(0).standard_normal(1000)
xd = torch.from_numpy(x).cuda()
for data, name in ((x, "host array"), (xd, "device tensor")):
    for _ in range(20):
        boot.ci(data, np.mean, n_samples=10000, seed=0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(300):
        boot.ci(data, np.mean, n_samples=10000, seed=i)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t) / 300 * 1e3:.3f} ms per BCa call; "
          f"pi: ", end="")
    t = time.perf_counter()
    for i in range(300):
        boot.ci(data, np.mean, n_samples=10000, seed=i, method="pi")
    torch.cuda.synchronize()
    print(f"{(time.perf_counter() - t) / 300 * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    boot.ci(xd, np.mean, n_samples=10000, seed=i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
