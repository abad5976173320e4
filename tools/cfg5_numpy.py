"""cfg5 from a pageable numpy array vs a CUDA tensor (development aid): the host staging of the
80 MB copy should hide under the occupancy chains."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
x = np.random.default_rng(0).standard_normal(10_000_000) + 1.0
xd = torch.from_numpy(x).cuda()
for name, data in (("cuda tensor", xd), ("numpy", x), ("cuda tensor", xd), ("numpy", x)):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = boot.ci(data, np.mean, n_samples=100_000, seed=1)
    torch.cuda.synchronize(); print(f"{name}: {1e3 * (time.perf_counter() - t):.2f} ms", out)
