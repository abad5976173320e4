"""cfg4-shaped (1000 x 10000) column-batched mean, for profiling (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device
x = torch.from_numpy(np.random.default_rng(4).standard_normal((1000, 10000))).cuda()
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    o = _device.resample_stat_columns(x, _abi.STAT_MEAN, it, 0, 10000)
    torch.cuda.synchronize(); print("K1c mean 1000x10000 n=1e4", time.perf_counter() - t)
for it in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    o6 = _device.resample_stat_gemm(x, _abi.STAT_MEAN, it, 0, 10000)
    torch.cuda.synchronize(); print("K6  mean 1000x10000 n=1e4", time.perf_counter() - t)
print("max |K6 - K1c| / scale", float((o6 - o).abs().max() / o.abs().max()))
for stat, name in ((boot.mean_axis0, "mean"), (boot.std_axis0, "std")):
    for it in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = boot.ci(x, stat, n_samples=10000, seed=it)
        torch.cuda.synchronize(); print(f"ci {name}_axis0 BCa 1000x10000 n=1e4", time.perf_counter() - t)
