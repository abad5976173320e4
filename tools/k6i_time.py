"""K6i at cfg4's shape (1000 x 10000 means, n = 1e4), three calls (ncu target; development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scikits_bootstrap_b200 import _abi, _device
x = torch.from_numpy(np.random.default_rng(0).standard_normal((1000, 10000))).cuda()
for stat in (_abi.STAT_MEAN, _abi.STAT_SUM, _abi.STAT_STD):
    for i in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = _device.resample_stat_gemm(x, stat, 5, 0, 10000); b.record()
        torch.cuda.synchronize()
        print(f"stat {stat}: {a.elapsed_time(b):.3f} ms", float(out[0, 0]))
