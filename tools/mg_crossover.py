"""K1m / K1m1-loop vs K1mg over table shapes (development aid; B200BOOT_K1MG=0 selects the
older kernels)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scikits_bootstrap_b200 import _device
rng = np.random.default_rng(0)
shapes = [(3000, 2, 10000), (3000, 8, 10000), (3000, 64, 10000), (10000, 2, 10000), (10000, 8, 10000), (10000, 32, 10000),
          (10000, 256, 10000), (28000, 4, 10000), (28000, 32, 10000), (5000, 1000, 2000), (2000, 100, 1000)]
for (n, d, m) in shapes:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    ts = []
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = _device.resample_median_columns(x, 7, 0, m, with_jackknife=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n} D={d} n={m}: {1e3 * min(ts[1:]):.3f} ms  [{float(out[0][0, 0]):.6f}]", flush=True)
