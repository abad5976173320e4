"""K1m vs K1mg over table shapes (development aid; B200BOOT_K1MG=0 selects K1m,
B200BOOT_MG_MIN_COLS lowers K1mg's column threshold)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scikits_bootstrap_b200 import _device
rng = np.random.default_rng(0)
shapes = [(1000, 2, 10000), (1000, 4, 10000), (1000, 8, 10000), (1000, 16, 10000), (1000, 32, 10000), (200, 8, 10000),
          (200, 32, 2000), (1000, 8, 1000), (64, 16, 10000)]
for (n, d, m) in shapes:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = _device.resample_median_columns(x, 7, 0, m, with_jackknife=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n} D={d} n={m}: {1e3 * min(ts[1:]):.3f} ms", flush=True)
