"""K1m vs K1mg over table shapes (development aid; B200BOOT_K1MG=0 selects K1m)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scikits_bootstrap_b200 import _device
rng = np.random.default_rng(0)
for (n, d, m) in [(1000, 64, 10000), (1000, 256, 10000), (1000, 1000, 1000), (500, 2000, 5000), (100, 10000, 10000),
                  (1000, 10000, 1000), (200, 64, 500), (1000, 128, 300), (64, 4096, 2000)]:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = _device.resample_median_columns(x, 7, 0, m, with_jackknife=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N={n} D={d} n={m}: {1e3 * min(ts[1:]):.3f} ms", flush=True)
