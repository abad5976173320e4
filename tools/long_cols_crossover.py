"""(N, D) moment statistics beyond one cell: per-column K1 passes vs K6i over materialised indices
(development aid for _device.GEMM_MIN_COLS)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _device
rng = np.random.default_rng(0)
for n_rows in (20000, 100000, 1000000):
    for d in (2, 4, 8, 16, 31):
        x = torch.from_numpy(rng.standard_normal((n_rows, d))).cuda()
        n = 2000
        res = []
        for min_cols in (1000, 2):
            _device.GEMM_MIN_COLS = min_cols
            for it in range(2):
                torch.cuda.synchronize(); t = time.perf_counter()
                boot.ci(x, boot.mean_axis0, n_samples=n, seed=1, method="pi")
                torch.cuda.synchronize(); dt = time.perf_counter() - t
            res.append(dt * 1e3)
        print(f"N={n_rows} D={d}: per-column K1 {res[0]:.2f} ms, K6i over indices {res[1]:.2f} ms")
