"""K1mg on a table with several units per CTA (development aid; run under `timeout`)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _device
rng = np.random.default_rng(0)
shapes = [(1000, 3000, 2000)] if len(sys.argv) > 1 else [(300, 64, 100), (1000, 600, 300), (1000, 3000, 2000), (1000, 10000, 10000)]
for (n, d, m) in shapes:
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    t0 = time.perf_counter()
    out = _device.resample_median_columns(x, 7, 0, m)
    torch.cuda.synchronize()
    print(n, d, m, "ok", f"{1e3 * (time.perf_counter() - t0):.2f} ms", float(out[0, 0]), flush=True)
