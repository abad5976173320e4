"""Device time of the fused small call for several shapes (which phase costs what):
n_samples tiny -> launch + row load (+ jackknife for bca); n_samples = 1e4 -> + resampling + tail."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device
L = _abi.lib()
a = np.array([0.025, 0.975])
s = _device.stream_ptr()
ws, wsz = _device._ws.get(4096)
for n_rows in (1000, 10000):
    x = torch.from_numpy(np.random.default_rng(0).standard_normal(n_rows)).cuda()
    ptrs = _device._ptr_array([x])
    for n in (32, 1000, 4704, 10000, 14112, 32768):
        ent = _device._small.get(n)
        for bca in (0, 1):
            def call(seed):
                _abi.check(L.b200boot_ci_small(ptrs, 1, n_rows, _abi.STAT_MEAN, seed, n, a.ctypes.data, 2, bca, bca,
                                               ent.stat_ptr, ent.res_dev_ptr, ent.res_host_ptr, ws, wsz, s))
            for i in range(20): call(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200): call(i)
            e1.record(); torch.cuda.synchronize()
            print(f"N={n_rows} n_samples={n} bca={bca}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call")
