"""Device time of the small-call kernel sequence (K1 + jackknife + tail) at steady clocks: 300
library calls enqueued back to back, timed by CUDA events; and the host cost of one enqueue."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device
L = _abi.lib()
x = torch.from_numpy(np.random.default_rng(0).standard_normal(1000)).cuda()
n = 10000
a = np.array([0.025, 0.975])
ent = _device._small.get(n)
stat, res_dev, res_host = ent.stat, ent.res_dev, ent.res_host
ws, wsz = _device._ws.get(4096)
ptrs = _device._ptr_array([x])
s = _device.stream_ptr()
def call(seed):
    _abi.check(L.b200boot_ci_small(ptrs, 1, 1000, _abi.STAT_MEAN, seed, n, a.ctypes.data, 2, 1, 1, stat.data_ptr(),
                                   res_dev.data_ptr(), res_host.data_ptr(), ws, wsz, s))
for i in range(50): call(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(300): call(i)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"device: {e0.elapsed_time(e1) / 300 * 1e3:.1f} us per call; host enqueue: {(t1 - t0) / 300 * 1e6:.1f} us per call")
t0 = time.perf_counter()
for i in range(300):
    call(i); torch.cuda.current_stream().synchronize()
print(f"enqueue + synchronize: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us per call")
xh = np.random.default_rng(0).standard_normal(1000)
for data, name in ((x, "device"), (xh, "numpy")):
    for i in range(50): boot.ci(data, np.mean, n_samples=n, seed=i)
    t0 = time.perf_counter()
    for i in range(500): boot.ci(data, np.mean, n_samples=n, seed=i)
    print(f"boot.ci from {name}: {(time.perf_counter() - t0) / 500 * 1e6:.1f} us per call")
