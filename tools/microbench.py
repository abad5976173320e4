"""Times individual library stages with CUDA events (development aid, not the bench)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


def main():
    n, m = 10**7, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    plan = _abi.make_plan(n, 1)
    counts = torch.empty((plan.n_tiles, m), dtype=torch.int32, device="cuda")
    cls = torch.empty((plan.n_full, m, 16), dtype=torch.int32, device="cuda")
    L = _abi.lib()
    s = _device.stream_ptr()
    t_all = timed(lambda: _abi.check(L.b200boot_tile_counts(1, n, 1, 0, m, counts.data_ptr(), cls.data_ptr(), s)))
    t_tiles = timed(lambda: _abi.check(L.b200boot_tile_counts(1, n, 1, 0, m, counts.data_ptr(), None, s)))
    print(f"K0 tile chain {t_tiles:.2f} ms, K0 + class split {t_all:.2f} ms (class split {t_all - t_tiles:.2f} ms)")
    x = torch.from_numpy(np.random.default_rng(5).standard_normal(n) + 1.0).cuda()
    t = timed(lambda: _device.resample_stat([x], _abi.STAT_MEAN, 1, 0, m))
    print(f"resample_stat total {t:.2f} ms -> {n * m / t / 1e9:.1f} Gpts/s")


if __name__ == "__main__":
    main()
