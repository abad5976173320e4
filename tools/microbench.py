"""Times individual library stages with CUDA events (development aid, not the bench).

    python tools/microbench.py [n_samples] [--check]      # cfg5 data, N = 1e7
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device
import ctypes as C


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n, m = 10**7, int(args[0]) if args else 20000
    L = _abi.lib()
    res = {"lib": os.path.basename(_abi.LIB_PATH), "n_samples": m}
    res["k0_ms"] = timed(lambda: _device.tile_counts_packed(1, n, 1, 0, m))
    x = torch.from_numpy(np.random.default_rng(5).standard_normal(n) + 1.0).cuda()
    L.b200boot_profile_enable(1)
    k1 = []
    def run():
        out = _device.resample_stat([x], _abi.STAT_MEAN, 1, 0, m)
        ms = C.c_float()
        _abi.check(L.b200boot_profile_last_k1_ms(C.byref(ms)))
        k1.append(ms.value)
        return out
    res["total_ms"] = timed(run)
    res["k1_ms"] = min(k1)
    res["gpts_per_s"] = n * m / res["total_ms"] / 1e6
    stat = run()
    torch.cuda.synchronize()
    res["stat_mean"] = float(stat.mean())
    res["stat_std"] = float(stat.std())
    res["stat0"] = [float(v) for v in stat[:3]]
    if "--check" in sys.argv:
        # fast K0 == sequential K0, and K1 == sum over the materialised indices
        c1, k1c = _device.tile_counts_packed(7, n, 1, 0, 96)
        c0, k0c = _device.tile_counts(7, n, 1, 0, 96, with_classes=True)
        res["k0_tiles_equal"] = bool((c0 == c1).all())
        res["k0_classes_equal"] = bool((k0c == k1c.to(torch.int32)).all())
        idx = _device.fill_indices(1, n, 1, 0, 4)
        want = x[idx].mean(dim=1)
        got = _device.resample_stat([x], _abi.STAT_MEAN, 1, 0, 4)
        res["k1_vs_indices_relerr"] = float(((got.flatten() - want) / want).abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
