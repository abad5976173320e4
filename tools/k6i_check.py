"""Checks the tcgen05 int8 contraction (K6i) against integer / FP64 references (development aid;
run under `timeout`: a wrong barrier would hang)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi, _device

L = _abi.lib()
s = _device.stream_ptr()
rng = np.random.default_rng(0)


def debug_gemm(m, n, k):
    a = torch.from_numpy(rng.integers(0, 256, (m, k), dtype=np.uint8)).cuda()
    b = torch.from_numpy(rng.integers(-64, 64, (n, k), dtype=np.int8)).cuda()
    c = torch.full((m, n), -7, dtype=torch.int32, device="cuda")
    _abi.check(L.b200boot_debug_i8_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, s))
    torch.cuda.synchronize()
    want = a.cpu().numpy().astype(np.int64) @ b.cpu().numpy().astype(np.int64).T
    got = c.cpu().numpy().astype(np.int64)
    bad = int((want != got).sum())
    print(f"debug_i8_gemm m={m} n={n} k={k}: mismatches {bad} of {want.size}", flush=True)
    if bad:
        i, j = np.argwhere(want != got)[0]
        print("  first:", i, j, want[i, j], got[i, j], " row0:", got[0, :8], want[0, :8])
    return bad == 0


ok = True
for shape in [(128, 256, 128), (128, 256, 256), (100, 512, 1024), (300, 768, 1152), (257, 256, 2048)]:
    ok = debug_gemm(*shape) and ok

# end to end: column means / stds of resamples
for (n_rows, n_cols, m) in [(1000, 100, 300), (50, 33, 129), (7169, 40, 256)]:
    x = torch.from_numpy(rng.standard_normal((n_rows, n_cols)) * 3 + 1).cuda()
    idx = _device.fill_indices(5, n_rows, 1, 0, m)
    for stat, f in ((_abi.STAT_MEAN, lambda t: t.mean(dim=1)), (_abi.STAT_STD, lambda t: t.std(dim=1, unbiased=False))):
        got = _device.resample_stat_gemm(x, stat, 5, 0, m)
        torch.cuda.synchronize()
        want = f(x[idx]).t()
        err = float(((got - want).abs() / want.abs().clamp_min(1e-300)).max())
        print(f"columns_gemm stat={stat} N={n_rows} D={n_cols} n={m}: max rel err {err:.3e}", flush=True)
        ok = ok and err < 1e-11
# a NaN row: resamples that do not draw it stay finite
x = torch.from_numpy(rng.standard_normal((200, 40))).cuda()
x[17, 3] = float("nan")
got = _device.resample_stat_gemm(x, _abi.STAT_MEAN, 9, 0, 64)
idx = _device.fill_indices(9, 200, 1, 0, 64)
want = x[idx].mean(dim=1).t()
same_nan = bool((torch.isnan(got) == torch.isnan(want)).all())
fin = ~torch.isnan(want)
err = float(((got[fin] - want[fin]).abs() / want[fin].abs()).max())
print(f"NaN column: nan pattern equal {same_nan}, finite part rel err {err:.2e}, nan count {int(torch.isnan(got).sum())}")
ok = ok and same_nan and err < 1e-12

# timing at cfg4's shape
n_rows, n_cols, m = 1000, 10000, 10000
x = torch.from_numpy(rng.standard_normal((n_rows, n_cols))).cuda()
def run():
    return _device.resample_stat_gemm(x, _abi.STAT_MEAN, 5, 0, m)
out = run(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"K6i mean of 1000 x 10000, n=1e4: {min(ts):.3f} ms (cuBLAS DGEMM path of round 1: 6.3 ms)")
cnt = torch.empty((m, n_rows), dtype=torch.float64, device="cuda")
def ref():
    _abi.check(L.b200boot_count_matrix(5, n_rows, 0, m, cnt.data_ptr(), s))
    return torch.mm(x.t(), cnt.t()) / n_rows
w = ref(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ref(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"count matrix + torch.mm DGEMM: {min(ts):.3f} ms; max abs diff vs K6i {float((out - w).abs().max()):.3e}")
print("OK" if ok else "FAILED")
