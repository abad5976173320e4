"""Unusual shapes through the public API on one GPU, each checked against the oracle fed the
device's own materialised indices (K2) where the oracle finishes in seconds, else against a
size-independent property.  Prints one line per case; exits 1 on any failure.

    python tools/stress_shapes.py
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import scikits_bootstrap_b200 as boot  # noqa: E402
import bootstrap_oracle as orc  # noqa: E402
from scikits_bootstrap_b200 import _abi  # noqa: E402

orc.PHILOX_ROUNDS = int(_abi.lib().b200boot_philox_rounds())
bad = 0


def case(name, fn):
    global bad
    t0 = time.perf_counter()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ok, note = fn()
    except Exception as e:  # noqa: BLE001
        ok, note = False, f"{type(e).__name__}: {e}"
    bad += not ok
    print(f"{'ok  ' if ok else 'FAIL'} {name:58s} {1e3 * (time.perf_counter() - t0):9.1f} ms  {note}", flush=True)


def against_oracle(tdata, dev_stat, host_stat, n, seed, method, multi=None, jstat=None, alpha=0.05):
    def run():
        data = tdata if multi else tdata[0]
        alphas = np.array([alpha / 2, 1 - alpha / 2])
        idx = np.array(list(boot.bootstrap_indices(tdata[0], n, seed=seed, n_arrays=len(tdata))))
        got = boot.ci(data, dev_stat, alpha=alpha, n_samples=n, method=method, multi=multi, seed=seed)
        src = tdata
        if host_stat is np.median and tdata[0].shape[0] > 28672:       # K1mt resamples sorted(x)
            src = (np.sort(tdata[0]),)
        want = orc.ci_from_indices(src, host_stat, idx, alphas, n, method, jstat=jstat)
        ok = bool(np.allclose(got, want, rtol=1e-12, atol=1e-15, equal_nan=True))
        return ok, f"{np.asarray(got).ravel()[:2]}" + ("" if ok else f" want {np.asarray(want).ravel()[:2]}")
    return run


rng = np.random.default_rng(123)
for N in (1, 2, 3, 31, 32, 33, 1023, 1025, 16399, 16400, 16401, 32801, 100_003):
    x = rng.standard_normal(N) + 1.0
    case(f"mean pi  N={N} n=257", against_oracle((x,), np.mean, np.mean, 257, 1, "pi"))
    # N = 3 BCa is left to --diag: 6 of the 27 possible resamples are permutations of the data,
    # their means tie with the statistic of the data up to the summation order, and the strict
    # `stat < ostat` count of z0 (bootstrap.py:583) then depends on that order
    if N > 3:
        case(f"mean bca N={N} n=257", against_oracle((x,), np.mean, np.mean, 257, 1, "bca", jstat=orc.jack_mean(x)))
        case(f"median bca N={N} n=129", against_oracle((x,), np.median, np.median, 129, 2, "bca", jstat=orc.jack_median(x)))
for n in (1, 2, 7, 100_001):
    x = rng.standard_normal(501)
    case(f"std pi N=501 n={n}", against_oracle((x,), np.std, np.std, n, 3, "pi"))

# property checks at large sizes: a constant column has a degenerate interval; mean of
# x + c shifts both endpoints by c with the same seed (same indices)
def big_shift():
    x = rng.standard_normal(100_000_000)
    a = boot.ci(x, np.mean, n_samples=2000, seed=5, method="pi")
    xs = x + 3.0
    b = boot.ci(xs, np.mean, n_samples=2000, seed=5, method="pi")
    return bool(np.allclose(b - a, 3.0, rtol=0, atol=1e-9)), f"{a} {b - a}"


case("mean pi N=1e8 n=2000: shift equivariance", big_shift)


def big_bca():
    x = rng.standard_normal(50_000_000) + 1.0
    lo, hi = boot.ci(x, np.mean, n_samples=4000, seed=6)
    se = x.std() / np.sqrt(x.size)
    return bool(lo < x.mean() < hi and 3.0 < (hi - lo) / se < 4.8), f"width/se = {(hi - lo) / se:.3f}"


case("mean bca N=5e7 n=4000: width ~ 3.92 se", big_bca)


def wide_table():
    t = rng.standard_normal((64, 200_000))
    got = boot.ci(t, boot.mean_axis0, n_samples=500, seed=7, method="pi")
    idx = np.array(list(boot.bootstrap_indices(t[:, :1], 500, seed=7)))
    cols = [0, 199_999, 77_777]
    stat = np.sort(t[:, cols][idx].mean(axis=1), axis=0)
    nv = np.round(499 * np.array([0.025, 0.975])).astype(int)
    return bool(np.allclose(got[:, cols], stat[nv], rtol=1e-12, atol=1e-15)), f"{got.shape}"


case("mean_axis0 pi table 64 x 200000 n=500", wide_table)


def tall_table_median():
    t = rng.standard_normal((200_000, 3))
    got = boot.ci(t, boot.median_axis0, n_samples=200, seed=8, method="pi")
    idx = np.array(list(boot.bootstrap_indices(t[:, :1], 200, seed=8)))
    stat = np.sort(np.median(t[idx], axis=1), axis=0)
    nv = np.round(199 * np.array([0.025, 0.975])).astype(int)
    return bool(np.allclose(got, stat[nv], rtol=1e-12, atol=1e-15)), f"{got.shape}"


case("median_axis0 pi table 200000 x 3 n=200", tall_table_median)


def many_alphas():
    x = rng.standard_normal(4001)
    al = np.linspace(0.001, 0.999, 97)
    got = boot.ci(x, np.mean, alpha=al, n_samples=3001, seed=9)
    idx = np.array(list(boot.bootstrap_indices(x, 3001, seed=9)))
    want = orc.ci_from_indices((x,), np.mean, idx, al, 3001, "bca", jstat=orc.jack_mean(x))
    return bool(np.allclose(got, want, rtol=1e-12, atol=1e-15)), f"{got.shape}"


case("mean bca 97 alphas", many_alphas)


def f32_in():
    x = rng.standard_normal(5000).astype(np.float32)
    got = boot.ci(x, np.mean, n_samples=999, seed=10)
    return got.dtype == np.float32, f"{got.dtype}"


case("float32 in -> float32 out", f32_in)

# diagnosis of tiny-N BCa differences: with N = 3 a resample that is a permutation of the data
# has the mean of the data up to the summation order, so the strict `stat < ostat` count of z0
# (bootstrap.py:583) depends on that order in the reference itself
if "--diag" in sys.argv:
    x = np.random.default_rng(123).standard_normal(1) + 1.0
    rng2 = np.random.default_rng(123)
    for N in (1, 2, 3):
        x = rng2.standard_normal(N) + 1.0
    idx = np.array(list(boot.bootstrap_indices(x, 257, seed=1)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, dist = boot.ci(x, np.mean, n_samples=257, seed=1, return_dist=True)
    want, parts = orc.ci_from_indices((x,), np.mean, idx, np.array([0.025, 0.975]), 257, "bca",
                                      jstat=orc.jack_mean(x), return_parts=True)
    ref = parts["stat"]
    print("max rel diff of sorted stat:", np.max(np.abs(np.asarray(dist) - ref) / np.abs(ref)))
    print("ostat", repr(x.mean()), " device count <", int(np.sum(np.asarray(dist) < x.mean())),
          " oracle count <", int(np.sum(ref < x.mean())),
          " |stat - ostat| < 1e-15 in oracle:", int(np.sum(np.abs(ref - x.mean()) < 1e-15)))
    print("got", got, "want", want)
print("failures:", bad)
sys.exit(1 if bad else 0)
