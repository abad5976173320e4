"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Parity has two legs:
  (i)  index-supplied mode (K1i): the GPU consumes the REFERENCE's own PCG64 index
       arrays; gathers are bit-exact, statistics agree to 1e-12 relative and the
       interval endpoints reproduce the reference's golden values;
  (ii) fused mode (K1): the GPU's own Philox draws, materialised by K2, are fed to
       the oracle; statistics agree to 1e-12 and endpoints are identical.
Tolerance: bit-exact for indices/gathers/counts/order statistics; 1e-12 relative
for FP64 statistics (north_star)."""
import functools

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import bootstrap_oracle as orc
import scikits_bootstrap_b200 as boot
from cases import CASES, DATA20, SEED, STATS, gen_data
from scikits_bootstrap_b200 import _abi, _device

pytestmark = pytest.mark.gpu

RTOL = 1e-12

X = [1, 2, 3, 4, 5, 6]
Y = [2, 1, 2, 5, 1, 2]
Z = [2, 1, 1, -1, -1, -4, -8]

STATS = dict(STATS, sum=np.sum, linear_fit=lambda a, b: np.polyfit(a, b, 1))

DEVICE_STATS = {
    "sum": np.sum,
    "mean": np.mean, "average": np.average, "var": np.var, "std": np.std, "median": np.median,
    "ratio_of_means": boot.ratio_of_means, "weighted_average": boot.weighted_average,
    "pearson_r": boot.pearson_r, "linear_fit": boot.linear_fit, "median_axis0": boot.median_axis0, "mean_axis0": boot.mean_axis0,
    "diff_of_means": boot.diff_of_means,
    "percentile90": boot.percentile(90), "quantile25_axis0": boot.quantile_axis0(0.25),
}
JACK = {
    "sum": lambda x: np.sum(x) - np.asarray(x, dtype=float),
    "mean": orc.jack_mean, "average": orc.jack_mean, "var": orc.jack_var,
    "std": lambda x: orc.jack_var(x, True), "ratio_of_means": orc.jack_ratio_of_means,
    "weighted_average": orc.jack_weighted_average, "pearson_r": orc.jack_pearson, "linear_fit": orc.jack_linear_fit,
    "median": orc.jack_median,
}


def torch():
    import torch as t
    return t


def dev(x, dtype=None):
    t = torch().as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


def pcg_indices(n, n_samples, seed=SEED):
    return np.array(list(orc.bootstrap_indices_pcg64(n, n_samples, seed)))


# ------------------------------------------------------------------ gathers ------

def test_gather_is_bit_exact():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(100003)
    x[:4] = [0.0, -0.0, np.inf, 5e-324]
    idx = rng.integers(0, len(x), 500000)
    got = _device.gather(dev(x), dev(idx)).cpu().numpy()
    assert_array_equal(got.view(np.uint64), x[idx].view(np.uint64))


# ------------------------------------------------------------------ leg (i) ------

def test_reference_goldens_in_index_supplied_mode():
    """tests/bootstrap_test.py:163-169, 205-209, 357-363 with the reference's indices."""
    idx = pcg_indices(20, 10000)
    r = boot.ci_indexed(DATA20, np.average, idx, alpha=(0.1, 0.2, 0.8, 0.9), method="pi")
    assert_allclose(r, [0.401879, 0.517506, 0.945416, 1.052798], rtol=1e-6)
    r = boot.ci_indexed(DATA20, None, idx, alpha=(0.1, 0.2, 0.8, 0.9))
    assert_allclose(r, [0.386674, 0.506714, 0.935628, 1.039683], rtol=1e-6)
    r = boot.ci_indexed(DATA20, None, pcg_indices(20, 500), alpha=(0.1, 0.2, 0.8, 0.9))
    assert_allclose(r, [0.37248, 0.507976, 0.92783, 1.039755], rtol=1e-6)


def test_reference_pval_goldens_in_index_supplied_mode():
    """tests/bootstrap_test.py:396-411."""
    r = boot.pval_indexed(DATA20, compfunction=lambda s: 0.8 <= s <= 1.2, indices=pcg_indices(20, 500))
    assert_allclose(r, 0.368)
    r = boot.pval_indexed(DATA20, compfunction=boot.between(0.8, 1.2), indices=pcg_indices(20, 500))
    assert_allclose(r, 0.368)
    r = boot.pval_indexed(Z, indices=pcg_indices(7, 500))
    assert_allclose(r, 0.096)


MOMENT_CASES = [c for c in CASES if c[3] != "independent"]      # incl. median and (N, D) cases


@pytest.mark.parametrize("case", MOMENT_CASES, ids=[c[0] for c in MOMENT_CASES])
def test_index_supplied_matches_reference_fixture(case, golden):
    name, spec, stat, multi, n, alphas = case
    tdata = gen_data(spec)
    data = tdata if multi else tdata[0]
    f = DEVICE_STATS[stat]
    idx = pcg_indices(len(tdata[0]), n)
    # np.polyfit solves its least squares by SVD, the device from centred moment sums
    rtol = 1e-9 if stat == "linear_fit" else RTOL
    for method in ("pi", "bca"):
        out, dist = boot.ci_indexed(data, f, idx, alpha=alphas, method=method, multi=multi,
                                    return_dist=True)
        assert_allclose(dist, golden[f"{name}/dist"], rtol=rtol, atol=1e-15)
        assert_allclose(out, golden[f"{name}/{method}/lowhigh"], rtol=rtol, atol=1e-15)
        eb = boot.ci_indexed(data, f, idx, alpha=alphas, method=method, multi=multi, output="errorbar")
        assert eb.shape == golden[f"{name}/{method}/errorbar"].shape
        assert_allclose(eb, golden[f"{name}/{method}/errorbar"], rtol=1e-9, atol=1e-13)
    assert_allclose(boot.pval_indexed(data, f, indices=idx, multi=bool(multi)), golden[f"{name}/pval"])


# ------------------------------------------------------------------ K2 / K0 ------

@pytest.mark.parametrize("n", [1, 5, 20, 1000, 28672])
def test_device_indices_match_oracle_single_cell(n):
    got = np.array(list(boot.bootstrap_indices(np.zeros(n), 37, seed=SEED)))
    assert got.shape == (37, n) and got.dtype == np.int64
    for r in (0, 1, 36):
        assert_array_equal(got[r], orc.device_indices(SEED, r, n, 14))


@pytest.mark.parametrize("n,n_arrays", [(40000, 1), (100000, 1), (20000, 2), (16384 * 2, 1)])
def test_device_indices_match_oracle_tiled(n, n_arrays):
    plan = _abi.make_plan(n, n_arrays)
    counts, cls = _device.tile_counts(9, n, n_arrays, 0, 16, with_classes=True)
    counts, cls = counts.cpu().numpy(), cls.cpu().numpy()             # [n_tiles][16], [n_full][16][16]
    assert (counts.sum(axis=0) == n).all()
    assert (cls.sum(axis=2) == counts[:plan.n_full]).all() and (cls >= 0).all()
    got = _device.fill_indices(9, n, n_arrays, 0, 16).cpu().numpy()
    for r in (0, 7, 15):
        want = orc.device_indices(9, r, n, plan.log2_tile, counts=list(counts[:, r]), n_arrays=n_arrays,
                                  class_counts=cls[:, r, :])
        assert_array_equal(got[r], want)
    # host probe (same sources, host libm) normally gives the same occupancies
    host = np.empty(plan.n_tiles, dtype=np.int32)
    same = 0
    for r in range(16):
        _abi.check(_abi.lib().b200boot_host_tile_counts(9, n, n_arrays, r, host.ctypes.data))
        same += int((host == counts[:, r]).all())
    assert same >= 12


def test_indices_do_not_depend_on_the_shard():
    n = 50000
    full = _device.fill_indices(5, n, 1, 0, 24).cpu().numpy()
    part = np.concatenate([_device.fill_indices(5, n, 1, lo, hi).cpu().numpy()
                           for lo, hi in [(0, 5), (5, 6), (6, 24)]])
    assert_array_equal(full, part)


def test_tile_counts_multinomial_on_device():
    n, m = 10**7, 512
    cnt = _device.tile_counts(1, n, 1, 0, m).cpu().numpy().astype(float)     # [611][m]
    assert (cnt.sum(axis=0) == n).all()
    sizes = np.array([16384] * 610 + [5760])
    z = (cnt.mean(axis=1) - sizes) / np.sqrt(sizes * (1 - sizes / n) / m)
    assert np.abs(z).max() < 5.5
    assert abs(z.std() - 1) < 0.15
    # 16-way bank-class split of each full tile
    _, cls = _device.tile_counts(1, n, 1, 0, 64, with_classes=True)
    cls = cls.cpu().numpy().astype(float)                                    # [610][64][16]
    tile_n = cls.sum(axis=2, keepdims=True)
    zc = (cls - tile_n / 16) / np.sqrt(tile_n * (1 / 16) * (15 / 16))
    assert abs(zc.mean()) < 0.01 and abs(zc.std() - 1) < 0.01 and np.abs(zc).max() < 6.5


# ------------------------------------------------------------------ leg (ii) -----

def fused_vs_oracle(tdata, stat_name, n_samples, seed, multi=None, alphas=(0.025, 0.975), RTOL=RTOL):
    f = DEVICE_STATS[stat_name]
    data = tdata if multi else tdata[0]
    idx = np.array(list(boot.bootstrap_indices(tdata[0], n_samples, seed=seed, n_arrays=len(tdata))))
    want = orc.stat_over_indices(tdata, STATS[stat_name], idx)
    want.sort(axis=0)
    for method in ("pi", "bca"):
        out, dist = boot.ci(data, f, alpha=alphas, n_samples=n_samples, method=method, multi=multi,
                            seed=seed, return_dist=True)
        assert_allclose(dist, want, rtol=RTOL, atol=1e-15)
        jstat = JACK[stat_name](*tdata)
        ref = orc.ci_from_indices(tdata, STATS[stat_name], idx, alphas, n_samples, method,
                                  jstat=jstat)
        assert_allclose(out, ref, rtol=RTOL, atol=1e-15)


@pytest.mark.parametrize("stat_name", ["mean", "var", "std", "sum"])
@pytest.mark.parametrize("n", [20, 1000, 28672, 40000])
def test_fused_single_array(stat_name, n):
    x = np.random.default_rng(n).standard_normal(n) + 0.5
    fused_vs_oracle((x,), stat_name, 400, seed=n + 1)


@pytest.mark.parametrize("stat_name", ["ratio_of_means", "weighted_average", "pearson_r"])
@pytest.mark.parametrize("n", [50, 3000, 14336, 30000])
def test_fused_paired(stat_name, n):
    rng = np.random.default_rng(n)
    a = rng.gamma(2, 1, n)
    b = 0.5 * a + rng.gamma(3, 1, n)
    fused_vs_oracle((a, b), stat_name, 400, seed=n + 2, multi="paired")


@pytest.mark.parametrize("n", [10, 50, 3000, 14336, 30000])
def test_fused_linear_fit(n):
    """np.polyfit(a, b, 1): a statistic with two outputs per resample (slope, intercept) from one
    pass over the five centred moment sums; numpy solves the same least squares by SVD, hence
    the looser tolerance."""
    rng = np.random.default_rng(n)
    a = rng.gamma(2, 1, n) + 10.0
    b = 0.5 * a + rng.gamma(3, 1, n)
    fused_vs_oracle((a, b), "linear_fit", 300, seed=n + 3, multi="paired", RTOL=1e-9)
    m = 300
    both, dist = boot.ci((a, b), boot.linear_fit, n_samples=m, seed=n, return_dist=True)
    assert both.shape == (2, 2) and dist.shape == (m, 2)
    assert_array_equal(both[:, 0], boot.ci((a, b), boot.fit_slope, n_samples=m, seed=n))
    assert_array_equal(both[:, 1], boot.ci((a, b), boot.fit_intercept, n_samples=m, seed=n))
    p = boot.pval((a, b), boot.linear_fit, boot.greater_than(0.5), n_samples=m, seed=n)
    assert_array_equal(p, (dist > 0.5).mean(axis=0))
    # chunked by a small workspace budget: same values
    import os
    os.environ["B200BOOT_WS_BYTES"] = str(1 << 20)
    try:
        again = boot.ci((a, b), boot.linear_fit, n_samples=m, seed=n, return_dist=True)[1]
    finally:
        del os.environ["B200BOOT_WS_BYTES"]
    assert_array_equal(again, dist)


@pytest.mark.parametrize("n", [1, 2, 3, 20, 301, 1000, 1024, 1025, 5000, 28672])
def test_fused_median_1d(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) if n % 3 else np.round(rng.standard_normal(n), 1)     # ties too
    m = 300
    idx = np.array(list(boot.bootstrap_indices(x, m, seed=n)))
    want = np.sort(orc.stat_over_indices((x,), np.median, idx))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, dist = boot.ci(x, np.median, n_samples=m, seed=n, method="pi", return_dist=True)
    assert_array_equal(dist, want)                       # order statistics: bit-exact
    if n >= 20:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = boot.ci(x, np.median, n_samples=m, seed=n, method="bca")
            ref = orc.ci_from_indices((x,), np.median, idx, (0.025, 0.975), m, "bca", jstat=orc.jack_median(x))
        assert_array_equal(got, ref)


@pytest.mark.parametrize("n", [28673, 40000, 100001, 16384 * 4])
def test_long_median_runs_in_rank_space(n):
    """Beyond the resident limit the median is taken over sorted(x)[bootstrap_indices]:
    only the tile holding the middle rank is drawn (K1mt)."""
    import warnings
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) if n % 2 else np.round(rng.standard_normal(n), 3)      # ties too
    m = 200
    idx = np.array(list(boot.bootstrap_indices(x, m, seed=n)))
    xs = np.sort(x)
    want = np.sort(np.median(xs[idx], axis=1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, dist = boot.ci(x, np.median, n_samples=m, seed=n, method="pi", return_dist=True)
        assert_array_equal(dist, want)
        got = boot.ci(x, np.median, n_samples=m, seed=n, method="bca")
        ref = orc.ci_from_indices((xs,), np.median, idx, (0.025, 0.975), m, "bca", jstat=orc.jack_median(xs))
    assert_array_equal(got, ref)


@pytest.mark.parametrize("q", [0.0, 0.1, 0.25, 1.0 / 3.0, 0.5, 0.9, 0.975, 1.0])
@pytest.mark.parametrize("n", [1, 2, 3, 20, 301, 1000, 5000, 28672])
def test_fused_quantile_1d(n, q):
    """np.quantile(x, q) (method 'linear') in rank space: the two order statistics are found
    exactly and combined with numpy's own interpolation formula -> bit-exact."""
    import warnings
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) if n % 3 else np.round(rng.standard_normal(n), 1)     # ties too
    m = 200
    idx = np.array(list(boot.bootstrap_indices(x, m, seed=n)))
    want = np.sort(np.quantile(x[idx], q, axis=1))
    stat = boot.quantile(q)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, dist = boot.ci(x, stat, n_samples=m, seed=n, method="pi", return_dist=True)
        assert_array_equal(dist, want)
        if n >= 20:
            got = boot.ci(x, functools.partial(np.quantile, q=q), n_samples=m, seed=n, method="bca")
            ref = orc.ci_from_indices((x,), lambda u: np.quantile(u, q), idx, (0.025, 0.975), m, "bca",
                                      jstat=orc.jack_quantile(x, q))
            assert_array_equal(got, ref)
            assert_array_equal(boot.ci_indexed(x, stat, idx, method="bca"), ref)


def test_fused_quantile_columns_and_percentile():
    import warnings
    rng = np.random.default_rng(44)
    data = rng.standard_normal((700, 37))
    data[:, 5] = np.round(data[:, 5], 1)
    m = 300
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=3)))
    for p in (5, 50, 90):
        f = functools.partial(np.percentile, q=p, axis=0)
        want = np.sort(np.stack([f(data[i]) for i in idx]), axis=0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out, dist = boot.ci(data, f, n_samples=m, seed=3, return_dist=True)
            assert_array_equal(dist, want)
            jst = np.stack([orc.jack_quantile(data[:, d], np.true_divide(p, 100)) for d in range(37)], axis=1)
            ref = orc.ci_from_indices((data,), f, idx, (0.025, 0.975), m, "bca", jstat=jst)
            # identical leave-one-out values (the tied column): the acceleration is 0/0 here and
            # rounding noise of np.mean in numpy, so only the other columns are comparable
            ok = np.ptp(jst, axis=0) > 0
            assert ok.sum() >= 36
            assert_array_equal(out[:, ok], ref[:, ok])
            assert_array_equal(boot.ci(data, boot.percentile_axis0(p), n_samples=m, seed=3), out)
            assert_array_equal(boot.ci(data[:, 7], boot.percentile(p), n_samples=m, seed=3), out[:, 7])
    with pytest.raises(ValueError):
        boot.quantile(1.5)
    with pytest.raises(NotImplementedError):
        boot.ci(data[:, 0], functools.partial(np.quantile, q=0.5, method="nearest"), n_samples=10)


@pytest.mark.parametrize("n", [28673, 50001])
def test_long_quantile_runs_in_rank_space(n):
    import warnings
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    xs = np.sort(x)
    m = 200
    idx = np.array(list(boot.bootstrap_indices(x, m, seed=n)))
    for q in (0.02, 0.5, 0.9, 1.0):
        want = np.sort(np.quantile(xs[idx], q, axis=1))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out, dist = boot.ci(x, boot.quantile(q), n_samples=m, seed=n, method="pi", return_dist=True)
            assert_array_equal(dist, want)
            got = boot.ci(x, boot.quantile(q), n_samples=m, seed=n, method="bca")
            ref = orc.ci_from_indices((xs,), lambda u: np.quantile(u, q), idx, (0.025, 0.975), m, "bca",
                                      jstat=orc.jack_quantile(xs, q))
        assert_array_equal(got, ref)


@pytest.mark.parametrize("n,d,q", [(30001, 3, None), (28673, 5, None), (40000, 2, 0.9), (65536, 9, 0.25),
                                   (29001, 4, None), (33000, 8, 0.4)])      # whole groups of four columns: 128-bit rank loads only
def test_long_median_over_columns_is_row_aligned(n, d, q):
    """(N, D) data with more rows than K1m's resident limit stay in row space (K1mr): every
    resample gathers whole rows, x[indices] (bootstrap.py:431), so the returned distribution
    holds the oracle's joint rows on the materialised indices, for medians and quantiles."""
    import warnings
    rng = np.random.default_rng(n + d)
    data = rng.standard_normal((n, d))
    data[:, 1] = np.round(data[:, 1], 2 if q is None else 5)   # ties in one column (heavy for the median)
    f = boot.median_axis0 if q is None else boot.quantile_axis0(q)
    host = (lambda x: np.median(x, axis=0)) if q is None else (lambda x: np.quantile(x, q, axis=0))
    jack1 = orc.jack_median if q is None else (lambda c: orc.jack_quantile(c, q))
    m = 120
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, dist = boot.ci(data, f, n_samples=m, seed=5, return_dist=True)
        idx = np.array(list(boot.bootstrap_indices(data, m, seed=5)))
        want = np.stack([host(data[i]) for i in idx])
        assert_array_equal(dist, np.sort(want, axis=0))
        jst = np.stack([jack1(data[:, c]) for c in range(d)], axis=1)
        ref = orc.ci_from_indices((data,), host, idx, (0.025, 0.975), m, "bca", jstat=jst)
        # a column whose leave-one-out values are all equal (heavy ties) has no acceleration: the
        # device says NaN exactly, the reference is at the mercy of np.mean's rounding
        with np.errstate(invalid="ignore", divide="ignore"):
            defined = [c for c in range(d) if np.isfinite(orc.acceleration(jst[:, c]))]
        assert 0 in defined
        assert_array_equal(out[:, defined], ref[:, defined])
        eb = boot.ci(dev(data), f, n_samples=m, seed=5, output="errorbar")
        assert_array_equal(eb, np.abs(host(data) - out).T)
        p = boot.pval(data, f, boot.greater_than(0.005), n_samples=m, seed=5)
        assert_array_equal(p, (want > 0.005).mean(axis=0))
        # the joint rows are rows of ONE resample: the unsorted statistic vectors line up
        # (column c's vector equals the row-space statistic of that column alone)
        run_idx = boot.ci_indexed(data[:, 0], np.median if q is None else boot.quantile(q), indices=idx,
                                  method="pi", return_dist=True)[1]
        assert_array_equal(run_idx, np.sort(want[:, 0]))


@pytest.mark.parametrize("n,d", [(400, 1), (3000, 4), (30001, 1), (29000, 3)])
def test_rank_statistics_propagate_nan_like_numpy(n, d):
    """np.median / np.quantile of a sample holding a NaN are NaN, so the reference's statistic is NaN
    for every resample that draws a NaN row; all rank paths (resident K1m, rank-space K1mt,
    row-space K1mr, index-supplied) must say the same."""
    import warnings
    rng = np.random.default_rng(n)
    data = rng.standard_normal((n, d))
    data[3, 0] = np.nan                               # one NaN row in column 0 ...
    if d > 2:
        data[[5, 7, 11], 2] = np.nan                  # ... three in column 2, none in the others
    x = data[:, 0] if d == 1 else data
    m = 60
    for f, host in ((np.median if d == 1 else boot.median_axis0, lambda u: np.median(u, axis=0)),
                    (boot.quantile(0.3) if d == 1 else boot.quantile_axis0(0.3), lambda u: np.quantile(u, 0.3, axis=0))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out, dist = boot.ci(x, f, n_samples=m, seed=2, method="pi", return_dist=True)
            idx = np.array(list(boot.bootstrap_indices(x, m, seed=2)))
            src = np.sort(x) if (d == 1 and n > 28672) else x          # K1mt resamples sorted(x)
            want = np.stack([host(src[i]) for i in idx]).reshape(m, -1)
            assert np.isnan(want[:, 0]).any() and (d == 1 or not np.isnan(want[:, 1]).any())
            assert_array_equal(dist.reshape(m, -1), np.sort(want, axis=0))
            bca = boot.ci(x, f, n_samples=m, seed=2, method="bca")
            # column 0 holds a NaN: its leave-one-out values are NaN (bar one), the acceleration is NaN
            jst = np.full((n,) if d == 1 else (n, d), np.nan)
            ref = orc.ci_from_indices((src,), host, idx, (0.025, 0.975), m, "bca", jstat=jst)
            assert_array_equal(np.asarray(bca).reshape(2, -1)[:, 0], np.asarray(ref).reshape(2, -1)[:, 0])
            pcg = np.array(list(orc.bootstrap_indices_pcg64(n, 20, 4)))
            got = boot.ci_indexed(x, f, indices=pcg, method="pi", return_dist=True)[1]
            assert_array_equal(got.reshape(20, -1), np.sort(np.stack([host(x[i]) for i in pcg]).reshape(20, -1), axis=0))


@pytest.mark.parametrize("n,d", [(32, 64), (100, 130), (513, 64), (1000, 72), (1023, 200), (1024, 65)])
def test_many_short_columns_through_the_tensor_cores(n, d):
    """K1mg (N <= 1024 rows, D >= 64 columns): 32 prefix counts per (resample, column) as a count
    matrix x 0/1 indicator matrix product on tcgen05, finished by a short walk.  The returned
    distribution must hold the oracle's rows on the materialised indices — medians (one and two
    middle ranks), an interpolated quantile, heavy ties, resample counts off the tile size."""
    import warnings
    rng = np.random.default_rng(n * 1000 + d)
    data = rng.standard_normal((n, d))
    data[:, 1] = np.round(data[:, 1], 1)                 # heavy ties
    data[:, 2] = 1.5                                     # a constant column
    m = 130
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=8)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f, host in ((boot.median_axis0, lambda u: np.median(u, axis=0)),
                        (boot.quantile_axis0(0.33), lambda u: np.quantile(u, 0.33, axis=0))):
            out, dist = boot.ci(data, f, n_samples=m, seed=8, method="pi", return_dist=True)
            want = np.stack([host(data[i]) for i in idx])
            assert_array_equal(dist, np.sort(want, axis=0))
            p = boot.pval(data, f, boot.greater_than(0.1), n_samples=m, seed=8)
            assert_array_equal(p, (want > 0.1).mean(axis=0))
        for c in (0, 1, d - 1):                         # column c == the 1-D run with the same seed
            assert_array_equal(boot.ci(data[:, c], np.median, n_samples=m, seed=8, method="pi"),
                               boot.ci(data, boot.median_axis0, n_samples=m, seed=8, method="pi")[:, c])


def test_long_median_index_supplied():
    """Index-supplied (parity) mode beyond the resident limit: the reference's own PCG64 index
    sets, 1-D and (N, D) data."""
    n, m = 30011, 60
    rng = np.random.default_rng(8)
    x = rng.gamma(2.0, 1.0, n)
    idx = np.array(list(orc.bootstrap_indices_pcg64(n, m, 21)))
    want = np.sort(np.array([np.median(x[i]) for i in idx]))
    out, dist = boot.ci_indexed(x, np.median, indices=idx, method="pi", return_dist=True)
    assert_array_equal(dist, want)
    data = np.stack([x, -x, np.round(x, 1)], axis=1)
    out, dist = boot.ci_indexed(data, boot.median_axis0, indices=idx, method="pi", return_dist=True)
    assert_array_equal(dist, np.sort(np.stack([np.median(data[i], axis=0) for i in idx]), axis=0))


def test_fused_median_columns_share_the_index_set():
    """cfg4 shape in miniature: (N, D) data, median over axis 0, BCa, plus pval."""
    rng = np.random.default_rng(4)
    data = rng.standard_normal((1000, 70))
    m = 400
    out, dist = boot.ci(data, boot.median_axis0, n_samples=m, seed=9, return_dist=True)
    assert out.shape == (2, 70) and dist.shape == (m, 70)
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=9)))
    want = orc.stat_over_indices((data,), STATS["median_axis0"], idx)
    want.sort(axis=0)
    assert_array_equal(dist, want)
    jst = np.stack([orc.jack_median(data[:, d]) for d in range(70)], axis=1)
    ref = orc.ci_from_indices((data,), STATS["median_axis0"], idx, (0.025, 0.975), m, "bca", jstat=jst)
    assert_array_equal(out, ref)
    for d in (0, 33, 69):                                 # column d == the 1-D run with the same seed
        assert_array_equal(boot.ci(data[:, d], np.median, n_samples=m, seed=9), out[:, d])
    p = boot.pval(data, functools.partial(np.median, axis=0), n_samples=m, seed=9)
    assert_array_equal(p, (want > 0).mean(axis=0))


@pytest.mark.parametrize("n", [2, 3, 10, 301, 1000])
def test_jackknife_median_closed_form(n):
    rng = np.random.default_rng(n)
    data = np.round(rng.standard_normal((n, 3)), 1 if n > 100 else 6)
    out = _device.jackknife_median_columns(dev(data)).cpu().numpy()
    for d in range(3):
        j = orc.jack_median(data[:, d])
        assert_allclose(out[d, 0], np.median(data[:, d]), rtol=0, atol=0)
        assert_allclose(out[d, 1], j.mean(), rtol=1e-12, atol=1e-15)
        with np.errstate(invalid="ignore", divide="ignore"):
            a = orc.acceleration(j)
        # all-equal leave-one-out medians (ties): 0/0 here, rounding noise in numpy's mean
        if np.isfinite(a) and np.sum((j.mean() - j) ** 2) > 1e-20:
            assert_allclose(out[d, 4], a, rtol=1e-6, atol=1e-12)
        if n <= 301:
            assert_array_equal(j, orc.jackknife_stats_loop((data[:, d],), np.median))


def test_fused_odd_sizes_and_unaligned_views():
    """Remainder cells of odd length and 8-byte-aligned (not 16) device views take
    the non-TMA load path; results must not change."""
    base = np.random.default_rng(3).standard_normal(70001)
    for n in (16385 + 28672, 32769, 70000):
        x = base[1:1 + n]
        want = boot.ci(x.copy(), np.mean, n_samples=300, seed=4, return_dist=True)[1]
        view = dev(base)[1:1 + n]                      # data_ptr is 8 mod 16
        assert view.data_ptr() % 16 == 8
        got = boot.ci(view, np.mean, n_samples=300, seed=4, return_dist=True)[1]
        assert_array_equal(got, want)


def test_fuzz_shapes_against_oracle():
    """Random sizes around every internal boundary (cell capacity, tile and class sizes, block
    lengths, odd/even, one- and two-array plans): the fused kernels must reproduce the oracle on
    the device's own index sets."""
    import warnings
    rng = np.random.default_rng(2024)
    special = [1, 2, 3, 4, 5, 11, 12, 13, 31, 32, 33, 47, 48, 49, 191, 192, 193, 1023, 1024, 1025, 14335, 14336,
               14337, 16383, 16384, 16385, 28671, 28672, 28673, 32767, 32768, 32769, 45055, 49152, 57343]
    sizes = special + [int(v) for v in rng.integers(1, 60000, 25)]
    names = ["mean", "var", "std", "ratio_of_means", "weighted_average", "pearson_r", "median"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, n in enumerate(sizes):
            name = names[k % len(names)]
            paired = name in ("ratio_of_means", "weighted_average", "pearson_r")
            if name == "pearson_r" and n < 3:
                name, paired = "mean", False
            a = rng.gamma(2.0, 1.0, n) + 0.1
            b = 0.3 * a + rng.gamma(3.0, 1.0, n) + 0.1
            tdata = (a, b) if paired else (a,)
            m = int(rng.integers(1, 40))
            seed = int(rng.integers(0, 2**62))
            idx = np.array(list(boot.bootstrap_indices(a, m, seed=seed, n_arrays=len(tdata)))).reshape(m, n)
            assert idx.min() >= 0 and idx.max() < n
            src = tuple(np.sort(t) for t in tdata) if (name == "median" and n > 28672) else tdata
            want = np.sort(orc.stat_over_indices(src, STATS[name], idx))
            _, dist = boot.ci(tdata if paired else a, DEVICE_STATS[name], n_samples=m, seed=seed, method="pi",
                              return_dist=True, multi="paired" if paired else None)
            if name == "median":
                assert_array_equal(dist, want, err_msg=f"{name} n={n} m={m}")
            else:
                assert_allclose(dist, want, rtol=RTOL, atol=1e-13, err_msg=f"{name} n={n} m={m}")


def test_fuzz_newer_statistics_against_numpy():
    """The same fuzz for the statistics added after the first registry: quantiles (bit-exact),
    the two-output linear fit, and differences of independent samples."""
    import warnings
    rng = np.random.default_rng(77)
    special = [2, 3, 4, 5, 12, 13, 33, 193, 1025, 14337, 16385, 28671, 28672, 28673, 32769, 49152]
    sizes = special + [int(v) for v in rng.integers(2, 60000, 14)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, n in enumerate(sizes):
            a = np.round(rng.gamma(2.0, 1.0, n) + 0.1, int(rng.integers(1, 12)))
            b = 0.3 * a + rng.gamma(3.0, 1.0, n) + 0.1
            m = int(rng.integers(1, 30))
            seed = int(rng.integers(0, 2**62))
            kind = k % 3
            if kind == 0:                      # quantile
                q = float(rng.choice([0.0, 1.0, 0.5, rng.random(), rng.random()]))
                idx = np.array(list(boot.bootstrap_indices(a, m, seed=seed))).reshape(m, n)
                src = np.sort(a) if n > 28672 else a
                want = np.sort(np.quantile(src[idx], q, axis=1))
                _, dist = boot.ci(a, boot.quantile(q), n_samples=m, seed=seed, method="pi", return_dist=True)
                assert_array_equal(dist, want, err_msg=f"quantile q={q} n={n} m={m}")
            elif kind == 1 and n >= 3:         # linear fit
                idx = np.array(list(boot.bootstrap_indices(a, m, seed=seed, n_arrays=2))).reshape(m, n)
                keep = np.array([np.ptp(a[i]) > 0 for i in idx])      # a constant abscissa has no line
                want = np.array([np.polyfit(a[i], b[i], 1) for i in idx[keep]])
                _, dist = boot.ci((a, b), boot.linear_fit, n_samples=m, seed=seed, method="pi", return_dist=True)
                got = _device.resample_stat([dev(a), dev(b)], _abi.STAT_LINEAR_FIT, boot.bootstrap._seed_key(seed),
                                            0, m).cpu().numpy().T
                assert_allclose(got[keep], want, rtol=1e-7, atol=1e-9, err_msg=f"linear_fit n={n} m={m}")
                assert_array_equal(np.sort(got, axis=0), dist)
            else:                              # difference of independent medians / stds
                nb = int(rng.integers(2, 3000))
                c = rng.standard_normal(nb)
                f = [np.median, np.std][k % 2]
                idx = list(boot.bootstrap_indices_independent((a, c), m, seed=seed))
                src = np.sort(a) if (f is np.median and n > 28672) else a
                want = np.sort([f(src[i]) - f(c[j]) for i, j in idx])
                _, dist = boot.ci((a, c), boot.diff_of(f), n_samples=m, seed=seed, method="pi", return_dist=True,
                                  multi="independent")
                assert_allclose(dist, want, rtol=1e-10, atol=1e-12, err_msg=f"diff_of n={n} nb={nb} m={m}")


def test_fuzz_column_kernels_against_oracle():
    """(N, D) data around the column-block widths (896 / 1792 / 3584 / 7168 rows), the median
    batch limits (4095 / 8191 rows) and odd / even segment lengths: K1c, K3c, K1m, K3m and the
    BCa tail against the oracle on the device's own index sets."""
    import warnings
    rng = np.random.default_rng(77)
    shapes = [(2, 1), (3, 2), (31, 33), (32, 5), (33, 64), (63, 3), (64, 2), (65, 1), (895, 3), (896, 34), (897, 2),
              (1792, 17), (1793, 2), (3584, 9), (3585, 2), (4094, 2), (4095, 3), (4096, 2), (7168, 5), (7169, 2),
              (8191, 2), (8192, 2), (12000, 3)] + [(int(rng.integers(2, 9000)), int(rng.integers(1, 40))) for _ in range(8)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, (n, dcols) in enumerate(shapes):
            data = np.round(rng.standard_normal((n, dcols)) * rng.uniform(0.5, 2, dcols) + rng.uniform(-2, 2, dcols),
                            3 if k % 3 == 0 else 12)                       # some with ties
            m = int(rng.integers(2, 30))
            seed = int(rng.integers(0, 2**62))
            idx = np.array(list(boot.bootstrap_indices(data, m, seed=seed))).reshape(m, n)
            name = ["mean_axis0", "median_axis0", "std_axis0"][k % 3]
            f = {"mean_axis0": boot.mean_axis0, "median_axis0": boot.median_axis0, "std_axis0": boot.std_axis0}[name]
            g = {"mean_axis0": lambda t: np.mean(t, axis=0), "median_axis0": lambda t: np.median(t, axis=0),
                 "std_axis0": lambda t: np.std(t, axis=0)}[name]
            want = orc.stat_over_indices((data,), g, idx)
            want.sort(axis=0)
            _, dist = boot.ci(data, f, n_samples=m, seed=seed, method="pi", return_dist=True)
            msg = f"{name} shape=({n},{dcols}) m={m}"
            if name == "median_axis0":
                assert_array_equal(dist, want, err_msg=msg)
            else:
                assert_allclose(dist, want, rtol=RTOL, atol=1e-13, err_msg=msg)
            # jackknife pieces per column (K3c / K3m) against the closed forms
            jk = {"mean_axis0": orc.jack_mean, "median_axis0": orc.jack_median,
                  "std_axis0": lambda x: orc.jack_var(x, True)}[name]
            spec = boot._registry.require(f)
            jd = (_device.jackknife_median_columns(dev(data)) if name == "median_axis0"
                  else _device.jackknife_columns(dev(data), spec.stat_id)).cpu().numpy()
            for d in range(0, dcols, max(1, dcols // 3)):
                j = jk(data[:, d])
                assert_allclose(jd[d, 0], g(data)[d], rtol=1e-12, atol=1e-14, err_msg=msg)
                assert_allclose(jd[d, 1], j.mean(), rtol=RTOL, atol=1e-13, err_msg=msg)
                d2 = np.sum((j.mean() - j) ** 2)
                if d2 > 1e-18:
                    assert_allclose(jd[d, 2], d2, rtol=1e-6, err_msg=msg)


def test_cfg1_matches_oracle_end_to_end():
    """BASELINE cfg1: N=1000 f64, np.mean, BCa, n_samples=10000, seed=0."""
    x = np.random.default_rng(0).standard_normal(1000)
    fused_vs_oracle((x,), "mean", 10000, seed=0)
    # statistical agreement with the reference's PCG64 answer [-0.10978462, 0.01280137]
    out = boot.ci(x, np.mean, method="bca", n_samples=10000, seed=0)
    se = x.std() / np.sqrt(1000)
    assert_allclose(out, [-0.10978462, 0.01280137], atol=0.15 * se)


@pytest.mark.parametrize("stat_name,n,n_arrays", [("mean", 1000, 1), ("sum", 7, 1), ("var", 1000, 1), ("std", 28672, 1),
                                                  ("ratio_of_means", 3000, 2), ("weighted_average", 50, 2),
                                                  ("pearson_r", 14336, 2)])
def test_small_call_path_equals_general_path(stat_name, n, n_arrays, monkeypatch):
    """Short jobs run as one library call (K1 + single-CTA jackknife + single-CTA tail whose ranks
    the host verifies); they must return exactly what the general path returns."""
    rng = np.random.default_rng(n)
    arrays = tuple(rng.gamma(2, 1, n) + 0.25 * k for k in range(n_arrays))
    data = arrays if n_arrays > 1 else arrays[0]
    f = DEVICE_STATS[stat_name]
    calls = []
    real = _device.ci_small
    monkeypatch.setattr(_device, "ci_small", lambda *a, **k: calls.append(1) or real(*a, **k))
    kw = [dict(method="bca"), dict(method="pi"), dict(method="bca", output="errorbar"),
          dict(method="pi", output="errorbar", alpha=(0.1, 0.5, 0.9)), dict(method="bca", alpha=0.5, n_samples=32768)]
    small = [boot.ci(data, f, seed=7, **{"n_samples": 4000, **k}) for k in kw]
    assert len(calls) == len(kw)
    monkeypatch.setattr(_device, "ci_small_supported", lambda *a: False)
    general = [boot.ci(data, f, seed=7, **{"n_samples": 4000, **k}) for k in kw]
    assert len(calls) == len(kw)
    for s, g in zip(small, general):
        assert_array_equal(s, g)


def _edge_data(kind):
    rng = np.random.default_rng(11)
    if kind == "nan":
        x = rng.standard_normal(300)
        x[5] = np.nan                     # most resamples draw it: the statistic vector is mostly NaN
        return x
    if kind == "rare_nan":
        x = rng.standard_normal(3000)
        x[7] = np.nan
        return x
    if kind == "inf":
        x = rng.standard_normal(2000)
        x[3] = np.inf
        return x
    if kind == "ties":
        return rng.integers(0, 2, 20).astype(float)      # ~21 distinct means: crowded buckets
    if kind == "huge":
        return rng.standard_normal(500) * 1e300
    if kind == "tiny":
        return rng.standard_normal(500) * 1e-310          # subnormal range
    if kind == "signed_zero":
        return np.array([0.0, -0.0] * 8)
    if kind == "two_values":
        return np.array([1.0] * 30 + [1.0 + 2.0 ** -52])
    return rng.standard_normal(1000)


@pytest.mark.parametrize("kind", ["plain", "nan", "rare_nan", "inf", "ties", "huge", "tiny", "signed_zero", "two_values"])
def test_small_call_select_edge_cases(kind, monkeypatch):
    """The fused small call selects by a monotone bucket map with an exact ranking inside the
    bucket and falls back to the radix descent for degenerate vectors: in every case the
    interval is the general path's (sort(stat)[k] bit for bit), from numpy and from CUDA input."""
    import warnings
    x = _edge_data(kind)
    kws = [dict(method="bca", n_samples=4000), dict(method="pi", n_samples=1), dict(method="pi", n_samples=2),
           dict(method="bca", n_samples=3), dict(method="pi", n_samples=20000, alpha=(0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.999)),
           dict(method="bca", n_samples=32768), dict(method="pi", n_samples=16384, output="errorbar"),
           dict(method="bca", n_samples=16385)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in (np.mean, np.std):
            small = [boot.ci(x, f, seed=3, **k) for k in kws]
            small_dev = [boot.ci(dev(x), f, seed=3, **k) for k in kws]
            monkeypatch.setattr(_device, "ci_small_supported", lambda *a: False)
            general = [boot.ci(x, f, seed=3, **k) for k in kws]
            monkeypatch.undo()
            for a, b, c in zip(small, small_dev, general):
                assert_array_equal(a, c)
                assert_array_equal(b, c)


def test_small_call_is_reentrant_on_its_buffers():
    """The fused kernel's ticket counter is left at zero by every call: many calls in a row, of
    different shapes, from one thread."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(777)
    want = {n: boot.ci(x, np.mean, n_samples=n, seed=9) for n in (10, 500, 9000)}
    for _ in range(20):
        for n in (9000, 10, 500):
            assert_array_equal(boot.ci(x, np.mean, n_samples=n, seed=9), want[n])


@pytest.mark.parametrize("stat_name,n,n_arrays", [("mean", 40000, 1), ("std", 30000, 1), ("median", 3000, 1),
                                                  ("median", 40000, 1), ("ratio_of_means", 20000, 2),
                                                  ("mean", 500, 1)])
def test_single_launch_tail_equals_general_tail(stat_name, n, n_arrays, monkeypatch):
    """Scalar statistics outside the small-call shape finish with ONE single-CTA launch (count,
    speculative levels and ranks, bucket select over up to 2^21 values) whose ranks the host
    verifies; the interval must be what the count / radix-select launches give."""
    import warnings
    rng = np.random.default_rng(n)
    arrays = tuple(rng.gamma(2, 1, n) + 0.25 * k for k in range(n_arrays))
    data = arrays if n_arrays > 1 else arrays[0]
    f = DEVICE_STATS[stat_name]
    kws = [dict(method="bca", n_samples=50000), dict(method="pi", n_samples=40000, output="errorbar"),
           dict(method="bca", n_samples=33000, alpha=(0.001, 0.01, 0.1, 0.3, 0.5, 0.7, 0.9, 0.999)),
           dict(method="bca", n_samples=70001, output="errorbar")]
    calls = []
    real = _device.ci_tail
    monkeypatch.setattr(_device, "ci_tail", lambda *a, **k: calls.append(1) or real(*a, **k))
    monkeypatch.setattr(_device, "ci_median_small_supported", lambda *a: False)   # the general job, not the one-call median
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = [boot.ci(data, f, seed=4, **k) for k in kws]
        assert len(calls) == len(kws)
        monkeypatch.setattr(_device, "ci_tail_supported", lambda *a: False)
        general = [boot.ci(data, f, seed=4, **k) for k in kws]
        assert len(calls) == len(kws)
    for a, b in zip(one, general):
        assert_array_equal(a, b)


def test_single_launch_tail_degenerate_vectors(monkeypatch):
    import warnings
    rng = np.random.default_rng(2)
    cases = [np.ones(40000), np.r_[rng.standard_normal(39999), np.nan], rng.integers(0, 2, 40000).astype(float),
             np.r_[rng.standard_normal(39999), np.inf]]
    for x in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = boot.ci(x, np.mean, n_samples=40000, seed=1)
            monkeypatch.setattr(_device, "ci_tail_supported", lambda *a: False)
            b = boot.ci(x, np.mean, n_samples=40000, seed=1)
            monkeypatch.undo()
        assert_array_equal(a, b)


def test_long_numpy_rows_travel_under_the_occupancy_chains(monkeypatch):
    """Long numpy inputs of the plain K1 path are copied by the resampling call itself, behind its
    occupancy chains (b200boot_resample_stat_host); same results as from a CUDA tensor."""
    rng = np.random.default_rng(6)
    a, b = rng.gamma(2.0, 1.0, 600_000), rng.gamma(3.0, 1.0, 600_000)
    calls = []
    real = _device.resample_stat

    def spy(*args, **kw):
        calls.append(kw.get("host_rows") is not None)
        return real(*args, **kw)
    monkeypatch.setattr(_device, "resample_stat", spy)
    for data, f, kw in [(a, np.mean, {}), (a, np.std, dict(method="pi")), ((a, b), boot.ratio_of_means, {}),
                        ((a, b), boot.pearson_r, dict(output="errorbar"))]:
        dev_data = tuple(dev(t) for t in data) if isinstance(data, tuple) else dev(data)
        before = _device.TRAFFIC["h2d"]
        got = boot.ci(data, f, n_samples=300, seed=12, **kw)
        assert calls[-1] is True
        n_arrays = len(data) if isinstance(data, tuple) else 1
        assert _device.TRAFFIC["h2d"] - before >= 8 * 600_000 * n_arrays
        want = boot.ci(dev_data, f, n_samples=300, seed=12, **kw)
        assert calls[-1] is False
        assert_array_equal(got, want)
    p = boot.pval(a, np.mean, boot.greater_than(2.0), n_samples=300, seed=3)
    assert calls[-1] is True and p == boot.pval(dev(a), np.mean, boot.greater_than(2.0), n_samples=300, seed=3)
    # short rows, read-only or strided arrays take the ordinary copy
    ro = a[:700_000].copy()
    ro.setflags(write=False)
    for data in (a[:1000], a[::2], ro):
        boot.ci(data, np.mean, n_samples=64, seed=1, method="pi")
    assert calls[-3:] == [False, False, True] or calls[-2:] == [False, True]


def test_small_pval_is_one_kernel_and_equals_general_path(monkeypatch):
    """Short pval() calls with a device criterion run the fused small-call kernel, counting the
    criterion instead of selecting order statistics; the value is the general path's."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal(800) + 0.05
    a, b = rng.gamma(2, 1, 500), rng.gamma(2.2, 1, 500)
    crits = [boot.positive, boot.less_than(0.04), boot.less_equal(0.05), boot.greater_equal(0.0),
             boot.between(0.0, 0.08)]
    calls = []
    real = _device.pval_small
    monkeypatch.setattr(_device, "pval_small", lambda *a_, **k: calls.append(1) or real(*a_, **k))
    jobs = [(x, np.mean), (x, np.std), (dev(x), np.mean), ((a, b), boot.ratio_of_means), ((a, b), boot.pearson_r)]
    small = [boot.pval(d, f, c, n_samples=3000, seed=5) for d, f in jobs for c in crits]
    assert len(calls) == len(jobs) * len(crits)
    assert all(isinstance(v, np.float64) for v in small)
    monkeypatch.setattr(_device, "ci_small_supported", lambda *a_: False)
    general = [boot.pval(d, f, c, n_samples=3000, seed=5) for d, f in jobs for c in crits]
    assert len(calls) == len(jobs) * len(crits)
    assert small == general
    assert 0.0 < small[0] < 1.0
    # what does not qualify takes the general path: host callables, medians, 2-D data
    boot.pval(x, np.mean, lambda s: s > 0, n_samples=200, seed=1)
    boot.pval(x, np.median, boot.positive, n_samples=200, seed=1)
    boot.pval(x.reshape(-1, 2), boot.mean_axis0, boot.positive, n_samples=200, seed=1)
    assert len(calls) == len(jobs) * len(crits)


@pytest.mark.parametrize("n", [1, 2, 7, 100, 1000, 2048, 2049, 10000, 28672])
def test_one_call_median_equals_general_path(n, monkeypatch):
    """ci() of np.median / a quantile of one resident column is one library call (sort, jackknife,
    K1m1 resamples, single-CTA tail); same interval as the general path and as the oracle on the
    materialised indices; NaN data fall back to the general path (numpy's rule)."""
    import warnings
    rng = np.random.default_rng(n)
    x = np.round(rng.gamma(2.0, 1.0, n), 3)
    calls = []
    real = _device.ci_median_small
    monkeypatch.setattr(_device, "ci_median_small", lambda *a, **k: calls.append(1) or real(*a, **k))
    kws = [dict(method="bca"), dict(method="pi"), dict(method="bca", output="errorbar", alpha=0.2)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in (np.median, boot.quantile(0.8)):
            one = [boot.ci(x, f, n_samples=700, seed=3, **k) for k in kws] + [boot.ci(dev(x), f, n_samples=700, seed=3)]
            assert len(calls) == len(kws) + 1
            monkeypatch.setattr(_device, "ci_median_small_supported", lambda *a: False)
            general = [boot.ci(x, f, n_samples=700, seed=3, **k) for k in kws] + [boot.ci(dev(x), f, n_samples=700, seed=3)]
            monkeypatch.undo()
            monkeypatch.setattr(_device, "ci_median_small", lambda *a, **k: calls.append(1) or real(*a, **k))
            for a, b in zip(one, general):
                assert_array_equal(a, b)
            calls.clear()
        if n >= 7:
            y = x.copy()
            y[n // 2] = np.nan
            got = boot.ci(y, np.median, n_samples=300, seed=4, method="pi")
            idx = np.array(list(boot.bootstrap_indices(y, 300, seed=4)))
            want = orc.ci_from_indices((y,), np.median, idx, (0.025, 0.975), 300, "pi")
            assert_array_equal(got, want)


def test_small_call_wrong_speculative_ranks_are_caught(monkeypatch):
    """The device's ranks are only a guess: when the host's own levels give other ranks the
    order statistics are selected again with the host's."""
    x = np.random.default_rng(3).standard_normal(500)
    want = boot.ci(x, np.mean, n_samples=2000, seed=1)
    real = _device.ci_small

    def skewed(*a, **k):
        stat, vals, ranks, less, jack = real(*a, **k)
        return stat, vals + 1.0, ranks + 1, less, jack
    monkeypatch.setattr(_device, "ci_small", skewed)
    assert_array_equal(boot.ci(x, np.mean, n_samples=2000, seed=1), want)
    # degenerate data: every level is NaN, ranks fall back to 0 (bootstrap.py:480)
    monkeypatch.setattr(_device, "ci_small", real)
    with pytest.warns(boot.InstabilityWarning):
        out = boot.ci(np.ones(50), np.mean, n_samples=300, seed=2)
    assert_array_equal(out, [1.0, 1.0])


def test_small_call_does_not_apply_where_it_must_not(monkeypatch):
    calls = []
    real = _device.ci_small
    monkeypatch.setattr(_device, "ci_small", lambda *a, **k: calls.append(1) or real(*a, **k))
    x = np.random.default_rng(4).standard_normal(40000)             # tiled plan
    boot.ci(x, np.mean, n_samples=500, seed=1)
    boot.ci(x[:1000], np.mean, n_samples=40000, seed=1)              # statistic vector too long
    boot.ci(x[:1000], np.median, n_samples=500, seed=1)              # rank statistic
    boot.ci(x[:1000], np.mean, n_samples=500, seed=1, return_dist=True)
    boot.ci(x[:1000].reshape(-1, 2), boot.mean_axis0, n_samples=500, seed=1)
    assert not calls
    boot.ci(x[:1000], np.mean, n_samples=500, seed=1)
    assert len(calls) == 1


# ------------------------------------------------------------------ K3 -----------

@pytest.mark.parametrize("stat_name", ["mean", "var", "std", "ratio_of_means", "weighted_average",
                                       "pearson_r"])
@pytest.mark.parametrize("n", [5, 301, 5000])
def test_jackknife_closed_form(stat_name, n):
    rng = np.random.default_rng(n)
    a = rng.gamma(2, 1, n)
    b = 0.3 * a + rng.gamma(3, 1, n)
    tdata = (a, b) if DEVICE_STATS[stat_name] in (boot.ratio_of_means, boot.weighted_average,
                                                  boot.pearson_r) else (a,)
    spec = boot._registry.require(DEVICE_STATS[stat_name])
    out = _device.jackknife([dev(t) for t in tdata], spec.stat_id).cpu().numpy()
    jstat = JACK[stat_name](*tdata)
    d = jstat.mean() - jstat
    assert_allclose(out[0], STATS[stat_name](*tdata), rtol=1e-12)
    assert_allclose(out[1], jstat.mean(), rtol=1e-12)
    assert_allclose(out[2], np.sum(d ** 2), rtol=1e-7)
    assert_allclose(out[4], orc.acceleration(jstat), rtol=1e-6, atol=1e-12)
    if n <= 301:   # and against the reference's O(N^2) loop itself
        loop = orc.jackknife_stats_loop(tdata, STATS[stat_name])
        assert_allclose(out[4], orc.acceleration(loop), rtol=1e-6, atol=1e-12)


# ------------------------------------------------------------------ K4 / K5 / sort

def _stat_matrix(n_cols, n, seed=0):
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((n_cols, n))
    m[0, : n // 3] = np.round(m[0, : n // 3], 1)          # ties
    if n > 10:
        m[-1, 3] = np.nan
        m[-1, 5] = -0.0
        m[-1, 6] = 0.0
        m[-1, 7] = np.inf
        m[-1, 8] = -np.inf
    return m


@pytest.mark.parametrize("n_cols,n", [(1, 1), (1, 10), (1, 100000), (3, 4097), (70, 1000)])
def test_select_equals_sorted_pick(n_cols, n):
    m = _stat_matrix(n_cols, n)
    srt = np.sort(m, axis=1)
    rng = np.random.default_rng(1)
    for n_alphas in (1, 2, 11):
        ranks = rng.integers(0, n, (n_alphas, n_cols))
        ranks[0, 0] = 0
        ranks[-1, -1] = n - 1
        got = _device.Select(dev(m), dev(ranks)).run().cpu().numpy()
        want = np.stack([srt[np.arange(n_cols), ranks[a]] for a in range(n_alphas)])
        assert_array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert_array_equal(got[ok], want[ok])


def test_select_with_allreduce_hook_emulating_two_ranks():
    """Sharded select: histograms of two shards are summed between hist and pick."""
    m = _stat_matrix(2, 9000)
    ranks = np.array([[0, 10], [4500, 8999], [123, 7777]])
    srt = np.sort(m, axis=1)
    a, b = dev(m[:, :4000].copy()), dev(m[:, 4000:].copy())
    sa, sb = _device.Select(a, dev(ranks)), _device.Select(b, dev(ranks))
    L, s = _abi.lib(), _device.stream_ptr()
    for sel in (sa, sb):
        _abi.check(L.b200boot_select_begin(sel.ranks.data_ptr(), 3, 2, sel.ws, sel.wsz, s))
    for p in range(8):
        for sel in (sa, sb):
            _abi.check(L.b200boot_select_hist(sel.stat.data_ptr(), sel.n_local, 2, 3, p, sel.ws, sel.wsz, s))
        tot = sa.hist_tensor() + sb.hist_tensor()
        sa.hist_tensor().copy_(tot)
        sb.hist_tensor().copy_(tot)
        for sel in (sa, sb):
            _abi.check(L.b200boot_select_pick(3, 2, p, sel.ws, sel.wsz, s))
    out = torch().empty((3, 2), dtype=torch().float64, device="cuda")
    _abi.check(L.b200boot_select_end(3, 2, out.data_ptr(), sa.ws, sa.wsz, s))
    want = np.stack([srt[np.arange(2), ranks[k]] for k in range(3)])
    assert_array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("op,fn", [(0, np.less), (1, np.less_equal), (2, np.greater), (3, np.greater_equal)])
def test_count_reduction(op, fn):
    m = _stat_matrix(5, 33333)
    t0 = np.array([0.0, 0.1, -0.5, 3.0, 0.0])
    got = _device.count(dev(m), op, dev(t0)).cpu().numpy()
    with np.errstate(invalid="ignore"):
        assert_array_equal(got, fn(m, t0[:, None]).sum(axis=1))
    got = _device.count(dev(m), 4, dev(t0 - 0.3), dev(t0 + 0.3)).cpu().numpy()
    with np.errstate(invalid="ignore"):
        assert_array_equal(got, ((m >= (t0 - 0.3)[:, None]) & (m <= (t0 + 0.3)[:, None])).sum(axis=1))


@pytest.mark.parametrize("n_cols,n", [(1, 2), (1, 1000), (1, 2048), (1, 2049), (2, 100000), (9, 5000)])
def test_sort_columns(n_cols, n):
    m = _stat_matrix(n_cols, n, seed=2)
    got = _device.sort_columns(dev(m)).cpu().numpy()
    want = np.sort(m, axis=1)
    assert_array_equal(np.isnan(got), np.isnan(want))
    assert_array_equal(got[~np.isnan(want)], want[~np.isnan(want)])


# ------------------------------------------------------------------ API behaviour -

def test_reference_property_tests():
    """Seed-independent properties from tests/bootstrap_test.py:171-195, 365-370."""
    out, dist = boot.ci(DATA20, return_dist=True, method="pi", alpha=[0.1, 0.9], n_samples=100)
    assert out[0] == dist[10] and out[1] == dist[89]
    assert_allclose(boot.ci(DATA20, seed=SEED), boot.ci(DATA20, alpha=(0.025, 1 - 0.025), seed=SEED))
    assert_allclose(boot.ci(DATA20, method="pi", seed=SEED),
                    boot.ci(DATA20, method="pi", alpha=(0.025, 1 - 0.025), seed=SEED))
    a = boot.ci(DATA20, seed=SEED)
    b = boot.ci(DATA20, output="errorbar", seed=SEED)
    assert b.shape == (2, 1)
    assert_allclose(b.T, abs(np.average(DATA20) - a)[np.newaxis])


def test_abc_for_the_weighted_mean():
    """method='abc' (bootstrap.py:507-567; tests/bootstrap_test.py:140-150, 190-195, 372-376).
    The device evaluates the reference's 2N + 3 + A weighted means with numpy's own arithmetic
    (pairwise summation order, single IEEE operations), so the finite differences — rounding
    noise included — and with them the interval agree with the oracle's faithful restatement
    for any epsilon, the default one included."""
    import pandas as pd
    got = boot.ci(DATA20, method="abc", seed=SEED)
    assert_allclose(got, orc.ci_abc(DATA20), rtol=RTOL)
    assert_allclose(got, [0.20982275, 1.20374686], rtol=1e-7)                     # the reference's golden
    got = boot.ci(DATA20, alpha=(0.1, 0.2, 0.8, 0.9), method="abc")
    assert_allclose(got, orc.ci_abc(DATA20, alpha=(0.1, 0.2, 0.8, 0.9)), rtol=RTOL)
    assert_allclose(got, [0.39472915, 0.51161304, 0.93789723, 1.04407254], rtol=1e-7)
    eb = boot.ci(DATA20, output="errorbar", method="abc")
    assert eb.shape == (2, 1)
    assert_allclose(eb, orc.ci_abc(DATA20, output="errorbar"), rtol=1e-11)
    assert_allclose(eb.T, abs(np.average(DATA20) - boot.ci(DATA20, method="abc"))[np.newaxis])
    assert_allclose(boot.ci(pd.Series(DATA20, index=np.arange(50, 70)), method="abc"), boot.ci(DATA20, method="abc"))
    rng = np.random.default_rng(0)
    for n, eps in ((7, 0.001), (128, 0.001), (129, 0.01), (500, 0.001), (500, 0.1), (1000, 0.001), (3001, 1e-4)):
        x = rng.gamma(2, 1, n)
        assert_allclose(boot.ci(x, method="abc", epsilon=eps), orc.ci_abc(x, epsilon=eps), rtol=RTOL), (n, eps)
    # the weighted means themselves: numpy's bits
    x = rng.standard_normal(777)
    w = rng.random((5, 777))
    got = _device.weighted_averages(dev(x), w)
    assert_array_equal(got, [np.average(x, weights=wi) for wi in w])
    assert _device.weighted_averages(dev(x), None)[0] == np.average(x)
    with pytest.raises(ValueError):
        boot.ci(DATA20, method="abc", return_dist=True)
    with pytest.raises(TypeError):
        boot.ci(DATA20, np.mean, method="abc")


def test_warnings_match_reference():
    """tests/bootstrap_test.py:250-260."""
    with pytest.warns(boot.InstabilityWarning, match="(NaN|all equal)"):
        boot.ci(np.ones(20), seed=SEED)
    with pytest.warns(boot.InstabilityWarning, match="extremal"):
        boot.ci([1, 2], n_samples=10, seed=SEED)
    with pytest.warns(boot.InstabilityWarning, match="top 10"):     # seed-independent under 'pi'
        boot.ci(np.arange(1, 20), n_samples=50, seed=SEED, method="pi")
    with pytest.warns(boot.InstabilityWarning, match="top 10"):     # the reference's own draws
        boot.ci_indexed(np.arange(1, 20), None, pcg_indices(19, 50))


def test_edge_shapes():
    """Smallest inputs, single resample, odd alpha shapes, non-finite data."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = boot.ci([3.5], np.mean, n_samples=5, seed=1, method="pi")
        assert_array_equal(out, [3.5, 3.5])
        out, dist = boot.ci([1.0, 2.0, 4.0], np.mean, n_samples=1, seed=1, method="pi", return_dist=True)
        assert dist.shape == (1,) and out[0] == out[1] == dist[0]
        out = boot.ci(DATA20, np.mean, alpha=[0.5], n_samples=101, seed=2, method="pi")
        assert out.shape == (1,)
        a2 = np.array([[0.1, 0.9], [0.2, 0.8]])
        out = boot.ci(DATA20, np.mean, alpha=a2, n_samples=200, seed=2, method="pi")
        flat = boot.ci(DATA20, np.mean, alpha=a2.ravel(), n_samples=200, seed=2, method="pi")
        assert out.shape == (2, 2) and np.array_equal(out.ravel(), flat)
        x = DATA20.copy()
        x[3] = np.nan
        out, dist = boot.ci(x, np.mean, n_samples=300, seed=3, method="pi", return_dist=True)
        k = int(np.isnan(dist).sum())
        assert 0 < k < 300 and np.all(np.isnan(dist[-k:])) and not np.isnan(dist[:-k]).any()   # NaNs sort last
        idx = np.array(list(boot.bootstrap_indices(x, 300, seed=3)))
        assert k == int(np.isnan(x[idx].mean(axis=1)).sum())
        x[3] = np.inf
        _, dist = boot.ci(x, np.mean, n_samples=300, seed=3, method="pi", return_dist=True)
        assert np.isinf(dist[-1]) and np.isfinite(dist[0])
    with pytest.raises(ValueError):
        boot.ci(DATA20, np.mean, n_samples=0)
    with pytest.raises(ValueError):
        boot.ci(np.zeros(0), np.mean)


def test_float32_data_give_float32_results_like_the_reference():
    """np.mean / np.std of float32 resamples are float32 in the reference: results
    carry the data's floating dtype and agree with the float64 run to float32 rounding."""
    rng = np.random.default_rng(8)
    x64 = rng.gamma(2, 1, 3000)
    x32 = x64.astype(np.float32)
    for f in (np.mean, np.std, np.median):
        lo32 = boot.ci(x32, f, n_samples=2000, seed=3)
        lo64 = boot.ci(x32.astype(np.float64), f, n_samples=2000, seed=3)
        assert lo32.dtype == np.float32 and lo64.dtype == np.float64
        assert_array_equal(lo32, lo64.astype(np.float32))
        out, dist = boot.ci(x32, f, n_samples=500, seed=3, method="pi", return_dist=True)
        assert out.dtype == np.float32 and dist.dtype == np.float32
    # the oracle (the reference's arithmetic) on the same float32 data and the device's own indices
    idx = np.array(list(boot.bootstrap_indices(x32, 300, seed=4)))
    want = orc.ci_from_indices((x32,), np.mean, idx, (0.025, 0.975), 300, "pi")
    got = boot.ci(x32, np.mean, n_samples=300, seed=4, method="pi")
    assert want.dtype == np.float32
    assert_allclose(got, want, rtol=2e-6)
    assert boot.ci(np.arange(50), np.mean, n_samples=200, seed=1).dtype == np.float64       # ints -> float64
    import torch
    assert boot.ci(torch.from_numpy(x32).cuda(), np.mean, n_samples=200, seed=1).dtype == np.float32
    assert boot.ci((x32, x64), boot.ratio_of_means, n_samples=200, seed=1).dtype == np.float64


def test_input_forms_agree():
    import pandas as pd
    x = DATA20
    want = boot.ci(x, np.mean, seed=5, n_samples=2000)
    assert_array_equal(boot.ci(list(x), np.mean, seed=5, n_samples=2000), want)
    assert_array_equal(boot.ci(pd.Series(x, index=np.arange(50, 70)), np.mean, seed=5, n_samples=2000), want)
    assert_array_equal(boot.ci(torch().as_tensor(x), np.mean, seed=5, n_samples=2000), want)
    assert_array_equal(boot.ci(dev(x), np.mean, seed=5, n_samples=2000), want)
    assert_array_equal(boot.ci(x, np.mean, seed=np.random.default_rng(1), n_samples=2000),
                       boot.ci(x, np.mean, seed=np.random.default_rng(1), n_samples=2000))
    xi = np.arange(1, 20)                      # integer input -> float64 statistic
    assert boot.ci(xi, np.mean, seed=1, n_samples=500).dtype == np.float64
    x32 = x.astype(np.float32)
    assert_allclose(boot.ci(x32, np.mean, seed=5, n_samples=2000),
                    boot.ci(x32.astype(np.float64), np.mean, seed=5, n_samples=2000))


def test_input_on_another_device_and_threads():
    """A CUDA tensor on a non-current device runs there; concurrent calls from Python
    threads use separate scratch."""
    import threading
    want = boot.ci(DATA20, np.mean, seed=5, n_samples=2000)
    if torch().cuda.device_count() >= 2:
        x1 = torch().as_tensor(DATA20).to("cuda:1")
        assert torch().cuda.current_device() == 0
        assert_array_equal(boot.ci(x1, np.mean, seed=5, n_samples=2000), want)
        assert torch().cuda.current_device() == 0
    big = np.random.default_rng(0).standard_normal(60000)
    ref = [boot.ci(big, np.mean, seed=s, n_samples=500, method="pi") for s in range(4)]
    got = [None] * 4

    def work(i):
        with torch().cuda.stream(torch().cuda.Stream()):
            got[i] = boot.ci(big, np.mean, seed=i, n_samples=500, method="pi")

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for a, b in zip(got, ref):
        assert_array_equal(a, b)


def test_direct_paths_from_concurrent_threads():
    """The one-call paths (fused small call, one-call median, small pval, host rows under K0) keep
    their buffers per thread and stream: four threads at once give the single-thread results."""
    import threading
    rng = np.random.default_rng(1)
    small, mid, long_rows = rng.standard_normal(900), rng.gamma(2.0, 1.0, 5000), rng.standard_normal(700_000)

    def job(seed):
        return [boot.ci(small, np.mean, n_samples=3000, seed=seed),
                boot.ci(small, np.std, n_samples=500, seed=seed, method="pi"),
                boot.ci(mid, np.median, n_samples=800, seed=seed),
                boot.ci(small, boot.quantile(0.25), n_samples=800, seed=seed, method="pi"),
                np.asarray(boot.pval(small, np.mean, boot.less_than(0.02), n_samples=2000, seed=seed)),
                boot.ci((small, small[::-1].copy()), boot.pearson_r, n_samples=600, seed=seed),
                boot.ci(long_rows, np.mean, n_samples=64, seed=seed, method="pi")]
    ref = [job(s) for s in range(4)]
    got = [None] * 4
    errors = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = job(i)
        except Exception as e:                      # noqa: BLE001 - surfaced below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for a, b in zip(got, ref):
        for u, v in zip(a, b):
            assert_array_equal(u, v)


def test_paired_tuple_equals_axis0_columns():
    """A (N, D) array under an axis-0 statistic shares one index set across columns
    (bootstrap.py:431): each column equals the 1-D run with the same seed."""
    rng = np.random.default_rng(8)
    m = rng.standard_normal((64, 5))
    for f, g in [(boot.mean_axis0, np.mean), (functools.partial(np.std, axis=0), np.std)]:
        out, dist = boot.ci(m, f, alpha=(0.1, 0.9), n_samples=1500, seed=3, return_dist=True)
        assert out.shape == (2, 5) and dist.shape == (1500, 5)
        eb = boot.ci(m, f, alpha=(0.1, 0.9), n_samples=1500, seed=3, output="errorbar")
        assert eb.shape == (5, 2)
        for d in range(5):       # same draws; the two kernels sum in different orders
            o1, d1 = boot.ci(m[:, d], g, alpha=(0.1, 0.9), n_samples=1500, seed=3, return_dist=True)
            assert_allclose(out[:, d], o1, rtol=RTOL, atol=1e-15)
            assert_allclose(dist[:, d], d1, rtol=RTOL, atol=1e-15)


@pytest.mark.parametrize("shape", [(20, 3), (896, 33), (1000, 70), (3000, 9), (7168, 5), (7169, 3)])
@pytest.mark.parametrize("name", ["mean_axis0", "std_axis0"])
def test_column_batched_moments(shape, name):
    """(N, D) data, axis-0 moment statistic: K1c / K3c (row-major blocks of W columns in
    shared memory; beyond 7168 rows a per-column loop) against the oracle on the
    device's own index sets, BCa endpoints included."""
    import warnings
    n, dcols = shape
    rng = np.random.default_rng(n + dcols)
    data = rng.standard_normal((n, dcols)) * rng.uniform(0.5, 2.0, dcols) + rng.uniform(-3, 3, dcols)
    f = {"mean_axis0": boot.mean_axis0, "std_axis0": functools.partial(np.std, axis=0)}[name]
    g = {"mean_axis0": lambda a: np.mean(a, axis=0), "std_axis0": lambda a: np.std(a, axis=0)}[name]
    m = 300
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=11)))
    want = orc.stat_over_indices((data,), g, idx)
    want.sort(axis=0)
    jk = orc.jack_mean if name == "mean_axis0" else (lambda x: orc.jack_var(x, True))
    jst = np.stack([jk(data[:, d]) for d in range(dcols)], axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out, dist = boot.ci(data, f, n_samples=m, seed=11, return_dist=True)
        ref = orc.ci_from_indices((data,), g, idx, (0.025, 0.975), m, "bca", jstat=jst)
    assert dist.shape == (m, dcols)
    assert_allclose(dist, want, rtol=RTOL, atol=1e-15)
    assert_allclose(out, ref, rtol=RTOL, atol=1e-15)
    p = boot.pval(data, f, boot.greater_than(0.5), n_samples=m, seed=11)
    assert_array_equal(p, (want > 0.5).mean(axis=0))


def test_pval_device_and_host_criteria_agree():
    x = np.array(Z, dtype=float)
    p1 = boot.pval(x, n_samples=4000, seed=2)
    p2 = boot.pval(x, compfunction=lambda s: s > 0, n_samples=4000, seed=2)
    _, dist = boot.ci(x, np.average, method="pi", n_samples=4000, seed=2, return_dist=True)
    assert p1 == p2 == np.mean(dist > 0)          # pval and ci share the stream
    assert isinstance(p1, np.float64)
    p3 = boot.pval(x, compfunction=boot.between(-2.0, -1.0), n_samples=4000, seed=2)
    assert p3 == np.mean((dist >= -2.0) & (dist <= -1.0))


def test_independent_mode():
    """multi='independent' (bootstrap.py:438-444, 601-607) against the oracle on the
    device's own per-array index streams."""
    a = np.random.default_rng(20).gamma(2, 1, 60)
    b = np.random.default_rng(21).gamma(3, 1, 45)
    n = 1500
    idx = list(boot.bootstrap_indices_independent((a, b), n, seed=11))
    f = STATS["diff_of_means"]
    for method in ("pi", "bca"):
        want = orc.ci_from_indices((a, b), f, idx, (0.025, 0.975), n, method, multi="independent")
        got = boot.ci((a, b), boot.diff_of_means, n_samples=n, multi="independent", seed=11, method=method)
        assert_allclose(got, want, rtol=RTOL)
    # statistical agreement with the reference golden (tests/bootstrap_test.py:240-248)
    got = boot.ci((X, Z), boot.diff_of_means, n_samples=20000, multi="independent", seed=1)
    assert_allclose(got, [2.547619, 7.97619], atol=0.25)


@pytest.mark.parametrize("name,stat", [("mean", np.mean), ("median", np.median), ("std", np.std), ("var", np.var),
                                       ("sum", np.sum), ("q80", functools.partial(np.quantile, q=0.8))])
@pytest.mark.parametrize("na,nb", [(60, 45), (301, 300), (30000, 1000)])
def test_independent_difference_of_statistics(name, stat, na, nb):
    """diff_of(s): s(a) - s(b) of independent samples; the pooled jackknife (N_a + N_b
    leave-one-out values, bootstrap.py:601-607) comes from the two per-array K3 summaries."""
    a = np.random.default_rng(na).gamma(2, 1, na)
    b = np.random.default_rng(nb + 1).gamma(3, 1, nb)
    n = 600
    f = lambda u, v: stat(u) - stat(v)
    dstat = boot.diff_of(stat)
    idx = list(boot.bootstrap_indices_independent((a, b), n, seed=13))
    if name in ("median", "q80") and na > 28672:      # rank-space path: indices address sorted(a)
        a = np.sort(a)
    want_dist = np.sort([f(a[i], b[j]) for i, j in idx])
    out, dist = boot.ci((a, b), dstat, n_samples=n, multi="independent", seed=13, method="pi", return_dist=True)
    assert_allclose(dist, want_dist, rtol=RTOL, atol=1e-13)
    # closed-form per-array leave-one-out values, pooled array-major as the reference does
    jack1 = {"mean": orc.jack_mean, "median": orc.jack_median, "std": lambda x: orc.jack_var(x, True),
             "var": orc.jack_var, "sum": lambda x: np.sum(x) - x, "q80": lambda x: orc.jack_quantile(x, 0.8)}[name]
    jstat = np.concatenate([jack1(a) - stat(b), stat(a) - jack1(b)])
    if na <= 301:
        assert_allclose(jstat, orc.jackknife_stats_independent_loop((a, b), f), rtol=1e-9, atol=1e-12)
    want = orc.ci_from_indices((a, b), f, idx, (0.025, 0.975), n, "bca", multi="independent", jstat=jstat)
    got = boot.ci((a, b), dstat, n_samples=n, multi="independent", seed=13, method="bca")
    assert_allclose(got, want, rtol=RTOL, atol=1e-13)
    p = boot.pval((a, b), dstat, boot.less_than(0.0), n_samples=n, multi="independent", seed=13)
    assert p == np.mean(dist < 0.0)


OPS = {"sub": lambda u, v: u - v, "add": lambda u, v: u + v, "div": lambda u, v: u / v,
       "mul": lambda u, v: u * v, "log_ratio": lambda u, v: np.log(u / v)}


@pytest.mark.parametrize("op", ["div", "log_ratio", "mul", "add", "sub"])
@pytest.mark.parametrize("sa,sb,na,nb", [(np.mean, np.mean, 60, 45), (np.mean, np.std, 301, 120),
                                         (np.median, np.mean, 200, 77), (np.var, np.median, 64, 1000)])
def test_independent_general_two_sample_statistic(op, sa, sb, na, nb):
    """f(s_a(a), s_b(b)) for independent samples, any registered f: the bootstrap distribution
    against the oracle on the device's own per-array index streams, and the pooled jackknife
    (N_a + N_b values, bootstrap.py:601-607, 708-714) against the reference's own loop
    (oracle.jackknife_stats_independent_loop) to 1e-12."""
    a = np.random.default_rng(na).gamma(2, 1, na) + 1.0
    b = np.random.default_rng(nb + 7).gamma(3, 1, nb) + 1.0
    f = lambda u, v: OPS[op](sa(u), sb(v))
    dstat = boot.independent(op, sa, sb)
    assert_allclose(dstat(a, b), f(a, b), rtol=1e-15)
    n = 500
    idx = list(boot.bootstrap_indices_independent((a, b), n, seed=17))
    for method in ("pi", "bca"):
        want = orc.ci_from_indices((a, b), f, idx, (0.05, 0.95), n, method, multi="independent")
        got = boot.ci((a, b), dstat, alpha=(0.05, 0.95), n_samples=n, multi="independent", seed=17, method=method)
        assert_allclose(got, want, rtol=RTOL, atol=1e-14)
    # the pooled jackknife itself, through the C ABI
    jstat = orc.jackknife_stats_independent_loop((a, b), f)
    parts = []
    for x, s in ((a, sa), (b, sb)):
        spec = boot._registry.require(s)
        xd = dev(x)
        if spec.stat_id == _abi.STAT_MEDIAN:
            parts.append(_device.jackknife_values_sorted(xd.sort().values))
        else:
            parts.append(_device.jackknife_values([xd], spec.stat_id))
    (va, ja), (vb, jb) = parts
    assert_allclose(np.sort(va.cpu().numpy()), np.sort([sa(np.delete(a, i)) for i in range(na)]), rtol=RTOL)
    pooled = _device.jackknife_pooled(dstat.spec.op, va, ja[0:1], vb, jb[0:1]).cpu().numpy()[0]
    assert_allclose(pooled[0], f(a, b), rtol=RTOL)
    assert_allclose(pooled[1], jstat.mean(), rtol=RTOL)
    assert_allclose(pooled[4], orc.acceleration(jstat), rtol=1e-9, atol=1e-15)


def test_independent_ratio_of_means_matches_reference_style_run():
    """ratio of the means of two independent samples of different lengths: endpoints bracket the
    plain ratio, pval counts on the device, sharding-free determinism under the seed."""
    rng = np.random.default_rng(3)
    a, b = rng.gamma(4, 1, 5000), rng.gamma(2, 1, 40000)           # b spans several cells
    st = boot.ratio_of(np.mean)
    lo, hi = boot.ci((a, b), st, n_samples=4000, multi="independent", seed=5)
    assert lo < a.mean() / b.mean() < hi
    assert_array_equal([lo, hi], boot.ci((a, b), st, n_samples=4000, multi="independent", seed=5))
    p = boot.pval((a, b), st, boot.greater_than(2.0), n_samples=4000, multi="independent", seed=5)
    _, dist = boot.ci((a, b), st, n_samples=4000, multi="independent", seed=5, method="pi", return_dist=True)
    assert p == np.mean(dist > 2.0)
    with pytest.raises(NotImplementedError):
        boot.ci((a, a), st, n_samples=10)                            # independent statistic without the mode
    with pytest.raises(NotImplementedError):
        boot.independent("div", boot.ratio_of_means)                 # inner statistics take one array
    with pytest.raises(ValueError):
        boot.independent("pow", np.mean)


# ------------------------------------------------------------------ other plans --

@pytest.mark.parametrize("n,L,wrap", [(5, 3, False), (4, 3, True), (1000, 7, False), (1000, 7, True), (40001, 50, True)])
def test_moving_block_indices(n, L, wrap):
    """bootstrap.py:747-787 on the device stream: bit-equal to the oracle's restatement
    of that stream, same block structure as the reference's generator."""
    got = np.array(list(boot.bootstrap_indices_moving_block(np.zeros(n), 9, block_length=L, wrap=wrap, seed=77)))
    assert got.shape == (9, n) and got.dtype == np.int64
    for r in (0, 4, 8):
        assert_array_equal(got[r], orc.device_moving_block(77, r, n, L, wrap))
    assert got.min() >= 0 and got.max() < n
    body = got[:, : (n // L) * L].reshape(9, -1, L)
    step = np.diff(body, axis=2)
    assert np.all((step == 1) | (step == 1 - n))         # consecutive inside a block (mod N)
    if not wrap:
        assert body[:, :, 0].max() < n - L                # numpy's half-open start range
    with pytest.raises(ValueError):
        list(boot.bootstrap_indices_moving_block(np.zeros(3), 2, block_length=3))


@pytest.mark.parametrize("n,L,wrap", [(1000, 7, False), (1000, 7, True), (5000, 64, False), (40001, 50, True),
                                      (300, 1, False), (64, 100, True)])
def test_moving_block_plan_fused_with_the_statistic(n, L, wrap):
    """K1b: block starts, L-contiguous gathers and the moment statistic in one kernel; the
    resamples are exactly those of bootstrap_indices_moving_block (bootstrap.py:747-787), so the
    statistics equal the oracle's over those indices."""
    rng = np.random.default_rng(n + L)
    x = np.cumsum(rng.standard_normal(n)) * 0.1 + 2.0                 # dependent data
    y = 0.5 * x + rng.standard_normal(n)
    m = 200
    idx = np.array(list(boot.bootstrap_indices_moving_block(x, m, block_length=L, wrap=wrap, seed=5)))
    for stat_name, tdata in (("mean", (x,)), ("std", (x,)), ("pearson_r", (x, y)), ("ratio_of_means", (x, y))):
        data = tdata if len(tdata) > 1 else tdata[0]
        want = np.sort(orc.stat_over_indices(tdata, STATS[stat_name], idx))
        out, dist = boot.ci_moving_block(data, DEVICE_STATS[stat_name], n_samples=m, block_length=L, wrap=wrap,
                                         seed=5, return_dist=True, multi="paired" if len(tdata) > 1 else None)
        assert_allclose(dist, want, rtol=RTOL, atol=1e-14)
        ref = orc.ci_from_indices(tdata, STATS[stat_name], idx, (0.025, 0.975), m, "pi")
        assert_allclose(out, ref, rtol=RTOL, atol=1e-14)
    # shards reproduce the whole: resample r depends on (seed, r) only
    flat = [dev(x)]
    whole = _device.resample_stat_moving_block(flat, _abi.STAT_MEAN, 5, L, wrap, 0, m).cpu().numpy()
    parts = np.concatenate([_device.resample_stat_moving_block(flat, _abi.STAT_MEAN, 5, L, wrap, a, b).cpu().numpy()
                            for a, b in ((0, 33), (33, 34), (34, m))])
    assert_array_equal(whole, parts)
    with pytest.raises(NotImplementedError):
        boot.ci_moving_block(x, np.median, n_samples=10, block_length=L, wrap=wrap)


def test_moving_block_plan_rejects_empty_start_range():
    with pytest.raises(ValueError, match="low >= high"):
        boot.ci_moving_block(np.arange(5.0), np.mean, block_length=5, wrap=False)


def test_moving_block_starts_are_uniform():
    n, L = 64, 4
    got = np.array(list(boot.bootstrap_indices_moving_block(np.zeros(n), 4000, block_length=L, wrap=True, seed=3)))
    starts = got[:, ::L].ravel()
    cnt = np.bincount(starts, minlength=n)
    chi2 = ((cnt - cnt.mean()) ** 2 / cnt.mean()).sum()
    assert chi2 < 130                                      # dof 63


def test_subsample_indices():
    """bootstrap.py:717-744 / bootstrap_test.py:114-138."""
    idx = boot.subsample_indices(DATA20, 1000, 0.5, seed=5)
    assert idx.shape == (1000, 10) and idx.dtype == np.int64
    assert all(len(np.unique(r)) == 10 for r in idx)
    for r in (0, 500, 999):
        assert_array_equal(idx[r], orc.device_subsample(5, r, 20, 10))
    assert all(len(np.unique(r)) == 10 for r in boot.subsample_indices(DATA20, 50, 10))
    perm = boot.subsample_indices(np.arange(0, 50), 300, -1, seed=1)
    assert perm.shape == (300, 50) and np.all(np.sort(perm, axis=1) == np.arange(50))
    assert not np.all(perm[0] == perm[1:])
    # every position equally likely to hold every row
    first = np.bincount(perm[:, 0], minlength=50)
    big = boot.subsample_indices(np.arange(0, 50), 20000, 1, seed=2)[:, 0]
    cnt = np.bincount(big, minlength=50)
    assert ((cnt - 400.0) ** 2 / 400.0).sum() < 110        # dof 49
    assert first.sum() == 300
    with pytest.raises(ValueError):
        boot.subsample_indices(DATA20, 1000, 30)
    with pytest.raises(ValueError, match="size cannot be -5"):
        boot.subsample_indices(DATA20, size=-5)
    big = boot.subsample_indices(np.zeros(5000), 3, 0.5, seed=9)      # beyond one shared-memory segment
    assert big.shape == (3, 2500) and all(len(np.unique(r)) == 2500 for r in big)
    assert_array_equal(big[2], orc.device_subsample(9, 2, 5000, 2500))


# ------------------------------------------------------------------ sharding -----

def test_shards_reproduce_the_single_gpu_statistics():
    """Emulates 1/2/4/8 ranks on one GPU: the union of the shards' statistic vectors
    is bit-identical to the unsharded vector."""
    from scikits_bootstrap_b200 import _dist
    x = dev(np.random.default_rng(1).standard_normal(100000) + 1.0)
    n = 203
    full = _device.resample_stat([x], _abi.STAT_MEAN, 42, 0, n).cpu().numpy()
    for world in (2, 4, 8):
        parts = []
        for rank in range(world):
            lo, hi = _dist.shard_range(n, rank, world)
            parts.append(_device.resample_stat([x], _abi.STAT_MEAN, 42, lo, hi).cpu().numpy())
        assert_array_equal(np.concatenate(parts), full)


def test_workspace_budget_chunks_do_not_change_results(monkeypatch):
    """Large jobs are run as consecutive resample sub-ranges under a scratch budget."""
    x = dev(np.random.default_rng(2).standard_normal(70000))
    full = _device.resample_stat([x], _abi.STAT_MEAN, 3, 0, 900).cpu().numpy()
    monkeypatch.setenv("B200BOOT_WS_BYTES", str(4 << 20))
    part = _device.resample_stat([x], _abi.STAT_MEAN, 3, 0, 900).cpu().numpy()
    assert_array_equal(part, full)


# ------------------------------------------------------------------ statistics ---

def test_philox_path_agrees_statistically_with_pcg64_path():
    """Two-sample KS between the device's bootstrap distribution and the oracle's
    PCG64 one (the reference compares its own two generators at rtol=1e-2,
    tests/bootstrap_test.py:51-58)."""
    from scipy import stats as sps
    x = np.random.default_rng(3).gamma(2.0, 1.0, 200)
    n = 20000
    _, d_gpu = boot.ci(x, np.mean, method="pi", n_samples=n, seed=1, return_dist=True)
    d_ref = orc.stat_over_indices((x,), np.mean, orc.bootstrap_indices_pcg64(200, n, 2))
    assert sps.ks_2samp(d_gpu, d_ref).pvalue > 1e-4
    assert_allclose(boot.ci(x, np.mean, n_samples=n, seed=1), orc.ci(x, np.mean, n_samples=n, seed=2), rtol=1e-2)


@pytest.mark.parametrize("stat_name", ["var", "std", "median", "ratio_of_means", "weighted_average", "pearson_r"])
def test_every_registry_statistic_agrees_statistically_with_pcg64(stat_name):
    """Two-sample KS of the device's bootstrap distribution against the oracle's PCG64 one,
    per registry statistic (the mean is covered above)."""
    from scipy import stats as sps
    rng = np.random.default_rng(31)
    n_obs, n = 150, 6000
    a = rng.gamma(2.0, 1.0, n_obs)
    b = 0.4 * a + rng.gamma(3.0, 1.0, n_obs)
    paired = stat_name in ("ratio_of_means", "weighted_average", "pearson_r")
    tdata = (a, b) if paired else (a,)
    data = tdata if paired else a
    _, d_gpu = boot.ci(data, DEVICE_STATS[stat_name], method="pi", n_samples=n, seed=7, return_dist=True)
    d_ref = orc.stat_over_indices(tdata, STATS[stat_name], orc.bootstrap_indices_pcg64(n_obs, n, 8))
    assert sps.ks_2samp(d_gpu, d_ref).pvalue > 1e-4
    assert abs(d_gpu.mean() - d_ref.mean()) < 6 * d_ref.std() / np.sqrt(n / 2)


def test_tiled_path_agrees_statistically_with_pcg64():
    """Same for N spanning several cells and bank classes (tile-multinomial draws), mean and std."""
    from scipy import stats as sps
    x = np.random.default_rng(32).gamma(2.0, 1.0, 40000)
    n = 3000
    idx = orc.bootstrap_indices_pcg64(40000, n, 9)
    ref = np.array([[x[i].mean(), x[i].std()] for i in idx])
    for col, f in ((0, np.mean), (1, np.std)):
        _, d_gpu = boot.ci(x, f, method="pi", n_samples=n, seed=10, return_dist=True)
        assert sps.ks_2samp(d_gpu, ref[:, col]).pvalue > 1e-4
        assert abs(d_gpu.std() / ref[:, col].std() - 1) < 0.08


def test_bca_coverage_matches_the_pcg64_path():
    """Coverage of nominal 90 % BCa intervals for the mean of skewed data (N = 30): the device
    (Philox) and the oracle (PCG64) must cover the truth equally often on the same data sets,
    and their endpoints must agree to bootstrap noise."""
    import warnings
    rng = np.random.default_rng(44)
    trials, n_obs, n = 250, 30, 1000
    hit_gpu = hit_ref = 0
    diffs = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(trials):
            x = rng.exponential(1.0, n_obs)
            lo, hi = boot.ci(x, np.mean, alpha=0.1, n_samples=n, seed=k)
            rlo, rhi = orc.ci(x, np.mean, alpha=0.1, n_samples=n, seed=10_000 + k)
            hit_gpu += lo <= 1.0 <= hi
            hit_ref += rlo <= 1.0 <= rhi
            diffs.append([(lo - rlo) / x.std(), (hi - rhi) / x.std()])
    assert abs(hit_gpu - hit_ref) <= 12                       # same data: differences are rare
    assert 0.78 <= hit_gpu / trials <= 0.95
    d = np.array(diffs)
    assert np.abs(d.mean(axis=0)).max() < 0.01 and d.std(axis=0).max() < 0.05


def test_coverage_of_the_tiled_path():
    """Nominal 90 % percentile intervals for the mean cover the truth ~90 % of the
    time when N spans several cells (tile-multinomial draws)."""
    rng = np.random.default_rng(12)
    hits, trials, n = 0, 120, 40000
    for k in range(trials):
        x = rng.standard_normal(n)
        lo, hi = boot.ci(x, np.mean, alpha=0.1, method="pi", n_samples=400, seed=k)
        hits += lo <= 0.0 <= hi
    assert 0.80 <= hits / trials <= 0.97


# ------------------------------------------------------------------ full size ----

def test_cfg2_shape_parity():
    """BASELINE cfg2: N = 1e5, np.average, percentile interval — the statistic over the device's
    own materialised indices (oracle) at a reduced resample count, and the full call's shape."""
    x = np.random.default_rng(1).standard_normal(100_000)
    fused_vs_oracle((x,), "average", 200, seed=3)
    out = boot.ci(x, np.average, method="pi", n_samples=100_000, seed=0)
    se = x.std() / np.sqrt(x.size)
    assert_allclose(out, x.mean() + np.array([-1.96, 1.96]) * se, atol=0.05 * se)


def test_cfg3_shape_parity():
    """BASELINE cfg3: two paired arrays of N = 1e6, ratio of means, BCa: the oracle on the
    device's K2 indices (50 resamples), the O(N) jackknife against the closed form, and the
    size-independent scale identity ratio(c a, b) = c ratio(a, b) under one seed."""
    rng = np.random.default_rng
    a, b = rng(2).gamma(2, 1, 10**6), rng(3).gamma(3, 1, 10**6)
    fused_vs_oracle((a, b), "ratio_of_means", 50, seed=4, multi="paired")
    m = 2000
    _, d1 = boot.ci((a, b), boot.ratio_of_means, n_samples=m, seed=9, multi="paired", method="pi", return_dist=True)
    _, d2 = boot.ci((4.0 * a, b), boot.ratio_of_means, n_samples=m, seed=9, multi="paired", method="pi",
                    return_dist=True)
    assert_allclose(d2, 4.0 * d1, rtol=1e-13)
    j = _device.jackknife([dev(a), dev(b)], _abi.STAT_RATIO_OF_MEANS).cpu().numpy()
    js = orc.jack_ratio_of_means(a, b)
    assert_allclose(j[0], a.mean() / b.mean(), rtol=RTOL)
    assert_allclose(j[1], js.mean(), rtol=RTOL)
    # The acceleration sum d^3 / (6 (sum d^2)^1.5) is built from differences d = jmean - j_i of
    # leave-one-out values that agree to 6 digits at N = 1e6: every d carries eps * |j| / |d| ~ 1e-10
    # of rounding noise ON BOTH SIDES (numpy's as much as the device's), and the cubes cancel;
    # 1e-7 bounds the noise, and a change of 1e-7 * a = 6e-12 in a moves no order-statistic index.
    assert_allclose(j[4], orc.acceleration(js), rtol=1e-7)
    lo, hi = boot.ci((a, b), boot.ratio_of_means, n_samples=100_000, seed=0, multi="paired")
    assert lo < a.mean() / b.mean() < hi and (hi - lo) < 0.01


def test_cfg4_shape_parity():
    """BASELINE cfg4: 1000 x 10 000 table, median of every column, BCa, plus pval.  All columns
    share the index set, so a column of the full-width run equals the 1-D run under the same
    seed; a column subset is checked against the oracle on the device's own indices."""
    table = np.random.default_rng(4).standard_normal((1000, 10_000))
    td = dev(table)
    n = 400
    full = boot.ci(td, boot.median_axis0, n_samples=n, seed=5)
    pfull = boot.pval(td, boot.median_axis0, n_samples=n, seed=5)
    assert full.shape == (2, 10_000) and pfull.shape == (10_000,)
    idx = np.array(list(boot.bootstrap_indices(table, n, seed=5)))
    for c in (0, 1, 4999, 9999):
        col = table[:, c]
        want = orc.ci_from_indices((col,), np.median, idx, np.array([0.025, 0.975]), n, "bca",
                                   jstat=orc.jack_median(col))
        assert_allclose(full[:, c], want, rtol=RTOL, atol=1e-15)
        assert_array_equal(full[:, c], boot.ci(col, np.median, n_samples=n, seed=5))
        assert pfull[c] == np.mean(np.median(col[idx], axis=1) > 0)
    # the config's resample count: shapes, ordering, and the known sampling error of a median
    out = boot.ci(td, boot.median_axis0, n_samples=10_000, seed=0)
    assert (out[0] < out[1]).all()
    width = out[1] - out[0]
    assert abs(np.median(width) / (2 * 1.96 * 1.2533 / np.sqrt(1000)) - 1) < 0.1



def test_full_size_properties_cfg5():
    """BASELINE cfg5 shape (N=1e7) at a reduced resample count: size-independent
    properties — affine equivariance of the mean under the same seed, the bootstrap
    mean/variance identities, and shard independence."""
    n, m = 10**7, 1024
    x = np.random.default_rng(5).standard_normal(n) + 1.0
    xd = dev(x)
    s1 = _device.resample_stat([xd], _abi.STAT_MEAN, 7, 0, m).cpu().numpy()
    s2 = _device.resample_stat([xd * 2.0 + 3.0], _abi.STAT_MEAN, 7, 0, m).cpu().numpy()
    assert_allclose(s2, 2.0 * s1 + 3.0, rtol=1e-13)
    se = x.std() / np.sqrt(n)
    assert abs(s1.mean() - x.mean()) < 5 * se / np.sqrt(m)
    assert abs(s1.std() / se - 1) < 0.12
    part = _device.resample_stat([xd], _abi.STAT_MEAN, 7, 300, 420).cpu().numpy()
    assert_array_equal(part, s1[300:420])
    # materialised indices of one resample reproduce its statistic
    idx = _device.fill_indices(7, n, 1, 5, 6).cpu().numpy()[0]
    assert idx.min() >= 0 and idx.max() < n
    assert_allclose(x[idx].mean(), s1[5], rtol=1e-12)
    lo, hi = boot.ci(xd, np.mean, n_samples=m, seed=7)
    assert lo < x.mean() < hi and (hi - lo) < 6 * se


def test_very_long_array_two_level_tile_chains():
    """N = 2e8 (12 208 tiles, two-level occupancy chains, 1.6 GB of data created on the
    device): the interval must bracket the mean with the width the standard error implies,
    and the rank-space median path must run at this length too."""
    import torch
    n, m = 200_000_000, 500
    x = torch.randn(n, dtype=torch.float64, device="cuda") + 1.0
    mu = float(x.mean())
    se = float(x.std()) / np.sqrt(n)
    lo, hi = boot.ci(x, np.mean, n_samples=m, seed=1)
    assert lo < mu < hi and 3.0 < (hi - lo) / se < 5.0
    again = boot.ci(x, np.mean, n_samples=m, seed=1)
    assert_array_equal(again, [lo, hi])
    mlo, mhi = boot.ci(x, np.median, n_samples=m, seed=1)
    assert mlo < 1.0 + 5 * se and mhi > 1.0 - 5 * se and 3.0 < (mhi - mlo) / (1.2533 * se) < 5.0


def test_pinned_host_tensors_copy_overlaps_but_results_do_not_change():
    """Pinned host tensors are copied on a side stream while the occupancy chains (which do not
    read the data) are already running; every statistic must see the complete data."""
    import torch
    rng = np.random.default_rng(8)
    for n in (1000, 300_000, 3_000_000):
        a = rng.gamma(2.0, 1.0, n)
        b = 0.5 * a + rng.gamma(3.0, 1.0, n)
        pa, pb = torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()
        for rep in range(3):                      # repeated: the destination block is recycled
            assert_array_equal(boot.ci(pa, np.mean, n_samples=300, seed=rep), boot.ci(a, np.mean, n_samples=300, seed=rep))
        assert_array_equal(boot.ci(pa, np.std, n_samples=300, seed=2), boot.ci(a, np.std, n_samples=300, seed=2))
        assert_array_equal(boot.ci((pa, pb), boot.pearson_r, n_samples=300, seed=3),
                           boot.ci((a, b), boot.pearson_r, n_samples=300, seed=3))
        assert_array_equal(boot.ci((pa, pb), boot.linear_fit, n_samples=300, seed=3),
                           boot.ci((a, b), boot.linear_fit, n_samples=300, seed=3))
        assert_array_equal(boot.ci(pa, np.median, n_samples=300, seed=4), boot.ci(a, np.median, n_samples=300, seed=4))
        assert boot.pval(pa, np.mean, boot.greater_than(2.0), n_samples=300, seed=5) == \
            boot.pval(a, np.mean, boot.greater_than(2.0), n_samples=300, seed=5)


@pytest.mark.parametrize("n,d", [(50, 32), (1000, 100), (7169, 40), (28672, 33)])
def test_count_matrix_gemm_columns(n, d):
    """K6: the column sums of every resample as (count matrix) x (data).  The count matrix holds the
    multiplicities of exactly the indices bootstrap_indices yields; statistics agree with numpy on
    those indices and with the gather kernels (different summation order: 1e-12)."""
    rng = np.random.default_rng(n + d)
    data = rng.standard_normal((n, d)) + 3.0
    m = 64
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=21)))
    import torch
    c = torch.empty((m, n), dtype=torch.float64, device="cuda")
    _abi.check(_abi.lib().b200boot_count_matrix(21, n, 0, m, c.data_ptr(), _device.stream_ptr()))
    want_counts = np.stack([np.bincount(i, minlength=n) for i in idx]).astype(float)
    assert_array_equal(c.cpu().numpy(), want_counts)
    part = torch.empty((10, n), dtype=torch.float64, device="cuda")
    _abi.check(_abi.lib().b200boot_count_matrix(21, n, 30, 40, part.data_ptr(), _device.stream_ptr()))
    assert_array_equal(part.cpu().numpy(), want_counts[30:40])
    xd = dev(data)
    for stat_id, f in ((_abi.STAT_MEAN, np.mean), (_abi.STAT_SUM, np.sum), (_abi.STAT_VAR, np.var), (_abi.STAT_STD, np.std)):
        got = _device.resample_stat_gemm(xd, stat_id, 21, 0, m).cpu().numpy()          # [D][m]
        want = np.stack([f(data[i], axis=0) for i in idx], axis=1)
        assert_allclose(got, want, rtol=RTOL, atol=1e-13)
        if _device.columns_block_width(stat_id, n) > 0:
            gather = _device.resample_stat_columns(xd, stat_id, 21, 0, m).cpu().numpy()
            assert_allclose(got, gather, rtol=RTOL, atol=1e-13)
    L = _abi.lib()
    assert L.b200boot_gemm_supported(_abi.STAT_MEAN, 28672) == 1 and L.b200boot_gemm_supported(_abi.STAT_MEAN, 28673) == 0
    assert _device.gemm_supported(_abi.STAT_MEAN, 10**6) and not _device.gemm_supported(_abi.STAT_MEDIAN, 100)
    # the public call takes this path for D >= 32 and equals the per-column 1-D runs
    out = boot.ci(data, boot.mean_axis0, n_samples=200, seed=5)
    one = boot.ci(data[:, 3], np.mean, n_samples=200, seed=5)
    assert_allclose(out[:, 3], one, rtol=1e-12, atol=1e-15)


def test_count_matrix_gemm_long_columns():
    """Beyond one cell the count matrix comes from the materialised indices of the tiled plan."""
    rng = np.random.default_rng(3)
    n, d, m = 40001, 35, 40
    data = rng.standard_normal((n, d)) + 1.0
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=8)))
    xd = dev(data)
    for stat_id, f in ((_abi.STAT_MEAN, np.mean), (_abi.STAT_STD, np.std)):
        got = _device.resample_stat_gemm(xd, stat_id, 8, 0, m).cpu().numpy()
        want = np.stack([f(data[i], axis=0) for i in idx], axis=1)
        assert_allclose(got, want, rtol=RTOL, atol=1e-13)
    import os
    os.environ["B200BOOT_GEMM_COUNT_BYTES"] = str(8 * n * 7)          # chunks of 7 resamples
    try:
        chunked = _device.resample_stat_gemm(xd, _abi.STAT_MEAN, 8, 0, m).cpu().numpy()
    finally:
        del os.environ["B200BOOT_GEMM_COUNT_BYTES"]
    assert_allclose(chunked, np.stack([data[i].mean(axis=0) for i in idx], axis=1), rtol=RTOL, atol=1e-13)
    out = boot.ci(data, boot.mean_axis0, n_samples=100, seed=5, method="pi")
    one = boot.ci(data[:, 7], np.mean, n_samples=100, seed=5, method="pi")
    assert_allclose(out[:, 7], one, rtol=1e-12, atol=1e-15)


def test_wide_table_with_nan_keeps_gather_semantics():
    """0 * nan is nan in a matrix product, so the GEMM form would poison every resample; data with
    non-finite entries take the gather kernels, where only resamples that draw the row see it."""
    rng = np.random.default_rng(5)
    data = rng.standard_normal((200, 40))
    data[17, 3] = np.nan
    data[60, 9] = np.inf
    m = 300
    idx = np.array(list(boot.bootstrap_indices(data, m, seed=2)))
    with np.errstate(invalid="ignore"):
        want = np.sort(np.stack([data[i].mean(axis=0) for i in idx]), axis=0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, dist = boot.ci(data, boot.mean_axis0, n_samples=m, seed=2, method="pi", return_dist=True)
    assert np.isfinite(dist[:, 3]).sum() == np.isfinite(want[:, 3]).sum() > 0
    assert_allclose(dist, want, rtol=1e-12, atol=1e-15, equal_nan=True)
