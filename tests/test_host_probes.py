"""CPU checks of the pieces that do not need a GPU: the C ABI surface, and the
generator / sampler code the kernels share with the host probes
(b200boot_host_*, compiled from the same __host__ __device__ sources)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from scipy import stats as sps

import bootstrap_oracle as orc
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "b200boot.h")).read()
    declared = set(re.findall(r"\b(b200boot_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200boot_status", "b200boot_stat", "b200boot_cmp", "b200boot_plan"}
    assert len(declared) >= 20
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    nm = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (b200boot_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported
    assert _abi.lib().b200boot_version() == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built library every entry point raises."""
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "libb200boot.so"))
    with pytest.raises(_abi.B200BootError, match="no CPU fallback"):
        _abi.lib()
    with pytest.raises(_abi.B200BootError):
        _abi.make_plan(1000, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scikits-bootstrap_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(import|from)\s+(bootstrap_oracle|oracle|scikits\.bootstrap)\b", src, re.M), name
            assert not re.search(r"sys\.path.*(oracle|reference)", src), name


def test_errors_do_not_cross_abi_as_exceptions():
    L = _abi.lib()
    assert L.b200boot_make_plan(0, 1, C.byref(_abi.Plan())) == -1
    assert b"make_plan" in L.b200boot_last_error()
    with pytest.raises(_abi.B200BootError):
        _abi.check(L.b200boot_make_plan(-5, 1, C.byref(_abi.Plan())))


def test_abi_argument_validation_returns_status_codes():
    """Every entry point validates before it touches CUDA, so the status codes of
    include/b200boot.h can be checked without a GPU: pointers below are fakes that
    are never dereferenced on these paths."""
    L = _abi.lib()
    EINVAL, EWORKSPACE, EUNSUPPORTED = -1, -3, -4
    fake = 1 << 20                                    # 256-byte aligned, never dereferenced
    ptrs = (C.c_void_p * 2)(fake, fake)
    rs = L.b200boot_resample_stat
    assert rs(ptrs, 1, 1000, 99, 0, 0, 10, fake, fake, 1 << 30, None) == EUNSUPPORTED
    assert b"registry" in L.b200boot_last_error()
    assert rs(ptrs, 2, 1000, _abi.STAT_MEAN, 0, 0, 10, fake, fake, 1 << 30, None) == EINVAL      # takes 1 array
    assert rs(ptrs, 1, 1000, _abi.STAT_PEARSON_R, 0, 0, 10, fake, fake, 1 << 30, None) == EINVAL  # takes 2
    assert rs(ptrs, 1, 0, _abi.STAT_MEAN, 0, 0, 10, fake, fake, 1 << 30, None) == EINVAL
    assert rs(ptrs, 1, 1 << 31, _abi.STAT_MEAN, 0, 0, 10, fake, fake, 1 << 30, None) == EINVAL
    assert rs(ptrs, 1, 1000, _abi.STAT_MEAN, 0, 10, 5, fake, fake, 1 << 30, None) == EINVAL       # hi < lo
    assert rs(ptrs, 1, 1000, _abi.STAT_MEAN, 0, 0, 0, fake, fake, 1 << 30, None) == 0             # empty shard
    assert rs(ptrs, 1, 1000, _abi.STAT_MEAN, 0, 0, 10, None, fake, 1 << 30, None) == EINVAL       # null output
    assert rs(ptrs, 1, 1000, _abi.STAT_MEAN, 0, 0, 10, fake, fake + 8, 1 << 30, None) == EWORKSPACE
    assert b"aligned" in L.b200boot_last_error()
    assert rs(ptrs, 1, 1000, _abi.STAT_MEAN, 0, 0, 10, fake, fake, 64, None) == EWORKSPACE
    assert b"too small" in L.b200boot_last_error()
    need = L.b200boot_workspace_bytes(_abi.STAT_MEAN, 10**7, 1, 1, 1000, 2)
    assert need > 611 * 1000 * (4 + 16 * 2)           # tile (int32) + class (uint16 x 16) occupancies at least
    assert L.b200boot_workspace_bytes(_abi.STAT_MEAN, 0, 1, 1, 10, 2) == 0

    # rank rule of the median / quantile entry points
    rule = _abi.RankRule(k_lo=5, k_hi=7, gamma=0.5, loo_k_lo=4, loo_k_hi=5, loo_gamma=0.5, mode=1)
    rp = C.cast(C.pointer(rule), C.c_void_p)
    rm = L.b200boot_resample_median_columns
    assert rm(fake, 100, 1, 1, 0, 0, 10, fake, fake, 1 << 30, None, rp, None) == EINVAL           # k_hi > k_lo + 1
    rule.k_hi = 100
    rule.k_lo = 99
    assert rm(fake, 100, 1, 1, 0, 0, 10, fake, fake, 1 << 30, None, rp, None) == EINVAL           # rank outside the sample
    rule.k_lo, rule.k_hi, rule.gamma = 5, 6, 1.5
    assert rm(fake, 100, 1, 1, 0, 0, 10, fake, fake, 1 << 30, None, rp, None) == EINVAL
    assert rm(fake, 40000, 1, 1, 0, 0, 10, fake, fake, 1 << 30, None, None, None) == EUNSUPPORTED # beyond resident rows
    assert L.b200boot_resample_median_sorted(fake, 1000, 0, 0, 10, fake, fake, 1 << 30, None, None) == EINVAL  # one tile

    assert L.b200boot_count(fake, 10, 1, 9, fake, None, fake, None) == EINVAL                     # bad operator
    assert L.b200boot_count(fake, 10, 1, _abi.CMP_BETWEEN, fake, None, fake, None) == EINVAL      # needs t1
    assert L.b200boot_fill_moving_block(0, 5, 5, 0, 0, 3, fake, None) == EINVAL                   # no admissible start
    assert L.b200boot_fill_subsample(0, 10, 11, 0, 3, fake, fake, 1 << 30, None) == EINVAL        # size > N
    assert L.b200boot_fill_indices(0, 0, 1, 0, 3, fake, fake, 1 << 30, None) == EINVAL
    assert L.b200boot_jackknife_columns(fake, 10, 3, 2, _abi.STAT_MEAN, fake, None) == EINVAL     # ld < n_cols
    assert L.b200boot_jackknife_columns(fake, 10, 3, 3, _abi.STAT_PEARSON_R, fake, None) == EUNSUPPORTED
    assert L.b200boot_count_matrix(0, 40000, 0, 3, fake, None) == EUNSUPPORTED                   # beyond one cell
    assert L.b200boot_count_matrix(0, 100, 5, 3, fake, None) == EINVAL
    assert L.b200boot_finish_columns(_abi.STAT_VAR, 100, 3, 10, fake, None, None) == EINVAL       # needs s2
    assert L.b200boot_finish_columns(_abi.STAT_PEARSON_R, 100, 3, 10, fake, fake, None) == EUNSUPPORTED
    assert L.b200boot_gemm_supported(_abi.STAT_STD, 28672) == 1 and L.b200boot_gemm_supported(_abi.STAT_STD, 28673) == 0
    assert L.b200boot_columns_block_width(_abi.STAT_MEAN, 100000) == 0                            # beyond K1c rows
    assert L.b200boot_columns_block_width(_abi.STAT_MEAN, 1000) in (4, 8, 16, 32)


def _philox(c, k):
    out = (C.c_uint32 * 4)()
    _abi.lib().b200boot_host_philox4x32((C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), out)
    return list(out)


def test_philox_kat_through_library():
    assert _philox([0] * 4, [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    rng = np.random.default_rng(1)
    for _ in range(20):
        c = [int(v) for v in rng.integers(0, 2**32, 4)]
        k = [int(v) for v in rng.integers(0, 2**32, 2)]
        assert _philox(c, k) == [int(v) for v in orc.philox4x32_10(np.array([c], dtype=np.uint64), k)[0]]
        assert _philox_r(7, c, k) == [int(v) for v in orc.philox4x32(np.array([c], dtype=np.uint64), k, 7)[0]]


def _philox_r(rounds, c, k):
    out = (C.c_uint32 * 4)()
    _abi.lib().b200boot_host_philox4x32_r(rounds, (C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), out)
    return list(out)


def test_philox7_kat_through_library():
    """Random123 kat_vectors, `philox4x32 7` lines — the round count of the library's streams."""
    assert _abi.lib().b200boot_philox_rounds() in (7, 10)
    assert orc.PHILOX_ROUNDS == _abi.lib().b200boot_philox_rounds()
    assert _philox_r(7, [0] * 4, [0, 0]) == [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]
    assert _philox_r(7, [0xffffffff] * 4, [0xffffffff] * 2) == [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]
    assert _philox_r(7, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]
    assert _philox_r(10, [0] * 4, [0, 0]) == _philox([0] * 4, [0, 0])


def _host_indices(seed, n, n_arrays, r):
    out = np.empty(n, dtype=np.int64)
    _abi.check(_abi.lib().b200boot_host_indices(seed, n, n_arrays, r, out.ctypes.data))
    return out


def _host_counts(seed, n, n_arrays, r):
    p = _abi.make_plan(n, n_arrays)
    out = np.empty(p.n_tiles, dtype=np.int32)
    _abi.check(_abi.lib().b200boot_host_tile_counts(seed, n, n_arrays, r, out.ctypes.data))
    return out


def _host_class_counts(seed, n, n_arrays, r, counts):
    p = _abi.make_plan(n, n_arrays)
    out = np.empty((p.n_full, 16), dtype=np.int32)
    for t in range(p.n_full):
        _abi.check(_abi.lib().b200boot_host_class_counts(seed, r, t, int(counts[t]), out[t].ctypes.data))
    return out


@pytest.mark.parametrize("n", [1, 2, 5, 20, 1000, 28672])
def test_single_cell_indices_match_oracle(n):
    for seed, r in [(0, 0), (1234567890, 7), (2**63 + 5, 2**33 + 1)]:
        got = _host_indices(seed, n, 1, r)
        want = orc.device_indices(seed, r, n, 14)
        np.testing.assert_array_equal(got, want)
        assert got.min() >= 0 and got.max() < n


def test_lemire_redraw_path_matches_oracle():
    # size 3*2^30 rejects a word with probability 1/4: exercises the redraw stream
    n = 3 << 30
    hit = 0
    for r in range(3):
        want = orc.cell_draws_lemire(11, r, 0, 64, n)
        # reproduce through the library one block at a time via philox probe
        words = [w for q in range(16) for w in _philox_r(orc.PHILOX_ROUNDS, [q, 0, r, 0], [11, 0])]
        thresh = (1 << 32) % n
        for p, w in enumerate(words):
            if (w * n) & 0xFFFFFFFF < thresh:
                hit += 1
            else:
                assert want[p] == (w * n) >> 32
        assert want.min() >= 0 and want.max() < n
    assert hit > 20


@pytest.mark.parametrize("n,n_arrays", [(28673, 1), (40000, 1), (16384 * 3, 1), (20000, 2), (100000, 1)])
def test_tiled_indices_match_oracle_given_counts(n, n_arrays):
    plan = _abi.make_plan(n, n_arrays)
    assert plan.n_tiles > 1
    for seed, r in [(3, 0), (99, 12345)]:
        counts = _host_counts(seed, n, n_arrays, r)
        assert counts.sum() == n and (counts >= 0).all()
        cls = _host_class_counts(seed, n, n_arrays, r, counts)
        assert (cls.sum(axis=1) == counts[:plan.n_full]).all() and (cls >= 0).all()
        got = _host_indices(seed, n, n_arrays, r)
        want = orc.device_indices(seed, r, n, plan.log2_tile, counts=list(counts), n_arrays=n_arrays,
                                  class_counts=cls)
        np.testing.assert_array_equal(got, want)


def test_plan_shapes():
    p = _abi.make_plan(10**7, 1)
    assert (p.log2_tile, p.n_full, p.rem, p.n_tiles) == (14, 610, 5760, 611)
    p = _abi.make_plan(10**6, 2)
    assert (p.log2_tile, p.n_full, p.rem, p.n_tiles) == (13, 122, 576, 123)
    p = _abi.make_plan(1000, 1)
    assert (p.n_full, p.rem, p.n_tiles) == (0, 1000, 1)
    assert orc.tile_plan(10**7, 14) == [16384] * 610 + [5760]


@pytest.mark.parametrize("n,p", [(10**7, 16384 / 10**7), (10**7, 0.5), (50000, 0.3), (1000, 0.003),
                                 (40, 0.2), (100, 0.97), (17, 0.5), (10**6, 1 - 1e-5)])
def test_binomial_sampler_distribution(n, p):
    L = _abi.lib()
    m = 20000
    xs = np.array([L.b200boot_host_binomial(5, r, 3, n, p) for r in range(m)])
    assert xs.min() >= 0 and xs.max() <= n
    mean, var = n * p, n * p * (1 - p)
    assert abs(xs.mean() - mean) < 5 * np.sqrt(var / m) + 1e-9
    assert abs(xs.var() - var) < 6 * var * np.sqrt(2.0 / m) + 1e-9
    # chi-square against the exact pmf on ~20 equiprobable bins
    qs = np.unique(sps.binom.ppf(np.linspace(0, 1, 21)[1:-1], n, p))
    edges = np.concatenate(([-0.5], qs + 0.5, [n + 0.5]))
    obs = np.histogram(xs, bins=edges)[0]
    cdf = sps.binom.cdf(np.floor(edges[1:-1]), n, p)
    exp = np.diff(np.concatenate(([0.0], cdf, [1.0]))) * m
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert chi2 < sps.chi2.ppf(1 - 1e-6, max(keep.sum() - 1, 1))


@pytest.mark.parametrize("n", [1, 2, 7, 19, 20, 21, 100, 1024, 16384, 40000])
def test_even_split_sampler_distribution(n):
    """Bin(n, 1/2) of the class split tree: set bits of n random bits below 20, BTRS above."""
    L = _abi.lib()
    m = 20000
    xs = np.array([L.b200boot_host_binomial_half(11, r, 5, n) for r in range(m)])
    assert xs.min() >= 0 and xs.max() <= n
    assert L.b200boot_host_binomial_half(11, 0, 5, 0) == 0
    mean, var = n / 2, n / 4
    assert abs(xs.mean() - mean) < 5 * np.sqrt(var / m)
    assert abs(xs.var() - var) < 6 * var * np.sqrt(2.0 / m)
    qs = np.unique(sps.binom.ppf(np.linspace(0, 1, 21)[1:-1], n, 0.5))
    edges = np.concatenate(([-0.5], qs + 0.5, [n + 0.5]))
    obs = np.histogram(xs, bins=edges)[0]
    cdf = sps.binom.cdf(np.floor(edges[1:-1]), n, 0.5)
    exp = np.diff(np.concatenate(([0.0], cdf, [1.0]))) * m
    keep = exp > 5
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert chi2 < sps.chi2.ppf(1 - 1e-6, max(keep.sum() - 1, 1))
    # symmetric: k and n - k equally likely
    assert abs((xs > n / 2).mean() - (xs < n / 2).mean()) < 5 / np.sqrt(m)


def test_tile_counts_are_multinomial():
    n, m = 100000, 4000           # 6 full tiles of 16384 + remainder 1696
    plan = _abi.make_plan(n, 1)
    sizes = np.array([16384] * 6 + [1696])
    assert plan.n_tiles == 7
    cnt = np.array([_host_counts(21, n, 1, r) for r in range(m)], dtype=float)
    assert (cnt.sum(axis=1) == n).all()
    pr = sizes / n
    z = (cnt.mean(axis=0) - n * pr) / np.sqrt(n * pr * (1 - pr) / m)
    assert np.abs(z).max() < 5
    # covariance of two cells: -n p_i p_j
    cov = np.cov(cnt[:, 0], cnt[:, 1])[0, 1]
    assert abs(cov - (-n * pr[0] * pr[1])) < 6 * n * pr[0] * (1 - pr[0]) / np.sqrt(m)


def test_tile_counts_two_level_chain():
    """More than 32 cells: occupancies are sampled per group of 32 cells, then inside each
    group; marginals, within-group and across-group covariances must stay multinomial."""
    n = 16384 * 70 + 999                 # 70 full tiles + remainder = 71 cells = 3 groups
    m = 1500
    plan = _abi.make_plan(n, 1)
    assert plan.n_tiles == 71
    cnt = np.array([_host_counts(8, n, 1, r) for r in range(m)], dtype=float)
    assert (cnt.sum(axis=1) == n).all()
    sizes = np.array([16384] * 70 + [999], dtype=float)
    pr = sizes / n
    z = (cnt.mean(axis=0) - n * pr) / np.sqrt(n * pr * (1 - pr) / m)
    assert np.abs(z).max() < 5 and abs(z.std() - 1) < 0.3
    v = cnt.var(axis=0) / (n * pr * (1 - pr))
    assert np.abs(v - 1).max() < 0.25
    for i, j in [(0, 1), (0, 40), (31, 32), (10, 69)]:          # same group / different groups
        cov = np.cov(cnt[:, i], cnt[:, j])[0, 1]
        assert abs(cov + n * pr[i] * pr[j]) < 6 * n * pr[i] / np.sqrt(m)
    g = cnt[:, :32].sum(axis=1)                                  # a whole group is binomial too
    pg = sizes[:32].sum() / n
    assert abs(g.mean() - n * pg) < 5 * np.sqrt(n * pg * (1 - pg) / m)
    assert 0.8 < g.var() / (n * pg * (1 - pg)) < 1.25


def test_class_counts_are_multinomial():
    m = 3000
    out = np.empty((m, 16), dtype=np.int32)
    for r in range(m):
        _abi.check(_abi.lib().b200boot_host_class_counts(4, r, 7, 16384, out[r].ctypes.data))
    assert (out.sum(axis=1) == 16384).all()
    z = (out.mean(axis=0) - 1024) / np.sqrt(1024 * (15 / 16) / m)
    assert np.abs(z).max() < 5
    v = out.var(axis=0) / (1024 * 15 / 16)
    assert np.abs(v - 1).max() < 0.2
    cov = np.cov(out[:, 0], out[:, 15])[0, 1]
    assert abs(cov + 64) < 6 * 960 / np.sqrt(m)


def test_tiled_draws_are_uniform_over_rows():
    n, m = 40000, 300             # 2 full tiles + remainder 7232
    hist = np.zeros(n)
    for r in range(m):
        hist += np.bincount(_host_indices(77, n, 1, r), minlength=n)
    chi2 = ((hist - m) ** 2 / m).sum()
    assert abs(chi2 - (n - 1)) < 6 * np.sqrt(2 * (n - 1))
    # cell boundaries carry no bias
    for lo, hi in [(0, 16384), (16384, 32768), (32768, n)]:
        share = hist[lo:hi].sum() / (n * m)
        assert abs(share - (hi - lo) / n) < 5 * np.sqrt((hi - lo) / n / (n * m))


def test_tiled_draws_have_iid_multiplicities():
    """Row multiplicities of a resample must look like N iid uniform draws:
    #rows drawn k times ~ N * Binomial(N, 1/N).pmf(k); in particular a fraction 1 - 1/e of
    the rows is hit.  Sensitive to any over- or under-dispersion the tile / bank-class
    multinomial splits could introduce."""
    n, m = 100000, 40           # 6 full tiles + remainder; 16 classes per full tile
    mult = np.zeros(8)
    distinct = []
    for r in range(m):
        c = np.bincount(_host_indices(123, n, 1, r), minlength=n)
        distinct.append((c > 0).mean())
        mult += np.bincount(np.minimum(c, 7), minlength=8)
    pmf = sps.binom.pmf(np.arange(7), n, 1.0 / n)
    exp = np.concatenate((pmf, [1 - pmf.sum()])) * n * m
    chi2 = ((mult - exp) ** 2 / exp).sum()
    assert chi2 < sps.chi2.ppf(1 - 1e-5, 7), (chi2, mult, exp)
    # Var(#distinct)/N = e^-1 (1 - 2 e^-1) ~ 0.0972 for iid draws
    d = np.array(distinct)
    assert abs(d.mean() - (1 - np.exp(-1))) < 5 * np.sqrt(0.0972 / n / m)
    assert 0.4 < d.var() * n / 0.0972 < 2.0
    # occupancy of an arbitrary row block (crossing a tile and several classes): Binomial(N, B/N)
    occ = np.array([np.count_nonzero((lambda i: (i >= 16000) & (i < 17000))(_host_indices(5, n, 1, r)))
                    for r in range(200)])
    assert abs(occ.mean() - 1000) < 5 * np.sqrt(1000 / 200)
    assert 0.7 < occ.var() / (1000 * (1 - 0.01)) < 1.4


def test_registry_and_errors_without_gpu():
    import functools
    from scikits_bootstrap_b200 import _registry as reg
    assert reg.lookup(np.mean).stat_id == _abi.STAT_MEAN
    assert reg.lookup(np.average).stat_id == _abi.STAT_MEAN
    assert reg.lookup(functools.partial(np.median, axis=0)).axis0
    assert reg.lookup(lambda x: x.mean()) is None
    x = np.arange(10.0)
    with pytest.raises(NotImplementedError, match="device statistic registry"):
        boot.ci(x, lambda a: np.mean(a))
    with pytest.raises(ValueError, match=r"Value `wrong` for multi is not recognized."):
        boot.ci(x, multi="wrong")
    with pytest.raises(ValueError, match=r"Method invalid is not supported"):
        boot.ci(x, method="invalid")
    with pytest.raises(ValueError, match=r"Output option invalid is not supported"):
        boot.ci(x, output="invalid")
    with pytest.raises(TypeError):
        boot.ci(x, "average")
    with pytest.raises(ValueError):
        boot.ci(([1, 2, 3], [1, 2]), boot.ratio_of_means, multi="paired")
    with pytest.raises(ValueError):
        boot.ci(x, method="abc", return_dist=True)
    with pytest.raises(NotImplementedError, match="not currently supported"):
        boot.ci((x, x), multi="independent", method="abc")
    with pytest.raises(TypeError, match="statfunction does not accept correct arguments"):
        boot.ci(x, np.mean, method="abc")             # bootstrap_test.py:270-274
    with pytest.raises(ValueError, match="Output option invalid is not"):
        boot.ci(x, output="invalid", method="abc")    # bootstrap_test.py:197-199
    with pytest.raises(NotImplementedError, match="Numba for independent data"):
        boot.ci((x, x), multi="independent", use_numba=True)
    with pytest.raises(ValueError, match="Numba can't be used"):
        boot.ci(x, np.median, use_numba=True)
    np.testing.assert_array_equal(np.array(list(boot.jackknife_indices(np.array([1, 2, 3])))),
                                  [[1, 2], [0, 2], [0, 1]])     # bootstrap_test.py:110-112


def test_interval_math_matches_oracle(golden):
    from scikits_bootstrap_b200 import _interval as iv
    np.testing.assert_array_equal(iv.nppf(golden["nppf/p"]), golden["nppf/val"])
    np.testing.assert_array_equal(iv.ncdf(golden["ncdf/x"]), golden["ncdf/val"])
    # the per-element path taken for a handful of levels equals the array path bit for bit,
    # on the reference's own grid and across the three branches and their edges
    grid = np.asarray(golden["nppf/p"], dtype=float).reshape(-1)
    rng = np.random.default_rng(1)
    grid = np.concatenate([grid, rng.random(20000), rng.random(5000) * 0.03, 1 - rng.random(5000) * 0.03,
                           [0.0, 1.0, 0.02425, 0.97575, 0.5, 1e-300, 1 - 1e-16, np.nan]])
    with np.errstate(all="ignore"):
        want = iv._nppf_array(grid)
    got = np.array([iv._nppf_scalar(float(v)) for v in grid])
    np.testing.assert_array_equal(got, want)
    for k in range(0, 64, 7):
        np.testing.assert_array_equal(iv.nppf(grid[k:k + 5]), want[k:k + 5])
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(50, 5000))
        av = rng.random((3, 4))
        a, fa = iv.order_indices(av, n)
        b, fb = orc.order_indices(av, n)
        np.testing.assert_array_equal(a, b)
        assert fa == fb
    # few levels take a per-element path: same integers and flags as the array form on exact
    # grid points, half-way cases (round half to even), NaN, the ends and out-of-range levels
    for t in range(3000):
        n = int(rng.integers(2, 3000))
        av = rng.random(int(rng.integers(1, 9)))
        mode = t % 6
        if mode == 1:
            av[rng.integers(av.size)] = np.nan
        elif mode == 2:
            av = np.round(av * (n - 1)) / (n - 1)
        elif mode == 3:
            av = (np.floor(av * (n - 1)) + 0.5) / (n - 1)
        elif mode == 4:
            av[rng.integers(av.size)] = rng.choice([0.0, 1.0, 1e-9, 1 - 1e-9])
        elif mode == 5:
            av = av * 3 - 1
        a, fa = iv.order_indices(av, n)
        with np.errstate(all="ignore"):
            b, fb = iv._order_indices_array(av, n)
        np.testing.assert_array_equal(a, b)
        assert fa == fb and a.dtype == b.dtype


def test_rank_rule_reproduces_numpy_quantile_bitwise():
    """The host fixes which two order statistics np.quantile(x, q) reads and the weight between
    them (_device.rank_rule); the kernels only fetch and combine.  Applying the rule on the host
    with numpy's _lerp formula must give np.quantile itself, bit for bit, for samples of n and of
    n - 1 values (the leave-one-out rule)."""
    rule = lambda n, q: _device_mod().rank_rule(n, q)
    rng = np.random.default_rng(0)

    def lerp(a, b, t):
        d = b - a
        return b - d * (1 - t) if t >= 0.5 else a + d * t

    qs = [0.0, 1.0, 0.5, 0.25, 0.1, 0.9, 0.975, 1 / 3, 2 / 3, 1e-9, 1 - 1e-9]
    for n in [1, 2, 3, 4, 5, 7, 10, 11, 100, 101, 1000, 4097, 28672, 100003]:
        x = np.sort(rng.standard_normal(n))
        for q in qs + list(rng.random(6)):
            r = rule(n, q)
            assert 0 <= r.k_lo <= r.k_hi <= min(r.k_lo + 1, n - 1) and r.mode == 1
            got = x[r.k_hi] if r.k_lo == r.k_hi else lerp(x[r.k_lo], x[r.k_hi], r.gamma)
            assert got == np.quantile(x, q), (n, q)
            if n > 1:
                y = x[:-1]
                got = y[r.loo_k_hi] if r.loo_k_lo == r.loo_k_hi else lerp(y[r.loo_k_lo], y[r.loo_k_hi], r.loo_gamma)
                assert got == np.quantile(y, q), (n, q, "loo")
    assert rule(10, None) is None                      # np.median: the library's own rule


def test_registry_recognises_quantiles_fits_and_differences():
    import functools
    reg = _registry_mod()
    spec = reg.lookup(functools.partial(np.percentile, q=90))
    assert spec.stat_id == _abi.STAT_MEDIAN and spec.q == 0.9 and not spec.axis0
    spec = reg.lookup(functools.partial(np.quantile, q=0.25, axis=0, method="linear"))
    assert spec.q == 0.25 and spec.axis0
    assert reg.lookup(functools.partial(np.quantile, q=0.5, method="nearest")) is None
    assert reg.lookup(functools.partial(np.quantile, q=0.5, axis=1)) is None
    assert reg.lookup(functools.partial(np.quantile, q=[0.1, 0.9])) is None
    with pytest.raises(ValueError, match="range"):
        reg.quantile(1.5)
    assert reg.percentile(50).spec.q == 0.5
    fit = reg.linear_fit
    assert fit.spec.n_out == 2 and fit.spec.n_arrays == 2 and fit.spec.stat_id == _abi.STAT_LINEAR_FIT
    a, b = np.arange(10.0), np.arange(10.0) * 2 + 1
    np.testing.assert_allclose(fit(a, b), [2.0, 1.0], atol=1e-12)
    d = reg.diff_of(np.median)
    assert d.spec.independent and d.spec.stat_id == _abi.STAT_MEDIAN and d(a, b) == np.median(a) - np.median(b)
    assert reg.diff_of(reg.quantile(0.8)).spec.q == 0.8
    with pytest.raises(NotImplementedError):
        reg.diff_of(reg.pearson_r)
    with pytest.raises(NotImplementedError):
        reg.require(lambda v: v.mean())
    # which form serves (N, D) moment statistics: the measured crossover
    dev = _device_mod()
    assert dev.gemm_pays(1000, 10000) and dev.gemm_pays(7000, 8) and dev.gemm_pays(100000, 64)
    assert dev.gemm_pays(1000000, 1024) and not dev.gemm_pays(1000000, 1023)
    assert not dev.gemm_pays(1000, 16) and not dev.gemm_pays(7000, 4) and not dev.gemm_pays(100000, 63)


def _device_mod():
    from scikits_bootstrap_b200 import _device
    return _device


def _registry_mod():
    from scikits_bootstrap_b200 import _registry
    return _registry


def test_bca_levels_scalar_path_is_bit_equal_to_the_array_path():
    """The per-call fast path of the BCa levels (Python floats) against the array form."""
    from scikits_bootstrap_b200 import _interval
    rng = np.random.default_rng(7)
    for _ in range(400):
        n = int(rng.integers(50, 200000))
        less = np.int64(rng.integers(1, n))
        accel = np.float64(rng.normal(0, 0.05))
        alphas = np.sort(rng.random(int(rng.integers(1, 6))))
        fast = _interval.bca_levels(less, n, accel, alphas)
        z0 = _interval.nppf((1.0 * np.asarray(less)) / n)
        with np.errstate(invalid="ignore", divide="ignore"):
            zs = z0 + _interval.nppf(alphas).reshape(alphas.shape + (1,) * z0.ndim)
            slow = _interval.ncdf(z0 + zs / (1 - accel * zs))
        assert fast.shape == slow.shape
        assert (fast.view(np.uint64) == slow.view(np.uint64)).all()
    # degenerate inputs take the array form (inf / nan arithmetic as numpy does it)
    out = _interval.bca_levels(np.int64(0), 100, np.float64(0.01), np.array([0.025, 0.975]))
    assert np.isnan(out).all() or ((out == 0) | (out == 1) | np.isnan(out)).all()
    assert np.isnan(_interval.bca_levels(np.int64(10), 100, np.float64("nan"), np.array([0.5]))).all()


def test_ncdf_array_path_equals_math_erf_bitwise():
    """Arrays of levels go through the library's host loop over the C library's erf; a handful
    of values through math.erf (the reference's np.vectorize(math.erf) path): same bits."""
    import math
    from scikits_bootstrap_b200 import _interval
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(50000) * 3, rng.standard_normal(20000) * 1e-3,
                        rng.standard_normal(5000) * 40, [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-320, -1e-320, 5.0, -5.0]])
    got = _interval.ncdf(x)
    want = np.array([0.5 * (1 + math.erf(v / math.sqrt(2))) for v in x.tolist()])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert _interval.ncdf(x.reshape(-1, 3)[:200]).shape == (200, 3)
    # and the BCa levels of many columns equal the per-column scalar path
    less = rng.integers(1, 999, 500)
    accel = rng.standard_normal(500) * 0.02
    alphas = np.array([0.025, 0.975])
    many = _interval.bca_levels(less, 1000, accel, alphas)
    for d in (0, 17, 499):
        one = _interval.bca_levels(less[d], 1000, accel[d], alphas)
        assert np.array_equal(many[:, d].view(np.uint64), one.view(np.uint64))
