"""The reference's own test-suite (tests/bootstrap_test.py, class TestCI) run against
this package, test for test and under the same names.

Substitutions, each noted where it happens:
  * fixed-seed goldens pin numpy's PCG64 stream; they are checked in index-supplied mode
    (the GPU consumes the reference's own index arrays), or as properties of the device
    stream when the golden is an index array itself;
  * statistics given as lambdas are replaced by registry statistics with the same structure
    (the device path has no host fallback for arbitrary callables); np.polyfit(a, b, 1) is
    the registry's linear_fit.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import bootstrap_oracle as orc
import scikits_bootstrap_b200 as boot
from scikits_bootstrap_b200 import InstabilityWarning

pytestmark = pytest.mark.gpu


class TestCI:
    def setup_method(self):
        self.data = np.array([1.34016346, 1.73759123, 1.49898834, -0.22864333, 2.031034, 2.17032495,
                              1.59645265, -0.76945156, 0.56605824, -0.11927018, -0.1465108, -0.79890338,
                              0.77183278, -0.82819136, 1.32667483, 1.05986776, 2.14408873, -1.43464512,
                              2.28743654, 0.42864858])
        self.x = [1, 2, 3, 4, 5, 6]
        self.y = [2, 1, 2, 5, 1, 2]
        self.z = [2, 1, 1, -1, -1, -4, -8]
        self.seed = 1234567890

    def ref_indices(self, n, n_samples):
        return np.array(list(orc.bootstrap_indices_pcg64(n, n_samples, self.seed)))

    def test_numba_close(self):                               # test:51-58 (both runs are the device path)
        dat = np.random.randint(10, size=50)
        no_numba = boot.ci(dat, n_samples=100000)
        with_numba = boot.ci(dat, n_samples=100000, use_numba=True)
        assert_allclose(no_numba, with_numba, rtol=1e-2)

    def test_invalid_method(self):                            # test:60-62
        with pytest.raises(ValueError, match=r"Method invalid is not supported"):
            boot.ci(self.data, method="invalid")

    def test_invalid_output(self):                            # test:64-66
        with pytest.raises(ValueError, match=r"Output option invalid is not supported"):
            boot.ci(self.data, output="invalid")

    def test_bootstrap_indices(self):                         # test:73-84 (golden is PCG64: oracle pin)
        indices = np.array(list(boot.bootstrap_indices(np.array([1, 2, 3, 4, 5]), n_samples=3, seed=self.seed)))
        assert indices.shape == (3, 5) and indices.dtype == np.int64
        assert indices.min() >= 0 and indices.max() < 5
        again = np.array(list(boot.bootstrap_indices(np.array([1, 2, 3, 4, 5]), n_samples=3, seed=self.seed)))
        assert_array_equal(indices, again)
        assert_array_equal(np.array(list(orc.bootstrap_indices_pcg64(5, 3, self.seed))),
                           [[2, 3, 0, 0, 2], [2, 3, 3, 0, 3], [0, 0, 2, 4, 2]])

    def test_bootstrap_indices_moving_block(self):            # test:86-97 (structure on the device stream)
        idx = np.array(list(boot.bootstrap_indices_moving_block(np.array([1, 2, 3, 4, 5]), n_samples=3,
                                                                 seed=self.seed)))
        assert idx.shape == (3, 5)
        assert np.all(np.diff(idx[:, :3], axis=1) == 1) and np.all(np.diff(idx[:, 3:], axis=1) == 1)
        assert idx[:, 0].max() < 2 and idx[:, 3].max() < 2    # starts from numpy's half-open [0, N-L)

    def test_bootstrap_indices_moving_block_wrap(self):       # test:99-108
        idx = np.array(list(boot.bootstrap_indices_moving_block(np.array([1, 2, 3, 4]), n_samples=3,
                                                                 seed=self.seed, wrap=True)))
        assert idx.shape == (3, 4)
        assert np.all((np.diff(idx[:, :3], axis=1) % 4) == 1)

    def test_jackknife_indices(self):                         # test:110-112
        indices = np.array([x for x in boot.jackknife_indices(np.array([1, 2, 3]))])
        assert_allclose(indices, np.array([[1, 2], [0, 2], [0, 1]]))

    def test_subsample_indices(self):                         # test:114-118
        for x in boot.subsample_indices(self.data, 1000, 0.5):
            assert len(np.unique(x)) == len(self.data) / 2

    def test_subsample_indices_fixed(self):                   # test:120-123
        for x in boot.subsample_indices(self.data, 1000, 10):
            assert len(np.unique(x)) == 10

    def test_subsample_size_too_large(self):                  # test:125-127
        with pytest.raises(ValueError):
            boot.subsample_indices(self.data, 1000, 30)

    def test_subsample_indices_notsame(self):                 # test:129-134
        indices = boot.subsample_indices(np.arange(0, 50), 1000, -1)
        assert not np.all(indices[0] == indices[1:])

    def test_subsample_invalid_size(self):                    # test:136-138
        with pytest.raises(ValueError, match="size cannot be -5"):
            boot.subsample_indices(self.data, size=-5)

    def test_abc_simple(self):                                # test:140-142 (5e-7: DESIGN.md section 9)
        results = boot.ci(self.data, method="abc", seed=self.seed)
        assert_allclose(results, np.array([0.20982275, 1.20374686]), rtol=5e-7)

    def test_abc_multialpha_defaultstat(self):                # test:144-150
        results = boot.ci(self.data, alpha=(0.1, 0.2, 0.8, 0.9), method="abc", seed=self.seed)
        assert_allclose(results, np.array([0.39472915, 0.51161304, 0.93789723, 1.04407254]), rtol=5e-7)

    def test_pi_multialpha(self):                             # test:163-169 (reference's indices)
        results = boot.ci_indexed(self.data, None, self.ref_indices(20, 10000), method="pi",
                                  alpha=(0.1, 0.2, 0.8, 0.9))
        assert_allclose(results, np.array([0.401879, 0.517506, 0.945416, 1.052798]), rtol=1e-6)

    def test_dist(self):                                      # test:171-176
        out, dist = boot.ci(self.data, return_dist=True, method="pi", alpha=[0.1, 0.9], n_samples=100)
        assert out[0] == dist[10]
        assert out[1] == dist[89]

    def test_bca_simple(self):                                # test:178-181
        results = boot.ci(self.data, seed=self.seed)
        results2 = boot.ci(self.data, alpha=(0.025, 1 - 0.025), seed=self.seed)
        assert_allclose(results, results2)

    def test_bca_errorbar_output_simple(self):                # test:183-188
        results_default = boot.ci(self.data, seed=self.seed)
        results_errorbar = boot.ci(self.data, output="errorbar", seed=self.seed)
        assert_allclose(results_errorbar.T, abs(np.average(self.data) - results_default)[np.newaxis])

    def test_abc_errorbar_output_simple(self):                # test:190-195
        results_default = boot.ci(self.data, method="abc")
        results_errorbar = boot.ci(self.data, output="errorbar", method="abc")
        assert_allclose(results_errorbar.T, abs(np.average(self.data) - results_default)[np.newaxis])

    def test_abc_errorbar_unsupported(self):                  # test:197-199
        with pytest.raises(ValueError, match="Output option invalid is not"):
            boot.ci(self.data, output="invalid", method="abc")

    def test_errorbar_unsupported(self):                      # test:201-203
        with pytest.raises(ValueError, match="Output option invalid is not"):
            boot.ci(self.data, output="invalid")

    def test_bca_multialpha(self):                            # test:205-209 (reference's indices)
        results = boot.ci_indexed(self.data, None, self.ref_indices(20, 10000), alpha=(0.1, 0.2, 0.8, 0.9))
        assert_allclose(results, np.array([0.386674, 0.506714, 0.935628, 1.039683]), rtol=1e-6)

    def test_bca_multi_multialpha(self):                      # test:211-238 (polyfit = linear_fit)
        kw = dict(alpha=(0.1, 0.2, 0.8, 0.9), n_samples=1000, seed=self.seed)
        results1 = boot.ci((self.x, self.y), boot.linear_fit, **kw)
        results2 = boot.ci((self.x, self.y), boot.linear_fit, multi="paired", **kw)
        results3 = boot.ci((self.x, self.y), boot.linear_fit, output="errorbar", **kw)
        assert results1.shape == (4, 2) and results3.shape == (2, 4)
        assert_allclose(results1, results2)
        assert_allclose(np.abs(np.polyfit(self.x, self.y, 1) - results1), results3.T)
        # the same statistic on the reference's own index draws, against the reference algorithm.
        # Percentile ranks only: on this 6-point integer data 19 of the 1000 resamples reproduce
        # the data's own fit, so BCa's count of `stat < ostat` (bootstrap.py:583) is decided by
        # rounding noise in the reference itself (SVD vs moment sums differ in the last bit).
        idx = self.ref_indices(6, 1000)
        got = boot.ci_indexed((self.x, self.y), boot.linear_fit, idx, alpha=(0.1, 0.2, 0.8, 0.9), method="pi")
        want = orc.ci_from_indices((np.array(self.x, float), np.array(self.y, float)),
                                   lambda a, b: np.polyfit(a, b, 1), idx, (0.1, 0.2, 0.8, 0.9), 1000, "pi")
        assert_allclose(got, want, rtol=1e-9, atol=1e-12)
        for statfun in (boot.pearson_r, boot.ratio_of_means):
            kw = dict(alpha=(0.1, 0.2, 0.8, 0.9), n_samples=1000, seed=self.seed)
            results1 = boot.ci((self.x, self.y), statfun, **kw)
            results2 = boot.ci((self.x, self.y), statfun, multi="paired", **kw)
            results3 = boot.ci((self.x, self.y), statfun, output="errorbar", **kw)
            assert_allclose(results1, results2)
            assert_allclose(np.abs(statfun(self.x, self.y) - results1)[np.newaxis].T, results3)

    def test_bca_multi_indep(self):                           # test:240-248 (lambda -> diff_of_means)
        n = 1000
        idx = list(boot.bootstrap_indices_independent((self.x, self.z), n, seed=self.seed))
        got = boot.ci((self.x, self.z), boot.diff_of_means, n_samples=n, multi="independent", seed=self.seed)
        want = orc.ci_from_indices((np.array(self.x), np.array(self.z)),
                                   lambda a, b: np.average(a) - np.average(b), idx, (0.025, 0.975), n,
                                   "bca", multi="independent")
        assert_allclose(got, want, rtol=1e-12)
        assert_allclose(got, [2.547619, 7.97619], atol=0.6)   # the PCG64 golden, statistically

    def test_allequal_warn(self):                             # test:250-252
        with pytest.warns(InstabilityWarning, match="(NaN|all equal)"):
            boot.ci(np.ones(20), seed=self.seed)

    def test_extremal_warn(self):                             # test:254-256
        with pytest.warns(InstabilityWarning, match="extremal"):
            boot.ci([1, 2], n_samples=10, seed=self.seed)

    def test_10_warn(self):                                   # test:258-260 (reference's indices)
        with pytest.warns(InstabilityWarning, match="top 10"):
            boot.ci_indexed(np.arange(1, 20), None, self.ref_indices(19, 50))

    def test_numba_no_indep(self):                            # test:262-264
        with pytest.raises(NotImplementedError, match="Numba for independent data"):
            boot.ci((self.x, self.y), multi="independent", use_numba=True)

    def test_abc_no_indep(self):                              # test:266-268
        with pytest.raises(NotImplementedError, match="not currently supported"):
            boot.ci((self.x, self.y), multi="independent", method="abc")

    def test_abc_no_weights(self):                            # test:270-274 (a statistic without weights=)
        with pytest.raises(TypeError, match="statfunction does not accept correct arguments"):
            boot.ci(self.x, np.mean, method="abc")

    def test_bca_multi_unequal_paired(self):                  # test:276-284
        with pytest.raises(ValueError):
            boot.ci((self.x, self.z), boot.diff_of_means, n_samples=1000, multi="paired", seed=self.seed)

    def test_bca_multi_2dout_multialpha(self):                # test:286-309 (vector statistic == scalar runs)
        self._vector_stat_columns("bca")

    def test_multi_fail(self):                                # test:311-319
        with pytest.raises(ValueError):
            boot.ci((self.x, self.z), boot.diff_of_means, n_samples=1000, multi="indepedent")

    def test_non_callable(self):                              # test:321-323
        with pytest.raises(TypeError):
            boot.ci(self.data, "average")

    def test_abc_with_returndist(self):                       # test:325-327
        with pytest.raises(ValueError):
            boot.ci(self.data, method="abc", return_dist=True)

    def test_pi_multi_2dout_multialpha(self):                 # test:329-355
        self._vector_stat_columns("pi")

    def _vector_stat_columns(self, method):
        kw = dict(alpha=(0.1, 0.2, 0.8, 0.9), n_samples=2000, method=method, seed=self.seed)
        both = boot.ci((self.x, self.y), boot.linear_fit, **kw)
        assert_allclose(both[:, 0], boot.ci((self.x, self.y), boot.fit_slope, **kw))
        assert_allclose(both[:, 1], boot.ci((self.x, self.y), boot.fit_intercept, **kw))
        table = np.vstack((self.x, self.y)).T.astype(float)
        results1 = boot.ci(table, boot.mean_axis0, alpha=(0.1, 0.2, 0.8, 0.9), n_samples=2000, method=method,
                           seed=self.seed)
        for col in (0, 1):
            single = boot.ci(table[:, col], np.mean, alpha=(0.1, 0.2, 0.8, 0.9), n_samples=2000, method=method,
                             seed=self.seed)
            assert_allclose(results1[:, col], single)

    def test_bca_n_samples(self):                             # test:357-363 (reference's indices)
        results = boot.ci_indexed(self.data, None, self.ref_indices(20, 500), alpha=(0.1, 0.2, 0.8, 0.9))
        assert_allclose(results, np.array([0.37248, 0.507976, 0.92783, 1.039755]), rtol=1e-6)

    def test_pi_simple(self):                                 # test:365-370
        results = boot.ci(self.data, method="pi", seed=self.seed)
        results2 = boot.ci(self.data, method="pi", alpha=(0.025, 1 - 0.025), seed=self.seed)
        assert_allclose(results, results2)

    def test_pandas_series(self):                             # test:372-388
        import pandas as pd
        pds = pd.Series(self.data, index=np.arange(50, 70))
        for method in ("abc", "bca", "pi"):
            assert_allclose(boot.ci(pds, method=method, seed=self.seed),
                            boot.ci(self.data, method=method, seed=self.seed))

    def test_invalid_multi(self):                             # test:390-394
        with pytest.raises(ValueError, match=r"Value `wrong` for multi is not recognized."):
            boot.ci(self.data, multi="wrong")

    def test_pval(self):                                      # test:396-403 (reference's indices)
        result = boot.pval_indexed(self.data, compfunction=lambda s: 0.8 <= s <= 1.2,
                                   indices=self.ref_indices(20, 500))
        assert_allclose(result, 0.368)

    def test_pval_default(self):                              # test:405-411
        assert_allclose(boot.pval_indexed(self.z, indices=self.ref_indices(7, 500)), 0.096)

    def test_pval_implicit_and_explicit_multi(self):          # test:413-430 (vector lambda -> two scalar stats)
        idx = self.ref_indices(6, 500)
        rx = boot.pval_indexed(self.x, np.average, boot.less_equal(3), indices=idx)
        ry = boot.pval_indexed(self.y, np.average, boot.less_equal(3), indices=idx)
        assert_allclose([rx, ry], [0.262, 0.936])
        a = boot.pval((self.x, self.y), boot.ratio_of_means, boot.less_equal(3), n_samples=500, seed=self.seed)
        b = boot.pval((self.x, self.y), boot.ratio_of_means, boot.less_equal(3), n_samples=500, seed=self.seed,
                      multi=True)
        assert_allclose(a, b)
