"""scikits-bootstrap_b200 — the scikits.bootstrap resampling hot path on B200.

Import as `scikits_bootstrap_b200` (repo-root shim) and use like the reference:

    import scikits_bootstrap_b200 as boot
    boot.ci(x, np.mean, method="bca", n_samples=100_000, seed=0)
"""
from . import _dist as distributed
from ._abi import B200BootError
from ._registry import (REGISTRY_NAMES, DeviceCompare, DeviceStat, between, diff_of, diff_of_means,
                        independent, ratio_of,
                        fit_intercept, fit_slope, greater_equal, greater_than, less_equal,
                        less_than, linear_fit, mean_axis0,
                        median_axis0, pearson_r, percentile, percentile_axis0, positive, quantile,
                        quantile_axis0, ratio_of_means, std_axis0, var_axis0,
                        weighted_average)
from .bootstrap import (InstabilityWarning, __version__, bootstrap_indices,
                        bootstrap_indices_independent, bootstrap_indices_moving_block, ci,
                        ci_indexed, ci_moving_block, jackknife_indices, pval, pval_indexed, subsample_indices)

__all__ = (
    "ci", "pval", "bootstrap_indices", "bootstrap_indices_independent", "subsample_indices",
    "jackknife_indices", "bootstrap_indices_moving_block", "InstabilityWarning",
    "ratio_of_means", "weighted_average", "pearson_r", "linear_fit", "fit_slope", "fit_intercept",
    "mean_axis0", "var_axis0", "std_axis0",
    "median_axis0", "quantile", "percentile", "quantile_axis0", "percentile_axis0", "diff_of_means", "diff_of", "ratio_of", "independent", "less_than", "less_equal", "greater_than", "greater_equal",
    "between", "positive", "ci_indexed", "ci_moving_block", "pval_indexed", "distributed", "REGISTRY_NAMES", "B200BootError",
)
