"""Case table shared by the golden generator (make_golden.py, runs the reference)
and the tests (which never import the reference).  Plain numpy stand-ins for
each registry statistic live in STATS so the same definitions feed both sides."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 1234567890

# the 20-point fixture of the reference's own tests (tests/bootstrap_test.py:20-43)
DATA20 = np.array([1.34016346, 1.73759123, 1.49898834, -0.22864333, 2.031034, 2.17032495,
                   1.59645265, -0.76945156, 0.56605824, -0.11927018, -0.1465108, -0.79890338,
                   0.77183278, -0.82819136, 1.32667483, 1.05986776, 2.14408873, -1.43464512,
                   2.28743654, 0.42864858])

STATS = {
    "mean": np.mean,
    "average": np.average,
    "median": np.median,
    "var": np.var,
    "std": np.std,
    "ratio_of_means": lambda a, b: np.mean(a) / np.mean(b),
    "weighted_average": lambda x, w: np.sum(x * w) / np.sum(w),
    "pearson_r": lambda a, b: np.corrcoef(a, b)[0, 1],
    "median_axis0": lambda a: np.median(a, axis=0),
    "mean_axis0": lambda a: np.mean(a, axis=0),
    "diff_of_means": lambda a, b: np.mean(a) - np.mean(b),
    # the statistic of the reference's paired tests (tests/bootstrap_test.py:211-238, 286-355)
    "linear_fit": lambda a, b: np.polyfit(a, b, 1),
    "percentile90": lambda x: np.percentile(x, 90),
    "quantile25_axis0": lambda a: np.quantile(a, 0.25, axis=0),
}


def gen_data(spec):
    kind = spec[0]
    if kind == "data20":
        return (DATA20,)
    if kind == "normal":
        _, seed, n = spec
        return (np.random.default_rng(seed).standard_normal(n),)
    if kind == "ints":
        _, seed, n = spec
        return (np.random.default_rng(seed).integers(0, 7, n).astype(float),)
    if kind == "gamma2":
        _, s1, s2, n = spec
        return (np.random.default_rng(s1).gamma(2, 1, n), np.random.default_rng(s2).gamma(3, 1, n))
    if kind == "gamma2_unequal":
        _, s1, s2, n1, n2 = spec
        return (np.random.default_rng(s1).gamma(2, 1, n1), np.random.default_rng(s2).gamma(3, 1, n2))
    if kind == "normal2d":
        _, seed, n, d = spec
        return (np.random.default_rng(seed).standard_normal((n, d)),)
    raise KeyError(kind)


# (case name, data spec, stat name, multi, n_samples, alphas)
CASES = [
    ("data20_mean", ("data20",), "mean", None, 2000, (0.1, 0.2, 0.8, 0.9)),
    ("data20_average_n500", ("data20",), "average", None, 500, (0.1, 0.2, 0.8, 0.9)),
    ("cfg1_mean", ("normal", 0, 1000), "mean", None, 10000, (0.025, 0.975)),
    ("normal300_var", ("normal", 11, 300), "var", None, 2000, (0.025, 0.975)),
    ("normal300_std", ("normal", 12, 300), "std", None, 2000, (0.05, 0.5, 0.95)),
    ("normal301_median", ("normal", 13, 301), "median", None, 2000, (0.025, 0.975)),
    ("normal300_median", ("normal", 14, 300), "median", None, 2000, (0.025, 0.975)),
    ("ints50_mean", ("ints", 15, 50), "mean", None, 2000, (0.025, 0.975)),
    ("ints50_median", ("ints", 16, 50), "median", None, 2000, (0.025, 0.975)),
    ("gamma200_ratio", ("gamma2", 2, 3, 200), "ratio_of_means", "paired", 2000, (0.025, 0.975)),
    ("gamma200_wavg", ("gamma2", 4, 5, 200), "weighted_average", "paired", 2000, (0.025, 0.975)),
    ("gamma200_pearson", ("gamma2", 6, 7, 200), "pearson_r", "paired", 2000, (0.025, 0.975)),
    ("normal2d_median_axis0", ("normal2d", 8, 101, 7), "median_axis0", None, 1500, (0.025, 0.975)),
    ("normal2d_mean_axis0", ("normal2d", 9, 64, 5), "mean_axis0", None, 1500, (0.1, 0.9)),
    ("indep_diff", ("gamma2_unequal", 20, 21, 60, 45), "diff_of_means", "independent", 1500,
     (0.025, 0.975)),
    ("gamma200_polyfit", ("gamma2", 22, 23, 200), "linear_fit", "paired", 2000, (0.025, 0.975)),
    ("normal300_p90", ("normal", 24, 300), "percentile90", None, 2000, (0.025, 0.975)),
    ("normal2d_q25_axis0", ("normal2d", 25, 101, 7), "quantile25_axis0", None, 1500, (0.05, 0.95)),
]
