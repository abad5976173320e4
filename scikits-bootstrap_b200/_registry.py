"""Device statistic registry.

The reference dispatches its compiled fast path by callable identity
(`statfunction in (np.average, np.mean)`, bootstrap.py:417, 588) and raises for
anything else (bootstrap.py:426-427).  This registry is the same idea with more
entries: a callable is recognised by identity, by being one of the exported
`DeviceStat` objects below, or as `functools.partial(np.<f>, axis=0)`.
Everything else raises — there is no host fallback.

Every `DeviceStat` is also an ordinary numpy callable, so the very same object
can be handed to the reference / the CPU oracle for parity checks."""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _abi


@dataclass(frozen=True)
class StatSpec:
    name: str
    stat_id: int
    n_arrays: int
    axis0: bool = False          # column-wise over 2-D data (rows share the index set)
    independent: bool = False    # defined for multi='independent'
    n_out: int = 1               # length of the statistic's value (1: scalar)
    q: Optional[float] = None    # rank statistics: None = np.median, else np.quantile(x, q)
    op: Optional[int] = None     # independent: the binary operation f of f(s_a(a), s_b(b))
    inner: tuple = ()            # independent: the specs of s_a and s_b


class DeviceStat:
    """A statistic the GPU path implements; callable on numpy arrays as well."""

    def __init__(self, spec: StatSpec, host: Callable, doc: str):
        self.spec = spec
        self._host = host
        self.__doc__ = doc
        self.__name__ = spec.name

    def __call__(self, *arrays, **kw):
        return self._host(*arrays, **kw)

    def __repr__(self):
        return f"<b200 device statistic {self.spec.name}>"


def _ratio(a, b):
    return np.mean(a) / np.mean(b)


def _wavg(x, w):
    x = np.asarray(x, dtype=float)
    w = np.asarray(w, dtype=float)
    return np.sum(x * w) / np.sum(w)


def _pearson(a, b):
    return np.corrcoef(a, b)[0, 1]


ratio_of_means = DeviceStat(StatSpec("ratio_of_means", _abi.STAT_RATIO_OF_MEANS, 2), _ratio,
                            "mean(a) / mean(b) of two paired arrays")
weighted_average = DeviceStat(StatSpec("weighted_average", _abi.STAT_WEIGHTED_AVERAGE, 2), _wavg,
                              "sum(x*w) / sum(w) of paired (x, w)")
pearson_r = DeviceStat(StatSpec("pearson_r", _abi.STAT_PEARSON_R, 2), _pearson,
                       "Pearson correlation of two paired arrays")
linear_fit = DeviceStat(StatSpec("linear_fit", _abi.STAT_LINEAR_FIT, 2, n_out=2),
                        lambda a, b: np.polyfit(a, b, 1),
                        "np.polyfit(a, b, 1) = [slope, intercept] of the least-squares line through paired "
                        "(a, b) (the statistic of the reference's paired tests)")
fit_slope = DeviceStat(StatSpec("fit_slope", _abi.STAT_SLOPE, 2), lambda a, b: np.polyfit(a, b, 1)[0],
                       "np.polyfit(a, b, 1)[0]")
fit_intercept = DeviceStat(StatSpec("fit_intercept", _abi.STAT_INTERCEPT, 2),
                           lambda a, b: np.polyfit(a, b, 1)[1], "np.polyfit(a, b, 1)[1]")
mean_axis0 = DeviceStat(StatSpec("mean_axis0", _abi.STAT_MEAN, 1, axis0=True),
                        lambda a: np.mean(a, axis=0), "np.mean(a, axis=0) over (N, D) data")
var_axis0 = DeviceStat(StatSpec("var_axis0", _abi.STAT_VAR, 1, axis0=True),
                       lambda a: np.var(a, axis=0), "np.var(a, axis=0) over (N, D) data")
std_axis0 = DeviceStat(StatSpec("std_axis0", _abi.STAT_STD, 1, axis0=True),
                       lambda a: np.std(a, axis=0), "np.std(a, axis=0) over (N, D) data")
median_axis0 = DeviceStat(StatSpec("median_axis0", _abi.STAT_MEDIAN, 1, axis0=True),
                          lambda a: np.median(a, axis=0), "np.median(a, axis=0) over (N, D) data")


def _quantile_spec(q, axis0: bool) -> StatSpec:
    q = float(q)
    if not 0.0 <= q <= 1.0:
        raise ValueError("Quantiles must be in the range [0, 1]")          # numpy's message
    return StatSpec("quantile_axis0" if axis0 else "quantile", _abi.STAT_MEDIAN, 1, axis0=axis0, q=q)


def quantile(q) -> DeviceStat:
    """np.quantile(x, q) (default method 'linear') of a 1-D array, as a device statistic."""
    spec = _quantile_spec(q, False)
    return DeviceStat(spec, lambda a: np.quantile(a, spec.q), f"np.quantile(x, {spec.q})")


def quantile_axis0(q) -> DeviceStat:
    """np.quantile(a, q, axis=0) over (N, D) data."""
    spec = _quantile_spec(q, True)
    return DeviceStat(spec, lambda a: np.quantile(a, spec.q, axis=0), f"np.quantile(a, {spec.q}, axis=0)")


def percentile(p) -> DeviceStat:
    """np.percentile(x, p): the quantile p / 100."""
    return quantile(np.true_divide(p, 100))


def percentile_axis0(p) -> DeviceStat:
    """np.percentile(a, p, axis=0) over (N, D) data."""
    return quantile_axis0(np.true_divide(p, 100))


def _partial_quantile(part: functools.partial) -> Optional[StatSpec]:
    """functools.partial(np.quantile | np.percentile, q=..[, axis=0][, method='linear'])."""
    kw = dict(part.keywords)
    if part.args or "q" not in kw or not np.isscalar(kw["q"]):
        return None
    if kw.pop("method", "linear") != "linear":
        return None
    q = kw.pop("q")
    axis0 = False
    if "axis" in kw:
        if kw.pop("axis") != 0:
            return None
        axis0 = True
    if kw:
        return None
    if part.func is np.percentile:
        q = np.true_divide(q, 100)
    return _quantile_spec(q, axis0)


# ---- multi='independent': f(s_a(a), s_b(b)) of two independent samples ---------------------
_OPS = {
    "sub": (_abi.OP_SUB, lambda u, v: u - v, "{a}(a) - {b}(b)"),
    "add": (_abi.OP_ADD, lambda u, v: u + v, "{a}(a) + {b}(b)"),
    "div": (_abi.OP_DIV, lambda u, v: u / v, "{a}(a) / {b}(b)"),
    "mul": (_abi.OP_MUL, lambda u, v: u * v, "{a}(a) * {b}(b)"),
    "log_ratio": (_abi.OP_LOG_RATIO, lambda u, v: np.log(u / v), "log({a}(a) / {b}(b))"),
}


def independent(op: str, stat_a, stat_b=None) -> DeviceStat:
    """f(s_a(a), s_b(b)) for two INDEPENDENT samples (multi='independent'; the arrays may differ
    in length, each is resampled on its own: bootstrap.py:438-444).  s_a, s_b: any scalar
    single-array statistics of the registry (np.mean, np.var, np.std, np.sum, np.median,
    quantile(q), percentile(p)); s_b defaults to s_a.  f: 'sub', 'add', 'div', 'mul' or
    'log_ratio'.  The BCa jackknife pools the N_a + N_b leave-one-out values as the reference
    does (bootstrap.py:601-607, 708-714)."""
    if op not in _OPS:
        raise ValueError(f"unknown operation {op!r}; one of {sorted(_OPS)}")
    host_a, host_b = stat_a, (stat_a if stat_b is None else stat_b)
    inner = tuple(require(f) for f in (host_a, host_b))
    for sp in inner:
        if sp.n_arrays != 1 or sp.axis0 or sp.n_out != 1 or sp.independent:
            raise NotImplementedError(f"independent() needs scalar statistics of one 1-D array, got {sp.name}")
    op_id, host_op, text = _OPS[op]
    name = text.format(a=inner[0].name, b=inner[1].name)
    spec = StatSpec(f"independent[{name}]", inner[0].stat_id, 2, independent=True, q=inner[0].q,
                    op=op_id, inner=inner)
    return DeviceStat(spec, lambda a, b: host_op(host_a(a), host_b(b)), name + " of independent samples")


def diff_of(statfunction) -> DeviceStat:
    """s(a) - s(b) of independent samples: independent('sub', s)."""
    return independent("sub", statfunction)


def ratio_of(statfunction) -> DeviceStat:
    """s(a) / s(b) of independent samples: independent('div', s)."""
    return independent("div", statfunction)


_BY_IDENTITY = {
    id(np.mean): StatSpec("mean", _abi.STAT_MEAN, 1),
    id(np.average): StatSpec("mean", _abi.STAT_MEAN, 1),
    id(np.var): StatSpec("var", _abi.STAT_VAR, 1),
    id(np.std): StatSpec("std", _abi.STAT_STD, 1),
    id(np.median): StatSpec("median", _abi.STAT_MEDIAN, 1),
    id(np.sum): StatSpec("sum", _abi.STAT_SUM, 1),
}
_AXIS0 = {
    id(np.mean): mean_axis0.spec, id(np.average): mean_axis0.spec, id(np.var): var_axis0.spec,
    id(np.std): std_axis0.spec, id(np.median): median_axis0.spec,
}

REGISTRY_NAMES = ("np.mean, np.average, np.var, np.std, np.median, np.sum, "
                  "functools.partial(np.<mean|average|var|std|median>, axis=0), "
                  "quantile(q), percentile(p), quantile_axis0(q), percentile_axis0(p), "
                  "functools.partial(np.<quantile|percentile>, q=...[, axis=0]), "
                  "ratio_of_means, weighted_average, pearson_r, linear_fit, fit_slope, fit_intercept, mean_axis0, var_axis0, std_axis0, "
                  "median_axis0, diff_of_means, diff_of(s), ratio_of(s), "
                  "independent(<'sub'|'add'|'div'|'mul'|'log_ratio'>, s_a[, s_b])")


def lookup(statfunction: Callable) -> Optional[StatSpec]:
    if isinstance(statfunction, DeviceStat):
        return statfunction.spec
    if isinstance(statfunction, functools.partial):
        if statfunction.func is np.quantile or statfunction.func is np.percentile:
            return _partial_quantile(statfunction)
        if (not statfunction.args and statfunction.keywords == {"axis": 0}
                and id(statfunction.func) in _AXIS0):
            return _AXIS0[id(statfunction.func)]
        return None
    return _BY_IDENTITY.get(id(statfunction))


def require(statfunction: Callable) -> StatSpec:
    spec = lookup(statfunction)
    if spec is None:
        raise NotImplementedError(
            "statfunction {!r} is not in the device statistic registry and there is no host "
            "fallback; registered: {}".format(statfunction, REGISTRY_NAMES))
    return spec


diff_of_means = diff_of(np.mean)     # mean(a) - mean(b): the reference's own independent-mode test statistic


# ---- comparison functions for pval ------------------------------------------------

class DeviceCompare:
    """A pval criterion evaluated on the device (count reduction); also callable
    on host values."""

    def __init__(self, op: int, t0: float, t1: float = 0.0):
        self.op, self.t0, self.t1 = op, float(t0), float(t1)

    def __call__(self, s):
        if self.op == _abi.CMP_LT:
            return s < self.t0
        if self.op == _abi.CMP_LE:
            return s <= self.t0
        if self.op == _abi.CMP_GT:
            return s > self.t0
        if self.op == _abi.CMP_GE:
            return s >= self.t0
        return (self.t0 <= s) & (s <= self.t1)


def less_than(c):
    return DeviceCompare(_abi.CMP_LT, c)


def less_equal(c):
    return DeviceCompare(_abi.CMP_LE, c)


def greater_than(c):
    return DeviceCompare(_abi.CMP_GT, c)


def greater_equal(c):
    return DeviceCompare(_abi.CMP_GE, c)


def between(lo, hi):
    return DeviceCompare(_abi.CMP_BETWEEN, lo, hi)


positive = greater_than(0.0)   # the reference's default `lambda s: s > 0` (bootstrap.py:793)
