"""Host-side interval math: O(len(alphas) * D) scalar work that stays in numpy.

Follows /root/reference/src/scikits/bootstrap/bootstrap.py:
  :75-79   ncdf  (math.erf based)
  :83-119  nppf  (Acklam's rational approximation; its ~1e-9 error is part of the
                  reference's answer, so the same form and constants are used)
  :583, :627-629  z0 and the adjusted quantile levels
  :456-480        order-statistic indices and the three instability conditions
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

_A_TAIL = (-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e00,
           -2.549732539343734e00, 4.374664141464968e00, 2.938163982698783e00)
_B_TAIL = (7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e00,
           3.754408661907416e00)
_A_MID = (-3.969683028665376e01, 2.209460984245205e02, -2.759285104469687e02,
          1.383577518672690e02, -3.066479806614716e01, 2.506628277459239e00)
_B_MID = (-5.447609879822406e01, 1.615858368580409e02, -1.556989798598866e02,
          6.680131188771972e01, -1.328068155288572e01)
_SPLIT_LO, _SPLIT_HI = 0.02425, 0.97575
_INV_SQRT2_DIV = math.sqrt(2)


def _poly(coef, t):
    r = np.full_like(t, coef[0], dtype=float)
    for c in coef[1:]:
        r = r * t + c
    return r


def _tail(t):
    return _poly(_A_TAIL, t) / (_poly(_B_TAIL, t) * t + 1)


def _horner(coef, t):
    r = coef[0]
    for c in coef[1:]:
        r = r * t + c
    return r


def _nppf_scalar(p: float) -> float:
    """One element of nppf in Python floats: the same IEEE operations in the same order as
    the array form below (np.log on a float64 scalar runs the same ufunc loop; sqrt is
    correctly rounded everywhere), so the two agree bit for bit — tests/test_host_probes.py."""
    if p != p:
        return math.nan
    if p <= 0.0:
        return -math.inf if p == 0.0 else math.nan
    if p >= 1.0:
        return math.inf if p == 1.0 else math.nan
    if p < _SPLIT_LO:
        t = math.sqrt(-2 * float(np.log(np.float64(p))))
        return _horner(_A_TAIL, t) / (_horner(_B_TAIL, t) * t + 1)
    if p <= _SPLIT_HI:
        q = p - 0.5
        s = q * q
        return (_horner(_A_MID, s) * q) / (_horner(_B_MID, s) * s + 1)
    t = math.sqrt(-2 * float(np.log(np.float64(1 - p))))
    return -(_horner(_A_TAIL, t) / (_horner(_B_TAIL, t) * t + 1))


_SCALAR_LIMIT = 16     # below this many elements the per-element path beats numpy's masks


def nppf(p) -> np.ndarray:
    """Inverse standard-normal cdf, Acklam form; -inf / +inf at p == 0 / 1."""
    p = np.asarray(p, dtype=float)
    if p.size <= _SCALAR_LIMIT:
        out = np.empty(p.shape)
        flat = out.reshape(-1)
        for i, v in enumerate(p.reshape(-1).tolist()):
            flat[i] = _nppf_scalar(v)
        return out
    return _nppf_array(p)


def _nppf_array(p: np.ndarray) -> np.ndarray:
    out = np.full(p.shape, np.nan)
    with np.errstate(all="ignore"):
        lo = (p > 0) & (p < _SPLIT_LO)
        out[lo] = _tail(np.sqrt(-2 * np.log(p[lo])))
        mid = (p >= _SPLIT_LO) & (p <= _SPLIT_HI)
        q = p[mid] - 0.5
        s = q * q
        out[mid] = (_poly(_A_MID, s) * q) / (_poly(_B_MID, s) * s + 1)
        hi = (p > _SPLIT_HI) & (p < 1)
        out[hi] = -_tail(np.sqrt(-2 * np.log(1 - p[hi])))
    out[p == 0] = -np.inf
    out[p == 1] = np.inf
    return out


def ncdf(x) -> np.ndarray:
    """Standard-normal cdf as the reference evaluates it (np.vectorize over math.erf,
    bootstrap.py:72-76): a handful of values go through math.erf, arrays through the library's
    host loop over the C library's erf — the very function math.erf calls, so the bits are the
    same (tests/test_host_probes.py compares the two paths)."""
    x = np.asarray(x, dtype=float)
    out = np.empty(x.shape)
    if x.size > _SCALAR_LIMIT:
        from . import _abi
        xc = np.ascontiguousarray(x)
        _abi.lib().b200boot_host_ncdf(xc.ctypes.data, out.ctypes.data, xc.size)
        return out
    oi = out.reshape(-1)
    for i, v in enumerate(x.reshape(-1)):
        oi[i] = 0.5 * (1 + math.erf(v / _INV_SQRT2_DIV))
    return out


def _bca_levels_scalar(less: float, n_samples: int, accel: float, alphas: np.ndarray) -> np.ndarray:
    """bca_levels for a scalar statistic and a handful of levels in Python floats: the same IEEE
    operations in the same order as the array form (tests/test_host_probes.py compares them bit
    for bit).  Raises ZeroDivisionError / OverflowError where numpy would produce inf or nan;
    the caller then takes the array form."""
    z0 = _nppf_scalar((1.0 * less) / n_samples)
    out = np.empty(alphas.shape)
    flat = out.reshape(-1)
    for i, al in enumerate(alphas.reshape(-1).tolist()):
        zs = z0 + _nppf_scalar(al)
        flat[i] = 0.5 * (1 + math.erf((z0 + zs / (1 - accel * zs)) / _INV_SQRT2_DIV))
    return out


def bca_levels(less_counts: np.ndarray, n_samples: int, accel: np.ndarray,
               alphas: np.ndarray) -> np.ndarray:
    """Adjusted quantile levels: shape alphas.shape + stat shape."""
    if np.ndim(less_counts) == 0 and np.ndim(accel) == 0 and alphas.size <= _SCALAR_LIMIT:
        a, c = float(accel), float(less_counts)
        if math.isfinite(a) and 0.0 < c < n_samples:          # finite z0 and acceleration: no inf / nan arithmetic
            try:
                return _bca_levels_scalar(c, n_samples, a, alphas)
            except (ZeroDivisionError, OverflowError, ValueError):
                pass
    z0 = nppf((1.0 * np.asarray(less_counts)) / n_samples)
    with np.errstate(invalid="ignore", divide="ignore"):
        zs = z0 + nppf(alphas).reshape(alphas.shape + (1,) * z0.ndim)
        return ncdf(z0 + zs / (1 - accel * zs))


def order_indices(avals: np.ndarray, n_samples: int) -> Tuple[np.ndarray, List[str]]:
    """round((n-1)*avals) half-to-even as int (NaN -> 0) and which instability
    conditions hold: 'nan', then 'extremal' or else 'top10'."""
    avals = np.asarray(avals, dtype=float)
    if avals.size <= _SCALAR_LIMIT:
        # a handful of levels: Python's round() is half-to-even like np.round, and
        # (n - 1) * a is the same IEEE product
        top = n_samples - 1
        vals = [top * a for a in avals.reshape(-1).tolist()]
        if all(v != v or abs(v) < 4.0e18 for v in vals):       # else: numpy's own int conversion
            nan = extremal = top10 = False
            ints = []
            for v in vals:
                if v != v:
                    nan = True
                    ints.append(0)
                    continue
                k = round(v)
                ints.append(k)
                if k == 0 or k == top:
                    extremal = True
                elif k < 10 or k >= n_samples - 10:
                    top10 = True
            flags: List[str] = []
            if nan:
                flags.append("nan")
            if extremal:
                flags.append("extremal")
            elif top10:
                flags.append("top10")
            return np.array(ints, dtype="int").reshape(avals.shape), flags
    return _order_indices_array(avals, n_samples)


def _order_indices_array(avals: np.ndarray, n_samples: int) -> Tuple[np.ndarray, List[str]]:
    flags: List[str] = []
    with np.errstate(invalid="ignore"):
        nvals = np.round((n_samples - 1) * avals)
        if np.any(np.isnan(nvals)):
            flags.append("nan")
        if np.any(nvals == 0) or np.any(nvals == n_samples - 1):
            flags.append("extremal")
        elif np.any(nvals < 10) or np.any(nvals >= n_samples - 10):
            flags.append("top10")
    return np.nan_to_num(nvals).astype("int"), flags
