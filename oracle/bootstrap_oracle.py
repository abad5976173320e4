"""CPU oracle for the bootstrap resampling hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the algorithm of the reference
(`/root/reference/src/scikits/bootstrap/bootstrap.py`, cited below as
`bootstrap.py:L`).  It exists so the CUDA path can be checked against it; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it.  The product package never imports anything from `oracle/`.

Parity pin: the restatement is checked (tests/test_oracle_pins.py) against
  * every golden vector the reference's own test-suite holds for this path
    (`tests/bootstrap_test.py:73-84, 110-112, 163-169, 205-209, 240-248,
    357-363, 396-430`), and
  * fixtures under `tests/golden/` produced by importing the reference itself in
    the build container (`tests/golden/make_golden.py`).

Third-party arithmetic on the path: numpy (`numpy>=1.22`, un-pinned upstream;
2.3.5 here) supplies PCG64 + Lemire-32 bounded integers through
`Generator.integers`.  The oracle calls the same numpy entry point at the same
call sites the reference does (`bootstrap.py:359, 677-680, 688-691`), so the
PCG64 stream is reproduced by construction.

The second half of the file restates the *device* generator (Philox4x32-10 and
the tile/lane layout the kernels use) so device-materialised indices can be
predicted bit for bit on the host.
"""
from __future__ import annotations

import math
import warnings
from typing import Callable, Iterable, Iterator, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Normal cdf / Acklam inverse cdf                     bootstrap.py:72-119
# --------------------------------------------------------------------------

_SQRT2 = math.sqrt(2)

# Acklam's coefficients exactly as the reference carries them
# (bootstrap.py:89-93 tails, :100-105 centre).  The approximation error of this
# rational form (~1e-9) is part of the reference's answer, so an exact ppf must
# not be substituted.
_TAIL_NUM = (-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e00,
             -2.549732539343734e00, 4.374664141464968e00, 2.938163982698783e00)
_TAIL_DEN = (7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e00,
             3.754408661907416e00)
_MID_NUM = (-3.969683028665376e01, 2.209460984245205e02, -2.759285104469687e02,
            1.383577518672690e02, -3.066479806614716e01, 2.506628277459239e00)
_MID_DEN = (-5.447609879822406e01, 1.615858368580409e02, -1.556989798598866e02,
            6.680131188771972e01, -1.328068155288572e01)
_P_LOW = 0.02425
_P_HIGH = 0.97575


def _horner(coeffs: Sequence[float], t: np.ndarray) -> np.ndarray:
    acc = np.full_like(t, coeffs[0], dtype=float)
    for c in coeffs[1:]:
        acc = acc * t + c
    return acc


def ncdf(x) -> np.ndarray:
    """0.5*(1+erf(x/sqrt2)) element-wise via math.erf (bootstrap.py:75-79)."""
    xa = np.asarray(x, dtype=float)
    out = np.empty(xa.shape, dtype=float)
    flat_in = xa.ravel()
    flat_out = out.ravel()
    for i, v in enumerate(flat_in):
        flat_out[i] = 0.5 * (1 + math.erf(v / _SQRT2))
    return out


def nppf(p) -> np.ndarray:
    """Acklam inverse normal cdf (bootstrap.py:83-119): three branches split at
    0.02425 / 0.97575, -inf at p==0, +inf at p==1, NaN elsewhere outside (0,1)."""
    pa = np.asarray(p, dtype=float)
    out = np.full(pa.shape, np.nan, dtype=float)
    with np.errstate(all="ignore"):
        lo = (pa < _P_LOW) & (pa > 0)
        t = np.sqrt(-2 * np.log(pa[lo]))
        # denominator Horner ends "... * t + 1" (bootstrap.py:93)
        out[lo] = _horner(_TAIL_NUM, t) / (_horner(_TAIL_DEN, t) * t + 1)

        mid = (_P_LOW <= pa) & (pa <= _P_HIGH)
        q = pa[mid] - 0.5
        s = q * q
        out[mid] = (_horner(_MID_NUM, s) * q) / (_horner(_MID_DEN, s) * s + 1)

        hi = (pa > _P_HIGH) & (pa < 1)
        t = np.sqrt(-2 * np.log(1 - pa[hi]))
        out[hi] = -(_horner(_TAIL_NUM, t) / (_horner(_TAIL_DEN, t) * t + 1))
    out[pa == 0] = -np.inf
    out[pa == 1] = np.inf
    return out


# --------------------------------------------------------------------------
# Index plans (reference stream: numpy PCG64)     bootstrap.py:667-714
# --------------------------------------------------------------------------

def bootstrap_indices_pcg64(n_obs: int, n_samples: int, seed=None) -> Iterator[np.ndarray]:
    """`rng.integers(0, N, size=(N,))` n_samples times (bootstrap.py:677-680).
    A Generator passed as seed is used (and advanced) in place, as default_rng does."""
    rng = np.random.default_rng(seed)
    for _ in range(n_samples):
        yield rng.integers(low=0, high=n_obs, size=(n_obs,))


def bootstrap_indices_independent_pcg64(lengths: Sequence[int], n_samples: int,
                                        seed=None) -> Iterator[Tuple[np.ndarray, ...]]:
    """One index array per data array, drawn in tuple order from one shared
    stream (bootstrap.py:688-691)."""
    rng = np.random.default_rng(seed)
    for _ in range(n_samples):
        yield tuple(rng.integers(low=0, high=n, size=(n,)) for n in lengths)


def jackknife_indices(n_obs: int) -> Iterator[np.ndarray]:
    """Leave-one-out index sets in ascending order of the deleted row
    (bootstrap.py:704-705)."""
    base = np.arange(n_obs)
    for i in range(n_obs):
        yield np.concatenate((base[:i], base[i + 1:]))


# --------------------------------------------------------------------------
# Resample loop, jackknife, BCa tail, selection      bootstrap.py:429-499, 570-632
# --------------------------------------------------------------------------

def as_tdata(data, multi) -> Tuple[Tuple[np.ndarray, ...], object]:
    """Argument normalisation of bootstrap.py:362-379."""
    if multi not in [False, True, None, "independent", "paired"]:
        raise ValueError("Value `{}` for multi is not recognized.".format(multi))
    if multi is None:
        multi = isinstance(data, tuple)
    if multi and isinstance(data, Iterable):
        tdata = tuple(np.array(x) for x in data)
        lengths = [x.shape[0] for x in tdata]
        if len(set(lengths)) > 1 and multi != "independent":
            raise ValueError(
                "Data arrays have differing lengths {}, and multi is not set to 'independent'."
                .format(lengths))
    else:
        tdata = (np.array(data),)
    return tdata, multi


def stat_over_indices(tdata, statfunction: Callable, index_sets: Iterable) -> np.ndarray:
    """The hot loop: gather every array with the same index set and apply the
    statistic (bootstrap.py:429-432; independent mode :438-444)."""
    rows = []
    for idx in index_sets:
        if isinstance(idx, tuple):
            rows.append(statfunction(*(x[i] for x, i in zip(tdata, idx))))
        else:
            rows.append(statfunction(*(x[idx] for x in tdata)))
    return np.array(rows)


def jackknife_stats_loop(tdata, statfunction: Callable) -> np.ndarray:
    """The reference's O(N^2) paired leave-one-out loop (bootstrap.py:595-599)."""
    n = len(tdata[0])
    return np.array([statfunction(*(np.concatenate((x[:i], x[i + 1:])) for x in tdata))
                     for i in range(n)])


def jackknife_stats_independent_loop(tdata, statfunction: Callable) -> np.ndarray:
    """Independent mode: sum_k N_k leave-one-out samples, array-major
    (bootstrap.py:601-607, 708-714)."""
    rows = []
    for k, xk in enumerate(tdata):
        for j in range(len(xk)):
            args = [x if m != k else np.concatenate((x[:j], x[j + 1:]))
                    for m, x in enumerate(tdata)]
            rows.append(statfunction(*args))
    return np.array(rows)


def acceleration(jstat: np.ndarray) -> np.ndarray:
    """a = sum d^3 / (6 (sum d^2)^1.5) with d = mean(j) - j  (bootstrap.py:609-615)."""
    jstat = np.asarray(jstat, dtype=float)
    jmean = np.mean(jstat, axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        d = jmean - jstat
        return np.sum(d ** 3, axis=0) / (6.0 * np.sum(d ** 2, axis=0) ** 1.5)


def bca_avals(stat_sorted: np.ndarray, ostat, accel, alphas: np.ndarray, n_samples: int) -> np.ndarray:
    """z0 from the strict '<' count, then adjusted quantile levels
    (bootstrap.py:583, 627-629)."""
    z0 = nppf((1.0 * np.sum(stat_sorted < ostat, axis=0)) / n_samples)
    with np.errstate(invalid="ignore", divide="ignore"):
        zs = z0 + nppf(alphas).reshape(alphas.shape + (1,) * z0.ndim)
        return ncdf(z0 + zs / (1 - accel * zs))


def order_indices(avals: np.ndarray, n_samples: int):
    """nvals = round((n-1)*avals) half-to-even, the three instability flags and
    the NaN->0 integer cast (bootstrap.py:456-480).  Returns (nvals_int, flags)."""
    with np.errstate(invalid="ignore"):
        nvals = np.round((n_samples - 1) * avals)
        flags = []
        if np.any(np.isnan(nvals)):
            flags.append("nan")
        if np.any(nvals == 0) or np.any(nvals == n_samples - 1):
            flags.append("extremal")
        elif np.any(nvals < 10) or np.any(nvals >= n_samples - 10):
            flags.append("top10")
    return np.nan_to_num(nvals).astype("int"), flags


def pick(stat_sorted: np.ndarray, nvals: np.ndarray, output: str, ostat=None) -> np.ndarray:
    """Order-statistic pick and output formatting (bootstrap.py:482-499).
    Vector statistics use per-column nvals (:490)."""
    if nvals.ndim == 1:
        vals = stat_sorted[nvals]
    else:
        cols = np.indices(nvals.shape)[1:]
        vals = stat_sorted[(nvals,) + tuple(cols)]
    if output == "lowhigh":
        return vals
    if output == "errorbar":
        if nvals.ndim == 1:
            return np.abs(ostat - vals)[np.newaxis].T
        return np.abs(ostat - vals).T.squeeze()
    raise ValueError("Output option {0} is not supported.".format(output))


def ci_from_indices(tdata, statfunction, index_sets, alphas, n_samples, method="bca",
                    output="lowhigh", multi=False, jstat=None, return_parts=False):
    """Everything after index generation: bootstrap.py:429-504 with an explicit
    index source.  `jstat` may be supplied (closed form) to avoid the O(N^2) loop."""
    alphas = np.asarray(alphas, dtype=float)
    stat = stat_over_indices(tdata, statfunction, index_sets)
    stat.sort(axis=0)
    ostat = statfunction(*tdata)
    accel = None
    if method == "pi":
        avals = alphas
    elif method == "bca":
        if jstat is None:
            if multi == "independent":
                jstat = jackknife_stats_independent_loop(tdata, statfunction)
            else:
                jstat = jackknife_stats_loop(tdata, statfunction)
        accel = acceleration(jstat)
        avals = bca_avals(stat, ostat, accel, alphas, n_samples)
    else:
        raise ValueError("Method {0} is not supported.".format(method))
    nvals, flags = order_indices(avals, n_samples)
    out = pick(stat, nvals, output, ostat)
    if return_parts:
        return out, dict(stat=stat, ostat=ostat, accel=accel, avals=avals, nvals=nvals, flags=flags)
    return out


def ci(data, statfunction=None, alpha=0.05, n_samples=10000, method="bca", output="lowhigh",
       multi=None, return_dist=False, seed=None):
    """Restatement of the reference `ci` for method in {'pi','bca'} on the PCG64
    stream (bootstrap.py:352-504)."""
    alphas = np.array(alpha) if isinstance(alpha, Iterable) else np.array([alpha / 2, 1 - alpha / 2])
    rng = np.random.default_rng(seed)
    tdata, multi = as_tdata(data, multi)
    if statfunction is None:
        statfunction = np.average
    elif not callable(statfunction):
        raise TypeError("statfunction {} is not callable.".format(statfunction))
    if multi == "independent":
        idx = bootstrap_indices_independent_pcg64([len(x) for x in tdata], n_samples, rng)
    else:
        idx = bootstrap_indices_pcg64(len(tdata[0]), n_samples, rng)
    out, parts = ci_from_indices(tdata, statfunction, idx, alphas, n_samples, method, output,
                                 multi, return_parts=True)
    return (out, parts["stat"]) if return_dist else out


def pval(data, statfunction=np.average, compfunction=lambda s: s > 0, n_samples=10000,
         multi=None, seed=None):
    """Restatement of `pval` (bootstrap.py:839-866)."""
    rng = np.random.default_rng(seed)
    if multi is None:
        multi = isinstance(data, tuple)
    tdata = tuple(np.array(x) for x in data) if multi else (np.array(data),)
    stat = stat_over_indices(tdata, statfunction,
                             bootstrap_indices_pcg64(len(tdata[0]), n_samples, rng))
    stat.sort(axis=0)
    return np.mean([compfunction(s) for s in stat], axis=0)


def ci_abc(data, statfunction=np.average, alpha=0.05, epsilon=0.001, output="lowhigh"):
    """Restatement of the reference's ABC interval (bootstrap.py:507-567): finite-difference
    influence values t1, second differences t2, acceleration a, curvature cq, bias bhat,
    then statfunction at the tilted weights.  `statfunction(x, weights=...)` as in the reference."""
    alphas = np.array(alpha) if isinstance(alpha, Iterable) else np.array([alpha / 2, 1 - alpha / 2])
    x = np.array(data)
    nn = x.shape[0]
    n = nn * 1.0
    ep = epsilon / n
    p0 = np.repeat(1.0 / n, nn)
    t0 = statfunction(x, weights=p0)
    eye = np.identity(nn)
    tp = np.array([statfunction(x, weights=p0 + ep * (row - p0)) for row in eye], dtype=float)
    tm = np.array([statfunction(x, weights=p0 - ep * (row - p0)) for row in eye], dtype=float)
    t1 = (tp - tm) / (2 * ep)
    t2 = (tp - 2 * t0 + tm) / ep ** 2
    sighat = np.sqrt(np.sum(t1 ** 2)) / n
    a = np.sum(t1 ** 3) / (6 * n ** 3 * sighat ** 3)
    delta = t1 / (n ** 2 * sighat)
    cq = (statfunction(x, weights=p0 + ep * delta) - 2 * t0
          + statfunction(x, weights=p0 - ep * delta)) / (2 * sighat * ep ** 2)
    bhat = np.sum(t2) / (2 * n ** 2)
    curv = bhat / sighat - cq
    z0 = nppf(2 * ncdf(a) * ncdf(-curv))
    big_z = z0 + nppf(alphas)
    za = big_z / (1 - a * big_z) ** 2
    abc = np.array([statfunction(x, weights=p0 + z * delta) for z in za], dtype=float)
    if output == "lowhigh":
        return abc
    if output == "errorbar":
        return abs(abc - statfunction(x))[np.newaxis].T
    raise ValueError("Output option {0} is not supported.".format(output))


# --------------------------------------------------------------------------
# Closed-form leave-one-out statistics for the device registry.  These replace
# the O(N^2) loop (bootstrap.py:595-599) the same way the reference's own numba
# kernel does for the mean (bootstrap.py:639-650).  Each is validated against
# `jackknife_stats_loop` in tests/test_oracle_pins.py.
# --------------------------------------------------------------------------

def jack_mean(x):
    x = np.asarray(x, dtype=float)
    return (x.sum() - x) / (len(x) - 1)


def jack_var(x, take_sqrt=False):
    """Leave-one-out population variance (ddof=0) from centred moments:
    M2_{-i} = M2 - d_i^2 * N/(N-1),  d_i = x_i - mean."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    d = x - x.mean()
    m2 = np.sum(d * d) - d * d * n / (n - 1)
    v = m2 / (n - 1)
    return np.sqrt(v) if take_sqrt else v


def jack_ratio_of_means(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return (a.sum() - a) / (b.sum() - b)


def jack_weighted_average(x, w):
    x = np.asarray(x, dtype=float)
    w = np.asarray(w, dtype=float)
    xw = x * w
    return (xw.sum() - xw) / (w.sum() - w)


def jack_pearson(a, b):
    """Leave-one-out Pearson r from the five centred sums minus the i-th terms."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    n = len(a)
    da = a - a.mean()
    db = b - b.mean()
    m = n - 1
    sa = -da                      # sum of centred a without i (total is 0)
    sb = -db
    saa = np.sum(da * da) - da * da
    sbb = np.sum(db * db) - db * db
    sab = np.sum(da * db) - da * db
    return (sab - sa * sb / m) / np.sqrt((saa - sa * sa / m) * (sbb - sb * sb / m))


def jack_quantile(x, q):
    """Leave-one-out np.quantile(., q) (method 'linear'): like the median, the value depends
    only on whether the deleted element's rank is below, between or above the two order
    statistics the quantile of an (N-1)-sample reads."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    order = np.argsort(x, kind="stable")
    s = x[order]
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    m = n - 1
    vi = (m - 1) * np.float64(q)                  # numpy: _QuantileMethods['linear']
    if vi >= m - 1:
        k_lo = k_hi = m - 1
        gamma = np.float64(0.0)
    else:
        k_lo = int(np.floor(vi))
        k_hi = k_lo + 1
        gamma = vi - k_lo
    kth = lambda k: s[np.where(k < rank, k, k + 1)]          # k-th order statistic without `rank`
    a, b = kth(k_lo), kth(k_hi)
    diff = b - a                                             # numpy: _lerp
    return np.where(gamma >= 0.5, b - diff * (1 - gamma), a + diff * gamma)


def jack_linear_fit(a, b):
    """Leave-one-out np.polyfit(a, b, 1) = [slope, intercept] from centred sums minus the
    i-th terms -> float64[N][2]."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    n = len(a)
    ma, mb = a.mean(), b.mean()
    da = a - ma
    db = b - mb
    m = n - 1
    sa = -da
    sb = -db
    saa = np.sum(da * da) - da * da
    sab = np.sum(da * db) - da * db
    slope = (sab - sa * sb / m) / (saa - sa * sa / m)
    intercept = (mb + sb / m) - slope * (ma + sa / m)
    return np.stack([slope, intercept], axis=1)


def jack_median(x):
    """Leave-one-out medians take at most three distinct values, by rank of the
    deleted element in the sorted data."""
    x = np.asarray(x, dtype=float)
    n = len(x)
    order = np.argsort(x, kind="stable")
    s = x[order]
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    m = n - 1                      # length of each leave-one-out sample
    out = np.empty(n)
    for i in range(n):
        r = rank[i]

        def kth(k):                # k-th smallest (0-based) of s without position r
            return s[k] if k < r else s[k + 1]

        out[i] = kth(m // 2) if m % 2 else 0.5 * (kth(m // 2 - 1) + kth(m // 2))
    return out


# --------------------------------------------------------------------------
# numba restatement of the reference's compiled CPU path (bootstrap.py:639-664),
# used only as the timed CPU baseline when numba is importable.
# --------------------------------------------------------------------------

def numba_kernels():
    """Returns (boot_mean, jack_mean) njit kernels or None when numba is absent."""
    try:
        import numba
    except Exception:  # pragma: no cover
        return None

    @numba.njit(parallel=True, fastmath=True)
    def boot_mean(data, n_samples, seed):
        # one MT19937 reseed per resample id, then draw-with-replacement and mean
        n = data.shape[0]
        out = np.zeros(n_samples)
        for r in numba.prange(n_samples):
            np.random.seed(seed + r)
            out[r] = np.random.choice(data, n).mean()
        return out

    @numba.njit(parallel=True, fastmath=True)
    def jack_mean_nb(data):
        n = data.shape[0]
        tot = data.sum()
        out = np.zeros(n)
        for i in numba.prange(n):
            out[i] = (tot - data[i]) / (n - 1)
        return out

    return boot_mean, jack_mean_nb


# --------------------------------------------------------------------------
# Device generator restatement: Philox4x32-10 (Salmon et al., SC'11) and the
# counter layout of `csrc/philox.cuh` / `csrc/draw.cuh`.
# --------------------------------------------------------------------------

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


# Rounds of the device streams (csrc/philox.cuh: B2_PHILOX_ROUNDS, default 7); a test run
# against a library built with another count sets this from b200boot_philox_rounds().
PHILOX_ROUNDS = 7


def philox4x32(ctr: np.ndarray, key: Tuple[int, int], rounds: int = None) -> np.ndarray:
    """ctr: (..., 4) uint32 counters; key: (k0, k1).  Returns (..., 4) uint32."""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(PHILOX_ROUNDS if rounds is None else rounds):
        p0 = PHILOX_M0 * c[..., 0]
        p1 = PHILOX_M1 * c[..., 2]
        n0 = (p1 >> np.uint64(32)) ^ c[..., 1] ^ np.uint64(k0)
        n1 = p1 & _U32
        n2 = (p0 >> np.uint64(32)) ^ c[..., 3] ^ np.uint64(k1)
        n3 = p0 & _U32
        c[..., 0], c[..., 1], c[..., 2], c[..., 3] = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox4x32_10(ctr: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    return philox4x32(ctr, key, 10)


# Counter layout (must match csrc/philox.cuh):
#   key  = (seed & 0xffffffff, seed >> 32)
#   c0   = block index within the (resample, cell) stream
#   c1   = cell id (tile index; array id in bits 24..31 for independent draws)
#   c2   = resample id, low 32 bits
#   c3   = (resample id >> 32) & 0xffff  |  purpose << 16
PURPOSE_INDEX = 0      # index words
PURPOSE_REDRAW = 1     # Lemire rejection redraws
PURPOSE_BINOMIAL = 2   # tile occupancy (multinomial chain) uniforms


def _ctr(block, cell, resample, purpose):
    block = np.asarray(block, dtype=np.uint64)
    out = np.zeros(block.shape + (4,), dtype=np.uint64)
    out[..., 0] = block & _U32
    out[..., 1] = np.uint64(cell & 0xFFFFFFFF)
    out[..., 2] = np.uint64(resample & 0xFFFFFFFF)
    out[..., 3] = np.uint64(((resample >> 32) & 0xFFFF) | (purpose << 16))
    return out


def seed_key(seed: int) -> Tuple[int, int]:
    return (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


N_CLASSES = 16                 # csrc/draw.cuh: kClasses (shared-memory bank pairs)
CLASS_DRAWS_PER_BLOCK = 12


def class_cell_id(tile: int, c: int) -> int:
    return 0x80000000 | ((tile << 4) & 0x7FFFFFF0) | c


def class_draws(seed: int, resample: int, tile: int, c: int, n_draws: int, log2_tile: int) -> np.ndarray:
    """Rows (inside the tile) drawn for bank class c of full tile `tile`: each Philox
    word gives three (log2_tile-4)-bit row numbers j_A=(w>>7)&m, j_B=(w>>17)&m,
    j_C=rotl(w,5)&m; the row is j*16 + c.  12 draws per block."""
    nblk = (n_draws + CLASS_DRAWS_PER_BLOCK - 1) // CLASS_DRAWS_PER_BLOCK
    words = philox4x32(_ctr(np.arange(nblk), class_cell_id(tile, c), resample, PURPOSE_INDEX),
                          seed_key(seed)).astype(np.uint32)
    m = np.uint32((1 << (log2_tile - 4)) - 1)
    ja = (words >> np.uint32(7)) & m
    jb = (words >> np.uint32(17)) & m
    jc = ((words << np.uint32(5)) | (words >> np.uint32(27))) & m
    j = np.stack((ja, jb, jc), axis=-1).reshape(nblk, 12)      # w0:A,B,C, w1:A,B,C, ...
    return (j.reshape(-1)[:n_draws].astype(np.int64)) * N_CLASSES + c


def cell_draws_lemire(seed: int, resample: int, cell: int, n_draws: int, size: int,
                      purpose: int = 0) -> np.ndarray:
    """Indices in [0, size) for one (resample, cell): one 32-bit word per index,
    Lemire multiply-shift; a word whose low product falls below 2^32 mod size is
    rejected and redrawn from the PURPOSE_REDRAW stream keyed by the position."""
    nblk = (n_draws + 3) // 4
    key = seed_key(seed)
    words = philox4x32(_ctr(np.arange(nblk), cell, resample, purpose), key)
    words = words.reshape(-1)[:n_draws].astype(np.uint64)
    prod = words * np.uint64(size)
    idx = (prod >> np.uint64(32)).astype(np.int64)
    thresh = (1 << 32) % size
    bad = np.nonzero((prod & _U32) < np.uint64(thresh))[0]
    for p in bad:
        attempt = 0
        while True:
            c = _ctr(np.array([p]), cell, resample, PURPOSE_REDRAW)
            c[..., 3] |= np.uint64((attempt // 4) << 20)
            w = int(philox4x32(c, key)[0, attempt % 4])
            pr = w * size
            if (pr & 0xFFFFFFFF) >= thresh:
                idx[p] = pr >> 32
                break
            attempt += 1
    return idx


SMEM_DATA_BYTES = 224 * 1024     # csrc/draw.cuh: kSmemDataBytes


def tile_plan(n_obs: int, log2_tile: int, n_arrays: int = 1):
    """Cell sizes (csrc/draw.cuh: make_plan): one cell when all rows fit in shared
    memory, else full power-of-two tiles followed by one remainder cell."""
    if n_obs <= SMEM_DATA_BYTES // (8 * n_arrays):
        return [n_obs]
    m = 1 << log2_tile
    full = n_obs // m
    rem = n_obs - full * m
    return [m] * full + ([rem] if rem else [])


def device_indices(seed: int, resample: int, n_obs: int, log2_tile: int,
                   counts: Sequence[int] | None = None, cell_base: int = 0,
                   n_arrays: int = 1, class_counts=None) -> np.ndarray:
    """The canonical index array of resample `resample`: cells in order; a full tile
    lists its 16 bank classes in order, each class its draws in stream order; a
    general cell lists its Lemire draws in stream order.  `counts` are the tile
    occupancies and `class_counts[t][c]` the class split of full tile t (both come
    from the device multinomial chains); with a single cell the count is n_obs."""
    sizes = tile_plan(n_obs, log2_tile, n_arrays)
    if len(sizes) == 1:
        return cell_draws_lemire(seed, resample, cell_base, n_obs, n_obs)
    assert counts is not None and len(counts) == len(sizes) and sum(counts) == n_obs
    m = 1 << log2_tile
    parts = []
    for t, (sz, cnt) in enumerate(zip(sizes, counts)):
        if sz == m:
            cc = class_counts[t]
            assert sum(cc) == cnt
            for c in range(N_CLASSES):
                parts.append(class_draws(seed, resample, t, c, int(cc[c]), log2_tile) + t * m)
        else:
            parts.append(cell_draws_lemire(seed, resample, cell_base + t, cnt, sz) + t * m)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


# --------------------------------------------------------------------------
# Moving-block and subsample plans: reference restatements (PCG64) and the device
# streams (csrc/k1_resample.cuh: k2_moving_block_kernel, k2_shuffle_keys_kernel).
# --------------------------------------------------------------------------

PURPOSE_BLOCK = 3
PURPOSE_SHUFFLE = 4


def moving_block_indices_pcg64(n_obs: int, n_samples: int, block_length: int = 3, wrap: bool = False,
                               seed=None) -> Iterator[np.ndarray]:
    """bootstrap.py:772-787: ceil(N/L) block starts from rng.integers(0, last_block),
    expanded to start+0..L-1, cut to N (mod N when wrapping)."""
    rng = np.random.default_rng(seed)
    n_blocks = -(-n_obs // block_length)
    last_block = n_obs if wrap else n_obs - block_length
    steps = np.arange(block_length)
    for _ in range(n_samples):
        starts = rng.integers(0, last_block, size=n_blocks)
        idx = (starts[:, None] + steps[None, :]).ravel()[:n_obs]
        yield np.mod(idx, n_obs) if wrap else idx


def subsample_indices_pcg64(n_obs: int, n_samples: int, size: int, seed=None) -> np.ndarray:
    """bootstrap.py:741-744 for an absolute size: shuffle each row, keep the first `size`."""
    rng = np.random.default_rng(seed)
    base = np.tile(np.arange(n_obs), (n_samples, 1))
    for row in base:
        rng.shuffle(row)
    return base[:, :size]


def device_moving_block(seed: int, resample: int, n_obs: int, block_length: int, wrap: bool) -> np.ndarray:
    n_blocks = -(-n_obs // block_length)
    bound = n_obs if wrap else n_obs - block_length
    starts = cell_draws_lemire(seed, resample, 0, n_blocks, bound, purpose=PURPOSE_BLOCK)
    idx = (starts[:, None] + np.arange(block_length)[None, :]).ravel()[:n_obs]
    return np.mod(idx, n_obs) if wrap else idx


def device_subsample(seed: int, sample: int, n_obs: int, size: int) -> np.ndarray:
    """First `size` rows of the permutation that sorts the 63-bit Philox keys
    (ties broken by row number)."""
    nblk = (n_obs + 1) // 2
    w = philox4x32(_ctr(np.arange(nblk), 0, sample, PURPOSE_SHUFFLE), seed_key(seed)).astype(np.uint64)
    keys = np.stack(((w[:, 0] << np.uint64(32)) | w[:, 1], (w[:, 2] << np.uint64(32)) | w[:, 3]), axis=1)
    keys = (keys.reshape(-1)[:n_obs]) >> np.uint64(1)
    return np.lexsort((np.arange(n_obs), keys))[:size].astype(np.int64)


# --------------------------------------------------------------------------
# K5 restatement: exact k-th order statistic by an 8-pass MSB radix descent over
# order-preserving uint64 keys (csrc/k45_interval.cuh).  Equivalent to
# `stat.sort(axis=0); stat[k]` (bootstrap.py:445, 485-490); used to check the
# device select and to drive the sharded protocol on CPU (gloo) in the tests.
# --------------------------------------------------------------------------

def radix_key(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    b = x.view(np.uint64).copy()
    b[np.isnan(x)] = np.uint64(0x7FF8000000000000)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def radix_value(k: np.ndarray) -> np.ndarray:
    k = np.asarray(k, dtype=np.uint64)
    pos = (k >> np.uint64(63)).astype(bool)
    b = np.where(pos, k & np.uint64((1 << 63) - 1), ~k)
    return b.view(np.float64)


class RadixSelectState:
    """Host twin of the device select state for a single column."""

    def __init__(self, local_values: np.ndarray, ranks: Sequence[int]):
        self.keys = radix_key(local_values)
        self.prefix = np.zeros(len(ranks), dtype=np.uint64)
        self.krem = np.array(ranks, dtype=np.int64)
        self.hist = np.zeros((len(ranks), 256), dtype=np.int64)

    def do_hist(self, p: int) -> None:
        shift = np.uint64(56 - 8 * p)
        digit = ((self.keys >> shift) & np.uint64(0xFF)).astype(np.int64)
        for a in range(len(self.krem)):
            if p == 0:
                m = np.ones(len(self.keys), dtype=bool)
            else:
                m = ((self.keys ^ self.prefix[a]) >> (shift + np.uint64(8))) == 0
            self.hist[a] += np.bincount(digit[m], minlength=256)

    def do_pick(self, p: int) -> None:
        for a in range(len(self.krem)):
            cum = np.cumsum(self.hist[a])
            digit = int(np.searchsorted(cum, self.krem[a], side="right"))
            digit = min(digit, 255)
            below = cum[digit - 1] if digit > 0 else 0
            self.krem[a] -= below
            self.prefix[a] |= np.uint64(digit) << np.uint64(56 - 8 * p)
        self.hist[:] = 0

    def result(self) -> np.ndarray:
        return radix_value(self.prefix)
