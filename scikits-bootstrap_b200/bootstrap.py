"""Host-side mirror of the scikits.bootstrap API for the resampling hot path.

Same names, argument meaning and error behaviour as the reference
(/root/reference/src/scikits/bootstrap/bootstrap.py, cited as bootstrap.py:L):

    ci(data, statfunction, alpha, n_samples, method='pi'|'bca', output, epsilon,
       multi, return_dist, seed, use_numba)                       bootstrap.py:222-504
    pval(data, statfunction, compfunction, n_samples, multi, seed)  bootstrap.py:790-866
    bootstrap_indices / bootstrap_indices_independent             bootstrap.py:667-691
    jackknife_indices                                             bootstrap.py:694-705

What differs, by design:
  * the resample loop, jackknife, z0 count and order-statistic pick run as
    sm_100a kernels (libb200boot.so); there is no CPU path;
  * `statfunction` must be in the device registry (_registry.py) — anything
    else raises NotImplementedError, as the reference's use_numba switch raises
    for unsupported callables (bootstrap.py:426-427);
  * indices come from Philox4x32-10 keyed by (seed, resample id), not PCG64, so
    fixed-seed values differ from the reference while the distribution is the same.
"""
from __future__ import annotations

import math
import warnings
from collections.abc import Iterable
from typing import Any, Callable, Iterator, Optional, Sequence, Tuple, Union

import numpy as np

from . import _abi, _device, _dist, _interval
from ._registry import (DeviceCompare, DeviceStat, StatSpec, lookup, positive, require)

__all__ = (
    "ci",
    "ci_moving_block",
    "pval",
    "bootstrap_indices",
    "bootstrap_indices_independent",
    "subsample_indices",
    "jackknife_indices",
    "bootstrap_indices_moving_block",
)

__version__ = "0.1.0"


class InstabilityWarning(UserWarning):
    """Issued when results may be unstable."""


# Like the reference (bootstrap.py:130): never let these be filtered out.
warnings.simplefilter("always", InstabilityWarning)

SeedType = Union[None, int, np.ndarray, np.random.SeedSequence, np.random.BitGenerator,
                 np.random.Generator]

_MASK64 = (1 << 64) - 1
# sharded runs gather statistic vectors up to this size instead of reducing counts and histograms
_GATHER_LIMIT_BYTES = 32 << 20
# key offset between the arrays of multi='independent' (odd 64-bit constant)
_ARRAY_KEY_STRIDE = 0x9E3779B97F4A7C15
# K1mr: materialised index sets per batch (2 GB: at N = 1e6 that is 268 sets = two select CTAs per SM)
_ROWS_INDEX_BYTES = 2 << 30
# numpy rows from this size on are copied by the resampling call, under its occupancy chains
_DEFER_COPY_BYTES = 4 << 20


def _seed_key(seed: SeedType) -> int:
    """Maps every seed form numpy's default_rng accepts (bootstrap.py:145-152) to
    the 64-bit Philox key.  An int in [0, 2^64) is the key itself; a Generator /
    BitGenerator is advanced by one 64-bit draw (as a Generator passed to the
    reference is advanced by use)."""
    if seed is None:
        key = int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        return _dist.broadcast_seed(key)
    if isinstance(seed, np.random.Generator):
        return int(seed.integers(0, 1 << 64, dtype=np.uint64))
    if isinstance(seed, np.random.BitGenerator):
        return int(np.random.Generator(seed).integers(0, 1 << 64, dtype=np.uint64))
    if isinstance(seed, np.random.SeedSequence):
        return int(seed.generate_state(1, np.uint64)[0])
    if isinstance(seed, (int, np.integer)):
        s = int(seed)
        if s < 0:
            raise ValueError("expected non-negative integer")
        return s if s <= _MASK64 else int(np.random.SeedSequence(s).generate_state(1, np.uint64)[0])
    return int(np.random.SeedSequence(seed).generate_state(1, np.uint64)[0])


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _normalise(data, multi) -> Tuple[list, Any]:
    """multi validation and the data -> tuple-of-arrays step (bootstrap.py:362-379).
    Arrays stay as given (numpy / torch); only shapes are inspected here.  The reference
    copies its inputs at this point (np.array); here the copy that isolates the call from
    the caller's buffer is the host-to-device transfer itself, so no host copy is made."""
    if multi not in [False, True, None, "independent", "paired"]:
        raise ValueError("Value `{}` for multi is not recognized.".format(multi))
    if multi is None:
        multi = bool(isinstance(data, tuple))
    if multi and isinstance(data, Iterable):
        arrays = [x if _is_tensor(x) else np.asarray(x) for x in data]
        lengths = [int(x.shape[0]) for x in arrays]
        if len(set(lengths)) > 1 and multi != "independent":
            raise ValueError(
                "Data arrays have differing lengths {}, and ".format(lengths)
                + "multi is not set to 'independent'.")
    else:
        arrays = [data if _is_tensor(data) else np.asarray(data)]
    return arrays, multi


def _check_spec(spec: StatSpec, arrays: Sequence, multi) -> None:
    if multi == "independent" and not spec.independent:
        raise NotImplementedError(
            f"device statistic {spec.name} is not defined for multi='independent'")
    if spec.independent and multi != "independent":
        raise NotImplementedError(f"device statistic {spec.name} needs multi='independent'")
    if len(arrays) != spec.n_arrays:
        raise ValueError(f"device statistic {spec.name} takes {spec.n_arrays} array(s), "
                         f"got {len(arrays)}")
    for a in arrays:
        nd = len(a.shape)
        if spec.axis0:
            if nd not in (1, 2):
                raise NotImplementedError(f"{spec.name} needs 1-D or (N, D) data, got ndim={nd}")
        elif nd != 1:
            raise NotImplementedError(
                f"device statistic {spec.name} needs 1-D data (got ndim={nd}); use the *_axis0 "
                "registry statistics or functools.partial(np.<f>, axis=0) for (N, D) data")
        if a.shape[0] < 1:
            raise ValueError("data must have at least one row")


class _Run:
    """One resampling job on the current device: statistic vectors are kept
    column-major, float64[D][n_local]."""

    def __init__(self, spec: StatSpec, arrays: Sequence, multi, n_samples: int, key: int,
                 indices=None, blocks=None):
        self.spec, self.multi, self.n_samples, self.key = spec, multi, int(n_samples), key
        self.indices = indices       # parity mode: caller-supplied index sets (K1i)
        self.blocks = blocks         # moving-block plan: (block_length, wrap), else None
        self.rank, self.world = _dist.world()
        self.lo, self.hi = _dist.shard_range(self.n_samples, self.rank, self.world)
        self._host_rows = None
        if (self.world >= max(2, _device.SHARD_COPY_MIN_RANKS) and arrays[0].nbytes >= _device.SHARD_COPY_BYTES
                and self._split_copy(arrays)):
            pass                     # long host rows of a sharded run: 1/G per rank + all-gather over NVLink
        elif self._defer_copy(arrays):
            # long numpy rows of a plain K1 job: the copy (8 ms of host staging per 80 MB of pageable
            # memory) is issued by the resampling call itself, behind its occupancy chains
            torch = _device.require_cuda()
            self.dev = [torch.empty(a.shape, dtype=torch.float64, device="cuda") for a in arrays]
            self._host_rows = list(arrays)
            self._ready = None
        else:
            # pinned host tensors are copied on a side stream; `_ready` is recorded after the copies
            pairs = [_device.to_device_f64_async(a) for a in arrays]
            self.dev = [t for t, _ in pairs]
            events = [e for _, e in pairs if e is not None]
            self._ready = events[-1] if events else None      # one side stream: the last event covers all
        # the statistic has shape (D,): one value per data column, or a multi-output statistic
        self.vector = (spec.axis0 and self.dev[0].dim() == 2) or spec.n_out > 1
        self.n_cols = spec.n_out if spec.n_out > 1 else (self.dev[0].shape[1] if self.vector else 1)
        self._columns = None
        self._jack = None
        self._sorted_vals = None
        self._parts = None
        self._jack_dev = None
        self._rows = None            # K1mr tables: (sorted [D][N], rank [N][D])
        self._nan = None             # rank statistics: which columns hold NaNs
        self._local = False          # sharded run whose statistic vector was gathered: finish without collectives
        self._less = None            # small-call path: #{stat < ostat} and the speculative select
        self._spec = None
        # rank statistics: which order statistics (None: np.median's)
        self.rule = _device.rank_rule(self.dev[0].shape[0], spec.q) if spec.stat_id == _abi.STAT_MEDIAN else None

    def _split_copy(self, arrays) -> bool:
        """Every array a long contiguous 1-D float64 host array of a run sharded over NCCL:
        copied in parts and all-gathered (_device.shard_copy_async); `_ready` covers the gather."""
        if not all(type(a) is np.ndarray and a.ndim == 1 and a.dtype == np.float64 and a.flags.c_contiguous
                   and a.nbytes >= _device.SHARD_COPY_BYTES for a in arrays):
            return False
        pairs = [_device.shard_copy_async(a) for a in arrays]
        if any(p is None for p in pairs):          # (not an NCCL group: all of them are None)
            return False
        self.dev = [t for t, _ in pairs]
        self._ready = pairs[-1][1]                 # one copy stream: the last event covers all
        return True

    def _defer_copy(self, arrays) -> bool:
        """Host rows that can travel under K0: contiguous 1-D float64 numpy arrays of at least
        4 MB feeding the plain K1 path (scalar moment statistic, Philox draws, a non-empty shard) —
        the one path whose first device work does not read the data."""
        spec = self.spec
        return (self.indices is None and self.blocks is None and not spec.independent
                and spec.stat_id != _abi.STAT_MEDIAN and spec.n_out == 1 and self.hi > self.lo
                and all(type(a) is np.ndarray and a.ndim == 1 and a.dtype == np.float64
                        and a.flags.c_contiguous for a in arrays)
                and arrays[0].nbytes >= _DEFER_COPY_BYTES)

    # -- data views ------------------------------------------------------------
    def columns(self):
        """(N, D) data as D contiguous float64[N] rows (moment statistics per column)."""
        if self._columns is None:
            self._columns = self.dev[0].t().contiguous()
        return self._columns

    # -- long medians --------------------------------------------------------------------
    def _long_median(self) -> bool:
        """np.median / quantiles over more rows than K1m keeps resident."""
        return (self.spec.stat_id == _abi.STAT_MEDIAN
                and self.dev[0].shape[0] > _device.MEDIAN_RESIDENT_ROWS)

    def _row_space(self) -> bool:
        """Long rank statistics that must stay in ROW space (K1mr): (N, D) data with D > 1 —
        x[indices] gathers whole rows (bootstrap.py:431), so the columns of a resample share
        their rows — and index-supplied runs.  A single long column without supplied indices
        takes the rank-space path K1mt instead: its resample is sorted(x)[bootstrap_indices(...)],
        the same distribution for a permutation-invariant statistic at O(tiles) work per resample."""
        if not self._long_median():
            return False
        want = self.indices is not None or (self.vector and self.n_cols > 1)
        if want and self.dev[0].shape[0] > _device.MEDIAN_ROWS_MAX:
            if self.indices is not None:
                raise NotImplementedError("index-supplied median needs N <= {} rows".format(_device.MEDIAN_ROWS_MAX))
            return False          # beyond 2^24 rows: per-column rank space (marginals exact, rows not joint)
        return want

    def _rows_tables(self):
        """K1mr preparation, once per job: every column sorted, and each row's rank per column."""
        if self._rows is None:
            data2d = self.dev[0] if self.dev[0].dim() == 2 else self.dev[0].reshape(-1, 1)
            self._rows = _device.rank_columns(data2d.contiguous())
            self._sorted_vals = self._rows[0]
        return self._rows

    def _median_rows(self, index_batches, n_sets):
        torch = _device._torch()
        sorted_vals, rank = self._rows_tables()
        out = torch.empty((self.n_cols, n_sets), dtype=torch.float64, device=sorted_vals.device)
        at = 0
        for idx in index_batches:
            _device.median_rows_indexed(sorted_vals, rank, idx, out, at, rule=self.rule)
            at += idx.shape[0]
        return out

    def _sorted(self):
        """float64[D][N]: every column ascending."""
        if self._sorted_vals is None:
            if self.vector:
                s = self.dev[0].t().contiguous()
                if s.data_ptr() == self.dev[0].data_ptr():
                    s = s.clone()
            else:
                s = self.dev[0].reshape(1, -1).clone()
            _device.sort_columns(s)
            self._sorted_vals = s
        return self._sorted_vals

    # -- K1 ----------------------------------------------------------------------
    def resample(self):
        torch = _device._torch()
        spec, lo, hi = self.spec, self.lo, self.hi
        ready, self._ready = self._ready, None
        plain = (self.indices is None and not spec.independent and spec.stat_id != _abi.STAT_MEDIAN
                 and not self.vector and hi > lo)
        if ready is not None and not plain:
            torch.cuda.current_stream().wait_event(ready)     # only the plain K1 path overlaps the copy
            ready = None
        if self.indices is not None:
            return self._resample_indexed()
        if self.blocks is not None:
            if spec.independent or spec.stat_id == _abi.STAT_MEDIAN or self.vector and spec.n_out == 1:
                raise NotImplementedError("the moving-block plan covers moment statistics of 1-D (paired) arrays")
            flat = [a.reshape(-1) for a in self.dev]
            out = _device.resample_stat_moving_block(flat, spec.stat_id, self.key, self.blocks[0], self.blocks[1], lo, hi)
            return out.reshape(self.n_cols, -1)
        if spec.independent:
            a, b = (p.resample() for p in self.parts())
            return _device.combine(spec.op, a, b)
        if spec.stat_id == _abi.STAT_MEDIAN:
            return self._propagate_nan(self._resample_rank_statistic(lo, hi), lo, hi)
        if spec.n_out > 1:
            return _device.resample_stat([a.reshape(-1) for a in self.dev], spec.stat_id, self.key, lo, hi)
        if (self.vector and _device.gemm_pays(self.dev[0].shape[0], self.n_cols) and hi > lo
                and _device.gemm_supported(spec.stat_id, self.dev[0].shape[0])):
            # K6i: count matrix x data on the tensor cores (exact integer split); columns with
            # non-finite entries are redone the gather way inside the call, so no host check
            return _device.resample_stat_gemm(self.dev[0], spec.stat_id, self.key, lo, hi)
        if self.vector and _device.columns_block_width(spec.stat_id, self.dev[0].shape[0]) > 0:
            return _device.resample_stat_columns(self.dev[0], spec.stat_id, self.key, lo, hi)   # K1c
        if self.vector:
            out = torch.empty((self.n_cols, hi - lo), dtype=torch.float64, device=self.dev[0].device)
            cols = self.columns()
            for d in range(self.n_cols):
                _device.resample_stat([cols[d]], spec.stat_id, self.key, lo, hi, out=out[d])
            return out
        flat = [a.reshape(-1) for a in self.dev]
        host_rows, self._host_rows = self._host_rows, None
        return _device.resample_stat(flat, spec.stat_id, self.key, lo, hi, ready=ready,
                                     host_rows=host_rows).reshape(1, -1)

    def _resample_rank_statistic(self, lo, hi):
        """Median / quantile kernels (NaNs ordered last; _propagate_nan restores numpy's rule)."""
        torch = _device._torch()
        if self._row_space():
            return self._median_rows(self._index_batches(lo, hi), hi - lo)
        if self._long_median():
            cols = self._sorted()
            out = torch.empty((self.n_cols, hi - lo), dtype=torch.float64, device=cols.device)
            for d in range(self.n_cols):
                _device.resample_median_sorted(cols[d], self.key, lo, hi, out=out[d:d + 1], rule=self.rule)
            return out
        data2d = self.dev[0] if self.vector else self.dev[0].reshape(-1, 1)
        if hi > lo:
            # the columns are sorted for the resampling anyway: take the jackknife summary along
            out, self._jack_dev = _device.resample_median_columns(data2d, self.key, lo, hi, rule=self.rule,
                                                                  with_jackknife=True)
            return out
        return _device.resample_median_columns(data2d, self.key, lo, hi, rule=self.rule)

    def _index_batches(self, lo, hi):
        """The materialised draws of resamples [lo, hi) (K2), in batches of bounded size."""
        n_rows = self.dev[0].shape[0]
        step = max(1, min(hi - lo, _ROWS_INDEX_BYTES // (8 * n_rows)))
        return (_device.fill_indices(self.key, n_rows, 1, a, min(hi, a + step)) for a in range(lo, hi, step))

    def _nan_flags(self):
        """(device flags, host bool[D]) of the columns holding NaNs — rank statistics only: np.median
        and np.quantile of a sample with a NaN are NaN, the rank kernels order NaNs last."""
        if self._nan is None:
            data2d = self.dev[0] if self.dev[0].dim() == 2 else self.dev[0].reshape(-1, 1)
            self._nan = _device.nan_columns(data2d)
        return self._nan

    def _propagate_nan(self, out, lo, hi, index_batches=None):
        """numpy's NaN rule for rank statistics: a resample that draws a NaN row of column d has a
        NaN statistic there (statfunction(x[indices]), bootstrap.py:431).  Only columns with NaNs
        pay: their resamples' draws are replayed (K2) and looked up."""
        flags, host = self._nan_flags()
        if not host.any() or hi <= lo:
            return out
        data2d = self.dev[0] if self.dev[0].dim() == 2 else self.dev[0].reshape(-1, 1)
        n_rows = data2d.shape[0]
        rank_space = index_batches is None and self._long_median() and not self._row_space()
        at = 0
        for idx in (index_batches if index_batches is not None else self._index_batches(lo, hi)):
            for d in np.nonzero(host)[0].tolist():
                if rank_space:      # K1mt resamples sorted(x): the same draws address the sorted column
                    _device.nan_fixup(self._sorted()[d], 1, n_rows, idx, out[d, at:])
                else:
                    _device.nan_fixup(data2d[:, d], data2d.stride(0), n_rows, idx, out[d, at:])
            at += idx.shape[0]
        return out

    # -- short calls: everything in one library call ------------------------------------
    def small_call(self, alphas: np.ndarray, bca: bool, need_jack: bool):
        """cfg1-sized jobs (one resident cell, scalar moment statistic, short statistic vector,
        not sharded, Philox draws): K1, the jackknife and a single-CTA tail (z0 count, levels,
        ranks, select) are enqueued by one call and one 256-byte copy comes back.  Returns the
        statistic vector float64[1][n] or None when the job does not qualify.  The ranks the
        device used are speculative; `select` compares them with the host's."""
        spec = self.spec
        if (self.indices is not None or self.blocks is not None or spec.independent or self.vector or spec.n_out != 1
                or _dist.active() or spec.stat_id == _abi.STAT_MEDIAN or self.dev[0].dim() != 1):
            return None
        n_alphas = int(np.prod(alphas.shape)) if alphas.ndim else 1
        if not _device.ci_small_supported(spec.stat_id, len(self.dev), self.dev[0].shape[0], self.n_samples, n_alphas):
            return None
        torch = _device._torch()
        ready, self._ready = self._ready, None
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
        stat, vals, ranks, less, jack = _device.ci_small([a.reshape(-1) for a in self.dev], spec.stat_id, self.key,
                                                         self.n_samples, alphas, bca, need_jack)
        if need_jack:
            self._jack = jack.reshape(1, 8)
            self._less = np.array([less], dtype=np.int64)
        self._spec = (ranks.reshape(-1, 1), vals.reshape(-1, 1))
        return stat.reshape(1, -1)

    def tail_call(self, stat, alphas: np.ndarray, bca: bool, need_jack: bool) -> None:
        """Scalar statistics with a moderately long vector (cfg5: 1e5 values; a sharded run's
        gathered vector): the jackknife summary stays on the device and ONE single-CTA launch
        counts stat < ostat, forms speculative levels and ranks and selects the order
        statistics; one 256-byte block comes back (instead of count -> host -> 8 radix passes ->
        host).  `select` verifies the ranks against the host's, as for the small call."""
        if (self.n_cols != 1 or self.vector or (_dist.active() and not self._local) or self._spec is not None
                or stat.shape[0] != 1):
            return
        n_alphas = int(np.prod(alphas.shape)) if alphas.ndim else 1
        if not _device.ci_tail_supported(stat.shape[1], n_alphas):
            return
        jd = None
        if need_jack:
            if self._jack is not None:
                return                                   # already on the host: the general tail is as short
            jd = self._jackknife_tensor().reshape(-1)
        vals, ranks, less, jack = _device.ci_tail(stat.reshape(-1), jd, alphas, bca)
        if need_jack:
            self._jack = jack.reshape(1, 8)
            self._less = np.array([less], dtype=np.int64)
        self._spec = (ranks.reshape(-1, 1), vals.reshape(-1, 1))

    # -- K1i ---------------------------------------------------------------------
    def _resample_indexed(self):
        """Parity mode: gathers with the caller's index sets (e.g. the reference's own
        PCG64 draws) instead of drawing; moment statistics of 1-D arrays only."""
        torch = _device._torch()
        spec = self.spec
        if _dist.active() or spec.independent:
            raise NotImplementedError("index-supplied mode covers paired (non-sharded) statistics")
        n_rows = self.dev[0].shape[0]
        idx = torch.as_tensor(np.ascontiguousarray(self.indices, dtype=np.int64)).to(self.dev[0].device)
        if idx.dim() != 2 or idx.shape[0] != self.n_samples or idx.shape[1] != n_rows:
            raise ValueError("indices must have shape (n_samples, N)")
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n_rows):
            raise IndexError("index out of range")
        if spec.stat_id == _abi.STAT_MEDIAN:
            if self._row_space():
                out = self._median_rows([idx], idx.shape[0])
            else:
                data2d = self.dev[0] if self.vector else self.dev[0].reshape(-1, 1)
                out = _device.resample_median_columns_indexed(data2d, idx, rule=self.rule)
            return self._propagate_nan(out, 0, idx.shape[0], index_batches=[idx])
        if self.vector and spec.n_out == 1:
            cols = self.columns()
            return torch.stack([_device.resample_stat_indexed([cols[d]], spec.stat_id, idx)
                                for d in range(self.n_cols)])
        flat = [a.reshape(-1) for a in self.dev]
        return _device.resample_stat_indexed(flat, spec.stat_id, idx).reshape(self.n_cols, -1)

    # -- K3 ----------------------------------------------------------------------
    def jackknife(self):
        """float64[D][8] on the host: ostat, jmean, d2, d3, accel, ..."""
        if self._jack is None:
            self._jack = _device.to_host(self._jackknife_tensor())
        return self._jack

    def jackknife_and_less(self, stat):
        """BCa inputs: the jackknife summary (host, float64[D][8]) and #{stat < ostat} per
        column (bootstrap.py:583).  For a scalar statistic the count kernel reads ostat
        straight from the jackknife output on the device, so the two results come back
        with one wait instead of a host round trip in between."""
        if self._less is not None:                                 # small-call path: already on the host
            return self._jack, self._less
        if self.n_cols == 1 and not self.spec.independent and self._jack is None:
            jd = self._jackknife_tensor()
            c = _device.count(stat, _abi.CMP_LT, jd)              # &jd[0][0] is ostat
            self._reduce(c)
            self._jack = _device.to_host(jd)
            return self._jack, _device.to_host(c)
        jack = self.jackknife()
        return jack, self.count(stat, _abi.CMP_LT, jack[:, 0])

    def _jackknife_tensor(self):
        torch = _device._torch()
        spec = self.spec
        if spec.independent:
            j = self._jackknife_independent()
            for p in self.parts():        # a rank statistic of an array with a NaN: every pooled value is NaN
                if p.spec.stat_id == _abi.STAT_MEDIAN and p._nan_flags()[1].any():
                    _device.nan_poison_rows(j, p._nan_flags()[0])
        elif spec.stat_id == _abi.STAT_MEDIAN:
            if self._long_median():
                j = torch.cat([_device.jackknife_median_sorted(c, rule=self.rule) for c in self._sorted()])
            elif self._jack_dev is not None:
                j = self._jack_dev
            else:
                data2d = self.dev[0] if self.vector else self.dev[0].reshape(-1, 1)
                j = _device.jackknife_median_columns(data2d, rule=self.rule)
            flags, host = self._nan_flags()
            if host.any():                # statfunction(data) and the leave-one-out values are NaN (numpy's rule)
                _device.nan_poison_rows(j, flags)
        elif self.vector and spec.n_out == 1:
            j = _device.jackknife_columns(self.dev[0], spec.stat_id)                               # K3c
        else:
            j = _device.jackknife([a.reshape(-1) for a in self.dev], spec.stat_id).reshape(self.n_cols, 8)
        return j

    def parts(self):
        """multi='independent': one single-array job per data array, each with its own Philox
        key (its index stream is independent of the others: bootstrap.py:683-691)."""
        if self._parts is None:
            self._parts = [_Run(self.spec.inner[k], [a], False, self.n_samples,
                                (self.key + k * _ARRAY_KEY_STRIDE) & _MASK64)
                           for k, a in enumerate(self.dev)]
        return self._parts

    def loo_values(self):
        """Single-array job: the leave-one-out statistic of every row and the K3 summary,
        both on the device (order of the values is irrelevant to their pooled moments)."""
        if self.spec.stat_id == _abi.STAT_MEDIAN:
            return _device.jackknife_values_sorted(self._sorted()[0], rule=self.rule)
        return _device.jackknife_values([self.dev[0].reshape(-1)], self.spec.stat_id)

    def _jackknife_independent(self):
        """f(s_a, s_b) under multi='independent': sum_k N_k leave-one-out values, array-major
        (bootstrap.py:601-607, 708-714).  Deleting row j of array k moves only s_k, so the values
        are f(loo_a[j], s_b) and f(s_a, loo_b[j]) with the per-array leave-one-out vectors of K3 /
        K3m; their mean and the sums of squared and cubed deviations are reduced on the device."""
        (va, ja), (vb, jb) = (p.loo_values() for p in self.parts())
        return _device.jackknife_pooled(self.spec.op, va, ja[0:1], vb, jb[0:1])

    # -- sharded runs with a small statistic vector ---------------------------------
    def gather_small(self, stat):
        """Every rank takes the whole statistic vector (one all-gather) when it is small and
        finishes on its own: one collective instead of an all-reduce per count and per radix
        pass.  The gathered vector is ordered by resample id, i.e. it IS the unsharded run's."""
        if (_dist.active() and not self._local
                and self.n_samples * self.n_cols * 8 <= _GATHER_LIMIT_BYTES):
            self._local = True
            return _dist.all_gather_ragged(stat, self.n_samples).contiguous()
        return stat

    def _reduce(self, c):
        if not self._local:
            _dist.all_reduce_sum(c)

    # -- K4 ----------------------------------------------------------------------
    def count(self, stat, op: int, t0: np.ndarray, t1: Optional[np.ndarray] = None) -> np.ndarray:
        torch = _device._torch()
        t0d = _device.host_to_device(np.asarray(t0, dtype=np.float64), stat.device)
        t1d = None if t1 is None else _device.host_to_device(np.asarray(t1, dtype=np.float64), stat.device)
        c = _device.count(stat, op, t0d, t1d)
        self._reduce(c)
        return _device.to_host(c)

    # -- K5 ----------------------------------------------------------------------
    def select(self, stat, ranks: np.ndarray) -> np.ndarray:
        """ranks int[A][D] (global, 0-based) -> float64[A][D] order statistics."""
        torch = _device._torch()
        if self._spec is not None:
            spec_ranks, spec_vals = self._spec
            self._spec = None
            if spec_ranks.shape == np.shape(ranks) and np.array_equal(spec_ranks, ranks):
                return spec_vals                                  # the device's ranks were the host's
        r = _device.host_to_device(np.asarray(ranks, dtype=np.int64), stat.device)
        sel = _device.Select(stat, r)
        out = sel.run(sharded=_dist.active() and not self._local)
        return _device.to_host(out)

    # -- full sorted distribution (return_dist / host-side compfunctions) --------
    def sorted_dist(self, stat) -> np.ndarray:
        full = stat.clone() if self._local else _dist.all_gather_ragged(stat, self.n_samples).contiguous()
        _device.sort_columns(full)
        host = _device.to_host(full)
        return host.T.copy() if self.vector else host[0].copy()


def ci(
    data,
    statfunction: Optional[Callable] = None,
    alpha: Union[float, Iterable] = 0.05,
    n_samples: int = 10000,
    method: str = "bca",
    output: str = "lowhigh",
    epsilon: float = 0.001,
    multi=None,
    return_dist: bool = False,
    seed: SeedType = None,
    use_numba: bool = False,
):
    """Bootstrap confidence interval of `statfunction` on `data`; drop-in for the
    reference's `ci` with method 'pi' or 'bca' (bootstrap.py:222-504).

    data: array_like (N, ...) or tuple of array_likes; numpy arrays, sequences,
        pandas Series and torch tensors (CPU or CUDA; float64 CUDA tensors are used
        in place).  statfunction: a callable from the device registry (see
        `scikits_bootstrap_b200.REGISTRY_NAMES`); default np.average.
    Returns the interval endpoints, and the sorted bootstrap distribution as well
    when return_dist=True."""
    if (multi is None and not return_dist and not use_numba
            and (type(data) is np.ndarray or type(data) is tuple or _is_tensor(data))):
        res = _ci_small_direct(data, statfunction, alpha, n_samples, method, output, seed)
        if type(res) is _RetryWithKey:        # the general path, with the key the seed was already turned into
            seed = res.key
        elif res is not None:
            return res
    return _ci_impl(data, statfunction, alpha, n_samples, method, output, epsilon, multi,
                    return_dist, seed, use_numba, None)


class _RetryWithKey:
    """The direct path found something out of its ordinary (NaNs under a rank statistic) after
    the seed had been consumed: the general path runs with that key."""

    def __init__(self, key):
        self.key = key


class _SmallRun:
    """What _ci_tail needs from a job whose device work is already done (the fused small call)."""
    vector = False
    n_cols = 1

    def __init__(self, vals, ranks, less, jack):
        self._spec = (ranks.reshape(-1, 1), vals.reshape(-1, 1))
        self._jack = jack.reshape(1, 8)
        self._less = np.array([less], dtype=np.int64)

    def jackknife_and_less(self, stat):
        return self._jack, self._less

    def jackknife(self):
        return self._jack

    def select(self, stat, ranks):
        spec_ranks, spec_vals = self._spec
        if spec_ranks.shape == np.shape(ranks) and np.array_equal(spec_ranks, ranks):
            return spec_vals                                      # the device's ranks were the host's
        stat = stat.reshape(1, -1)
        r = _device.host_to_device(np.asarray(ranks, dtype=np.int64), stat.device)
        return _device.to_host(_device.Select(stat, r).run(sharded=False))


def _ci_small_direct(data, statfunction, alpha, n_samples, method, output, seed):
    """Short calls (cfg1: N = 1e3, n_samples = 1e4) are bound by per-call overheads, so the
    common form — one contiguous 1-D float64 numpy array or CUDA tensor, a scalar moment
    statistic, a float alpha — skips the general job set-up: host rows go to the device inside
    the one library call that also runs the fused kernel (b200boot_ci_small_host), and the
    interval follows from the returned block in Python floats (_small_tail; the general tail
    whenever anything is out of the ordinary).  Returns None when the call is not of that
    form; the result is what the general path returns, bit for bit."""
    if (type(alpha) is not float or type(n_samples) is not int or method not in ("pi", "bca")
            or output not in ("lowhigh", "errorbar") or _dist.active()):
        return None
    if type(data) is tuple:                      # paired arrays (multi=None on a tuple means paired)
        if not (len(data) == 2 and all(type(a) is np.ndarray and a.ndim == 1 and a.dtype == np.float64
                                       and a.flags.c_contiguous for a in data)
                and data[0].shape[0] == data[1].shape[0]):
            return None
        arrays, n_rows = data, data[0].shape[0]
    elif type(data) is np.ndarray:
        if data.ndim != 1 or data.dtype != np.float64 or not data.flags.c_contiguous:
            return None
        arrays, n_rows = (data,), data.shape[0]
    else:
        torch = _device.require_cuda()
        if (not data.is_cuda or data.dtype != torch.float64 or data.dim() != 1 or not data.is_contiguous()
                or data.device.index != torch.cuda.current_device()):
            return None
        arrays, n_rows = (data,), data.shape[0]
    spec = lookup(np.average if statfunction is None else statfunction)
    if (spec is None or spec.n_arrays != len(arrays) or spec.independent or spec.n_out != 1 or n_rows < 1):
        return None
    _device.require_cuda()
    bca = method == "bca"
    if spec.stat_id == _abi.STAT_MEDIAN:
        # rank statistics of one resident column: sort, jackknife, resamples (K1m1) and tail in one
        # library call; data holding NaNs go to the general path (numpy's NaN rule)
        if not _device.ci_median_small_supported(n_rows, n_samples, 2):
            return None
        key = _seed_key(seed)
        alphas = np.array([alpha / 2, 1 - alpha / 2])
        stat, vals, ranks, less, jack, has_nan = _device.ci_median_small(
            arrays[0], key, n_samples, alphas, bca, bca or output == "errorbar", rule=_device.rank_rule(n_rows, spec.q))
        if has_nan:
            return _RetryWithKey(key)
    else:
        if not _device.ci_small_supported(spec.stat_id, len(arrays), n_rows, n_samples, 2):
            return None
        key = _seed_key(seed)
        alphas = np.array([alpha / 2, 1 - alpha / 2])
        stat, vals, ranks, less, jack = _device.ci_small(arrays, spec.stat_id, key, n_samples, alphas, bca,
                                                         bca or output == "errorbar")
    out = _small_tail(vals, ranks, less, jack, alphas, n_samples, bca, output)
    if out is not None:
        return out
    return _ci_tail(_SmallRun(vals, ranks, less, jack), stat, alphas, n_samples, method, output, False, 4)


def _small_tail(vals, ranks, less, jack, alphas, n_samples, bca, output):
    """_ci_tail for a scalar statistic and two levels when nothing is out of the ordinary: finite
    acceleration, 0 < less < n_samples, finite levels, and the device's speculative ranks equal
    to the host's.  Python floats, the same IEEE operations in the same order as the array forms
    (_interval._bca_levels_scalar and order_indices' per-element path, which are the ones
    _ci_tail takes for such input).  Anything else returns None and the general tail decides."""
    levels = alphas.tolist()
    if bca:
        accel, c = float(jack[4]), float(less)
        if not (math.isfinite(accel) and 0.0 < c < n_samples):
            return None
        try:
            levels = _interval._bca_levels_scalar(c, n_samples, accel, alphas).tolist()
        except (ZeroDivisionError, OverflowError, ValueError):
            return None
    top = n_samples - 1
    extremal = top10 = False
    for a, r in zip(levels, ranks.tolist()):
        v = top * a
        if v != v or abs(v) >= 4.0e18:
            return None
        k = round(v)
        if k == 0 or k == top:
            extremal = True
        elif k < 10 or k >= n_samples - 10:
            top10 = True
        if min(max(k, 0), top) != r:
            return None
    if extremal:
        warnings.warn("Some values used extremal samples; " + "results are probably unstable.",
                      InstabilityWarning, stacklevel=4)
    elif top10:
        warnings.warn("Some values used top 10 low/high samples; " + "results may be unstable.",
                      InstabilityWarning, stacklevel=4)
    if output == "lowhigh":
        return vals
    return np.abs(jack[0] - vals)[np.newaxis].T                   # bootstrap.py:491-497


def _ci_abc(arrays, statfunction, alphas, output, epsilon):
    """ABC interval (bootstrap.py:507-567) for the statistic the reference's ABC supports among
    the registry: the weighted mean, np.average (the reference calls
    `statfunction(x, weights=...)`; np.mean & co. fail its weights probe with the TypeError
    reproduced below).

    The reference takes finite differences, with step epsilon / N, of 2N + 3 + A weighted means;
    at the default epsilon its second differences are dominated by the rounding of those means,
    so the answer is defined by numpy's arithmetic.  The device evaluates every weighted mean
    with that arithmetic (single IEEE operations, numpy's pairwise summation order:
    csrc/abc_method.cuh); what follows are the reference's own O(N) numpy expressions, so the
    interval agrees with the reference to rounding of `pow` (about 1e-15) for any epsilon."""
    if statfunction is not np.average:
        if lookup(statfunction) is not None or statfunction is np.mean:
            raise TypeError("statfunction does not accept correct arguments for ABC")
        raise NotImplementedError(
            "method='abc' is available for the weighted mean (np.average) only; "
            "statfunction {!r} is not in the device registry".format(statfunction))
    if len(arrays) != 1 or len(arrays[0].shape) != 1:
        raise NotImplementedError("method='abc' needs a single 1-D array")
    with _device.device_context(arrays):
        x = _device.to_device_f64(arrays[0]).reshape(-1)
        nn = int(x.numel())
        n = nn * 1.0
        ep = epsilon / n * 1.0
        p0 = np.repeat(1.0 / n, nn)
        tp, tm = _device.abc_identity_averages(x, ep)
        t0 = _device.weighted_averages(x, p0[np.newaxis])[0]
        t1 = (tp - tm) / (2 * ep)
        t2 = (tp - 2 * t0 + tm) / ep**2
        sighat = np.sqrt(np.sum(t1**2)) / n
        with np.errstate(invalid="ignore", divide="ignore"):
            a = (np.sum(t1**3)) / (6 * n**3 * sighat**3)
            delta = t1 / (n**2 * sighat)
            cqp, cqm = _device.weighted_averages(x, np.stack([p0 + ep * delta, p0 - ep * delta]))
            cq = (cqp - 2 * t0 + cqm) / (2 * sighat * ep**2)
            bhat = np.sum(t2) / (2 * n**2)
            curv = bhat / sighat - cq
            z0 = _interval.nppf(2 * _interval.ncdf(a) * _interval.ncdf(-curv))
            big_z = z0 + _interval.nppf(alphas)
            za = big_z / (1 - a * big_z) ** 2
            flat = np.atleast_1d(za).reshape(-1)
            abc = _device.weighted_averages(x, np.stack([p0 + z * delta for z in flat])).reshape(np.shape(za))
        if output == "lowhigh":
            return abc
        full = _device.weighted_averages(x, None)[0]                           # statfunction(*tdata)
        return np.abs(abc - full)[np.newaxis].T


def ci_moving_block(data, statfunction=None, alpha=0.05, n_samples=10000, block_length=3, wrap=False,
                    method="pi", output="lowhigh", multi=None, return_dist=False, seed=None):
    """`ci` with moving-block resamples (for dependent data): every resample is what
    `bootstrap_indices_moving_block(data, n_samples, block_length, wrap, seed)` yields
    (bootstrap.py:747-787), drawn, gathered and reduced in one fused kernel (K1b) without
    materialising the indices.  The reference offers the index generator only; looping
    `statfunction(data[idx])` over it gives the same statistics.  Moment statistics of 1-D data
    or paired 1-D arrays; method 'pi' (default) or 'bca' (the jackknife acceleration treats the
    rows as exchangeable, as the reference's `ci` would)."""
    n_rows = _n_rows(data[0] if isinstance(data, tuple) else data)
    block_length = int(block_length)
    if block_length < 1 or (n_rows if wrap else n_rows - block_length) < 1:
        raise ValueError("low >= high")          # numpy's error for an empty start range (:783)
    return _ci_impl(data, statfunction, alpha, n_samples, method, output, 0.001, multi, return_dist,
                    seed, False, None, (block_length, bool(wrap)))


def ci_indexed(data, statfunction=None, indices=None, alpha=0.05, method="bca", output="lowhigh",
               multi=None, return_dist=False):
    """Parity mode of `ci`: identical pipeline, but the resamples are the caller's
    `indices` (int array, shape (n_samples, N)) instead of Philox draws.  Feeding the
    reference's own `bootstrap_indices` output reproduces the reference's result."""
    idx = np.asarray(indices if not hasattr(indices, "__next__") else list(indices))
    return _ci_impl(data, statfunction, alpha, idx.shape[0], method, output, 0.001, multi,
                    return_dist, 0, False, idx)


def _ci_impl(data, statfunction, alpha, n_samples, method, output, epsilon, multi, return_dist,
             seed, use_numba, indices, blocks=None):
    """Body shared by ci and ci_indexed (warnings use stacklevel=3: the user's frame)."""
    # alphas exactly as given, or (alpha/2, 1-alpha/2)             bootstrap.py:353-356
    if isinstance(alpha, Iterable):
        alphas = np.array(alpha)
    else:
        alphas = np.array([alpha / 2, 1 - alpha / 2])

    arrays, multi = _normalise(data, multi)

    if statfunction is None:
        statfunction = np.average
    elif not callable(statfunction):
        raise TypeError(
            "statfunction {} is not callable.  If you tried ".format(statfunction)
            + "calling a function with arguments here, for example, by using "
            + "statfunction=myfunction(data, arg), then you probably need "
            + "to wrap your function in a lambda, eg, as "
            + "statfunction=(lambda data: myfunction(data, arg)). "
            + "If your function doesn't need any arguments other than the data, "
            + "you can alternatively use statfunction=myfunction (without "
            + "parentheses.")

    if method == "abc":                                           # bootstrap.py:400-413
        if return_dist:
            raise ValueError(
                "The ABC method is being used, but return_dist=True. The distribution cannot be"
                + "returned in this case, because the ABC method doesn't actually calculate it.")
        if multi == "independent":
            raise NotImplementedError("multi='independent' is not currently supported for ABC.")
        if output not in ("lowhigh", "errorbar"):
            raise ValueError("Output option {0} is not supported.".format(output))
        return _ci_abc(arrays, statfunction, alphas, output, epsilon)
    if use_numba:                                                 # bootstrap.py:415-437
        if multi == "independent":
            raise NotImplementedError(
                "Numba for independent data is not implemented, because the numba code does not support >1d data yet")
        if not (statfunction in (np.average, np.mean) and len(arrays) == 1):
            raise ValueError("Numba can't be used with these values currently.")
        # accepted for signature parity: the compiled path here is the device path
    if method not in ("pi", "bca"):
        raise ValueError("Method {0} is not supported.".format(method))
    if output not in ("lowhigh", "errorbar"):
        raise ValueError("Output option {0} is not supported.".format(output))

    spec = require(statfunction)
    _check_spec(spec, arrays, multi)
    n_samples = int(n_samples)
    if n_samples < 1:
        raise ValueError("n_samples must be positive")
    key = _seed_key(seed)

    with _device.device_context(arrays):
        res = _ci_on_device(spec, arrays, multi, n_samples, key, indices, alphas, method, output,
                            return_dist, blocks)
    # the reference's results carry the floating dtype of the data (np.mean of float32 is float32);
    # the device computes in float64 and the result is rounded to that dtype once, at the end
    low = _low_precision_dtype(arrays)
    if low is None:
        return res
    if return_dist:
        return res[0].astype(low), res[1].astype(low)
    return res.astype(low)


def _low_precision_dtype(arrays):
    """np.float32 / np.float16 when every data array has that floating dtype, else None."""
    kinds = set()
    for a in arrays:
        name = str(a.dtype).replace("torch.", "")
        kinds.add(name)
    if len(kinds) == 1:
        name = kinds.pop()
        if name in ("float32", "float16"):
            return np.dtype(name)
    return None


def _ci_on_device(spec, arrays, multi, n_samples, key, indices, alphas, method, output, return_dist,
                  blocks=None):
    """Device part of ci (warnings use stacklevel=5: the user's frame)."""
    run = _Run(spec, arrays, multi, n_samples, key, indices, blocks)
    stat = None
    if not return_dist:
        stat = run.small_call(alphas, method == "bca", method == "bca" or output == "errorbar")
    if stat is None:
        stat = run.gather_small(run.resample())                   # float64[D][n_local]
        if not return_dist:
            run.tail_call(stat, alphas, method == "bca", method == "bca" or output == "errorbar")
    return _ci_tail(run, stat, alphas, n_samples, method, output, return_dist, 5)


def _ci_tail(run, stat, alphas, n_samples, method, output, return_dist, stacklevel):
    """From the statistic vector to the interval: bootstrap.py:445-504.  `stacklevel` reaches
    the user's frame from here."""
    ostat = None
    if method == "pi":
        avals = alphas
    else:
        jack, less = run.jackknife_and_less(stat)                 # strict '<'  bootstrap.py:583
        ostat = jack[:, 0]
        accel = jack[:, 4]
        if not run.vector:
            less, accel_s = less[0], accel[0]
        else:
            accel_s = accel
        if np.any(np.isnan(accel_s)):                             # bootstrap.py:616-625
            nanind = np.nonzero(np.isnan(np.atleast_1d(accel_s)))
            warnings.warn(
                "BCa acceleration values for indices {} were undefined. \
Statistic values were likely all equal. Affected CI will \
be inaccurate.".format(nanind),
                InstabilityWarning,
                stacklevel=stacklevel,
            )
        avals = _interval.bca_levels(less, n_samples, accel_s, alphas)

    nvals, flags = _interval.order_indices(avals, n_samples)      # bootstrap.py:456-480
    if "nan" in flags:
        warnings.warn(
            "Some values were NaN; results are probably unstable "
            + "(all values were probably equal)",
            InstabilityWarning,
            stacklevel=stacklevel,
        )
    if "extremal" in flags:
        warnings.warn(
            "Some values used extremal samples; " + "results are probably unstable.",
            InstabilityWarning,
            stacklevel=stacklevel,
        )
    elif "top10" in flags:
        warnings.warn(
            "Some values used top 10 low/high samples; " + "results may be unstable.",
            InstabilityWarning,
            stacklevel=stacklevel,
        )

    # order statistics without sorting: ranks int[A][D]            bootstrap.py:482-490
    n_alphas = int(np.prod(alphas.shape)) if alphas.ndim else 1
    flat = nvals.reshape(n_alphas, -1)
    ranks = np.broadcast_to(flat, (n_alphas, run.n_cols)) if flat.shape[1] == 1 else flat
    ranks = np.clip(ranks, 0, n_samples - 1)
    picked = run.select(stat, ranks)                              # float64[A][D]

    if run.vector:
        vals = picked.reshape(alphas.shape + (run.n_cols,))
    else:
        vals = picked[:, 0].reshape(alphas.shape)

    if output == "lowhigh":
        out = vals
    else:                                                         # bootstrap.py:491-497
        if ostat is None:
            ostat = run.jackknife()[:, 0]
        o = ostat if run.vector else ostat[0]
        if nvals.ndim == 1:
            out = np.abs(o - vals)[np.newaxis].T
        else:
            out = np.abs(o - vals).T.squeeze()

    if return_dist:
        return out, run.sorted_dist(stat)
    return out


def pval(
    data,
    statfunction: Callable = np.average,
    compfunction: Callable[[Any], Any] = positive,
    n_samples: int = 10000,
    multi: Optional[bool] = None,
    seed: SeedType = None,
):
    """Bootstrap probability that `compfunction(statfunction(resample))` holds;
    drop-in for the reference's `pval` (bootstrap.py:790-866).  The default
    criterion is `s > 0`.  Criteria built with less_than / less_equal /
    greater_than / greater_equal / between are counted on the device; any other
    callable is applied on the host to the device-computed statistic values, row
    by row, as the reference does (bootstrap.py:861).  Beyond the reference (whose
    pval treats any truthy `multi` as paired, bootstrap.py:841-850): multi='independent'
    resamples each array on its own, as `ci` does."""
    if multi is None and isinstance(compfunction, DeviceCompare):
        res = _pval_small_direct(data, statfunction, compfunction, n_samples, seed)
        if res is not None:
            return res
    return _pval_impl(data, statfunction, compfunction, n_samples, multi, seed, None)


def _pval_small_direct(data, statfunction, compfunction, n_samples, seed):
    """Short pval() calls on one contiguous 1-D float64 array (numpy or CUDA) or a pair of such
    numpy arrays with a device criterion: one kernel (the fused small call counting the criterion
    instead of selecting order statistics).  None when the call is not of that form; the value is
    the general path's."""
    if type(n_samples) is not int or _dist.active():
        return None
    if type(data) is tuple:
        if not (len(data) == 2 and all(type(a) is np.ndarray and a.ndim == 1 and a.dtype == np.float64
                                       and a.flags.c_contiguous for a in data)
                and data[0].shape[0] == data[1].shape[0]):
            return None
        arrays, n_rows = data, data[0].shape[0]
    elif type(data) is np.ndarray:
        if data.ndim != 1 or data.dtype != np.float64 or not data.flags.c_contiguous:
            return None
        arrays, n_rows = (data,), data.shape[0]
    elif _is_tensor(data):
        torch = _device.require_cuda()
        if (not data.is_cuda or data.dtype != torch.float64 or data.dim() != 1 or not data.is_contiguous()
                or data.device.index != torch.cuda.current_device()):
            return None
        arrays, n_rows = (data,), data.shape[0]
    else:
        return None
    if not callable(statfunction):
        return None
    spec = lookup(statfunction)
    if (spec is None or spec.n_arrays != len(arrays) or spec.independent or spec.n_out != 1
            or spec.stat_id == _abi.STAT_MEDIAN or n_rows < 1):
        return None
    _device.require_cuda()
    if not _device.ci_small_supported(spec.stat_id, len(arrays), n_rows, n_samples, 1):
        return None
    key = _seed_key(seed)
    hits = _device.pval_small(arrays, spec.stat_id, key, n_samples, compfunction.op, compfunction.t0, compfunction.t1)
    return np.float64(hits / float(n_samples))


def pval_indexed(data, statfunction=np.average, compfunction=positive, indices=None, multi=None):
    """Parity mode of `pval`: resamples are the caller's `indices` (n_samples, N)."""
    idx = np.asarray(indices if not hasattr(indices, "__next__") else list(indices))
    return _pval_impl(data, statfunction, compfunction, idx.shape[0], multi, 0, idx)


def _pval_impl(data, statfunction, compfunction, n_samples, multi, seed, indices):
    if multi is None:
        multi = bool(isinstance(data, tuple))
    arrays, multi = _normalise(data, multi if multi == "independent" else bool(multi))
    if not callable(statfunction):
        raise TypeError("statfunction {} is not callable.".format(statfunction))
    spec = require(statfunction)
    _check_spec(spec, arrays, multi)
    n_samples = int(n_samples)
    if n_samples < 1:
        raise ValueError("n_samples must be positive")
    key = _seed_key(seed)
    with _device.device_context(arrays):
        return _pval_on_device(spec, arrays, multi, n_samples, key, indices, compfunction)


def _pval_on_device(spec, arrays, multi, n_samples, key, indices, compfunction):
    run = _Run(spec, arrays, multi, n_samples, key, indices)
    stat = run.gather_small(run.resample())
    if isinstance(compfunction, DeviceCompare):
        t0 = np.full(run.n_cols, compfunction.t0)
        t1 = np.full(run.n_cols, compfunction.t1)
        hits = run.count(stat, compfunction.op, t0, t1)
        frac = hits / float(n_samples)
        return frac if run.vector else np.float64(frac[0])
    dist = run.sorted_dist(stat)
    return np.mean([compfunction(s) for s in dist], axis=0)


def _n_rows(data) -> int:
    if hasattr(data, "shape"):
        return int(data.shape[0])
    return len(data)


def _index_batches(key: int, n_rows: int, n_arrays: int, n_samples: int):
    if n_rows < 1:
        raise ValueError("low >= high")              # numpy's message for rng.integers(0, 0, ...)
    budget = 64 << 20
    step = max(1, min(n_samples, budget // (8 * n_rows)))
    for lo in range(0, n_samples, step):
        hi = min(n_samples, lo + step)
        yield _device.to_host(_device.fill_indices(key, n_rows, n_arrays, lo, hi))


def bootstrap_indices(data, n_samples: int = 10000, seed: SeedType = None,
                      n_arrays: int = 1) -> Iterator[np.ndarray]:
    """Generator of bootstrap index sets for data whose axis 0 delineates points
    (bootstrap.py:667-680).  Yields int64[N] arrays: exactly the indices the fused
    kernel draws for resample 0, 1, ... under the same seed, materialised on the
    device.  Each array is a multiset of N iid uniform row draws listed cell by
    cell; `n_arrays` selects the shared-memory tile plan of a
    paired statistic over that many arrays (default 1)."""
    key = _seed_key(seed)
    n_rows = _n_rows(data)
    for batch in _index_batches(key, n_rows, n_arrays, int(n_samples)):
        for row in batch:
            yield row


def bootstrap_indices_independent(data: Tuple, n_samples: int = 10000,
                                  seed: SeedType = None) -> Iterator[Tuple[np.ndarray, ...]]:
    """One index set per array per resample (bootstrap.py:683-691); array k draws
    from its own Philox key."""
    key = _seed_key(seed)
    lens = [_n_rows(x) for x in data]
    gens = [
        (row for batch in _index_batches((key + k * _ARRAY_KEY_STRIDE) & _MASK64, n, 1, int(n_samples))
         for row in batch)
        for k, n in enumerate(lens)
    ]
    for _ in range(int(n_samples)):
        yield tuple(next(g) for g in gens)


def jackknife_indices(data) -> Iterator[np.ndarray]:
    """Leave-one-out index sets, ascending in the deleted row (bootstrap.py:694-705).
    Public API only: the device jackknife is closed-form and never builds these."""
    n = len(data)
    base = np.arange(0, n)
    return (np.concatenate((base[:i], base[i + 1:])) for i in range(n))


def subsample_indices(data, n_samples: int = 1000, size: float = 0.5, seed: SeedType = None):
    """Indices of `n_samples` subsamples without replacement, an int64 array of shape
    (n_samples, size) (bootstrap.py:717-744).  size >= 1 is an absolute size, 0 < size < 1
    a fraction of the data, -1 the data size itself (permuted samples).  Each row holds
    the first `size` entries of a uniform random permutation drawn on the device."""
    n_rows = len(data)
    if size == -1:
        size = n_rows
    elif 0 < size < 1:
        size = int(round(size * n_rows))
    elif size < 1:
        raise ValueError("size cannot be {0}".format(size))
    elif size > n_rows:
        raise ValueError(f"Size {size} is larger than data size {n_rows}.")
    size = int(size)
    key = _seed_key(seed)
    n_samples = int(n_samples)
    if size < 1 or n_samples < 1:
        return np.zeros((max(n_samples, 0), max(size, 0)), dtype=np.int64)
    n_pad = 1
    while n_pad < n_rows:
        n_pad <<= 1
    step = max(1, min(n_samples, 65535, (256 << 20) // (12 * n_pad)))
    parts = [_device.to_host(_device.fill_subsample(key, n_rows, size, lo, min(n_samples, lo + step)))
             for lo in range(0, n_samples, step)]
    return np.concatenate(parts, axis=0)


def bootstrap_indices_moving_block(data, n_samples: int = 10000, block_length: int = 3,
                                   wrap: bool = False, seed: SeedType = None) -> Iterator[np.ndarray]:
    """Generator of moving-block bootstrap index sets (bootstrap.py:747-787): blocks of
    `block_length` consecutive points starting at uniformly drawn positions, concatenated
    and cut to the data length; with wrap=True blocks may start anywhere and wrap around
    the end.  Block starts are Philox draws made on the device."""
    key = _seed_key(seed)
    n_rows = _n_rows(data)
    block_length = int(block_length)
    bound = n_rows if wrap else n_rows - block_length
    if block_length < 1 or bound < 1:
        raise ValueError("low >= high")          # numpy's error for an empty start range (:783)
    n_samples = int(n_samples)
    step = max(1, min(n_samples, (64 << 20) // (8 * n_rows)))
    for lo in range(0, n_samples, step):
        hi = min(n_samples, lo + step)
        for row in _device.to_host(_device.fill_moving_block(key, n_rows, block_length, wrap, lo, hi)):
            yield row
