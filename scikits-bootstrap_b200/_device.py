"""Device plumbing: tensors, streams and workspaces around the C ABI.

PyTorch is used for device memory, the current stream and (in _dist.py)
torch.distributed only; every computation is a libb200boot kernel."""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _abi, _dist


def _torch():
    import torch
    return torch


_cuda_seen = False


def require_cuda():
    global _cuda_seen
    torch = _torch()
    if not _cuda_seen:                      # asked once: a device that was there stays there
        if not torch.cuda.is_available():
            raise _abi.B200BootError(
                "no CUDA device: scikits-bootstrap_b200 runs its resampling path on a B200 only "
                "(there is no CPU fallback)")
        _cuda_seen = True
    return torch


# bytes moved between host and device by the package (bench.py reads these)
TRAFFIC = {"h2d": 0, "d2h": 0}


def to_host(t) -> np.ndarray:
    """Device tensor -> numpy array (synchronises), counted as D2H traffic."""
    TRAFFIC["d2h"] += t.numel() * t.element_size()
    return t.cpu().numpy()


def host_to_device(arr: np.ndarray, device):
    """Small host array -> device tensor, counted as H2D traffic."""
    torch = _torch()
    t = torch.as_tensor(np.ascontiguousarray(arr))
    TRAFFIC["h2d"] += t.numel() * t.element_size()
    return t.to(device)


_copy_streams = {}
SHARD_COPY_BYTES = 32 << 20      # sharded runs: host rows from this size on are copied 1/G per rank ...
SHARD_COPY_MIN_RANKS = 8         # ... from this many ranks on (measured: the split costs ~0.4 ms of host
                                 # time before K0 is enqueued; 80 MB whole per rank hides under K0 up to 4 ranks)


def _copy_stream(dev):
    torch = _torch()
    side = _copy_streams.get(dev.index)
    if side is None:
        side = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))      # the block may have pending users there
    return side


def shard_copy_async(x, device=None):
    """Sharded run (NCCL), replicated host rows: a contiguous 1-D float64 numpy array or CPU
    tensor of SHARD_COPY_BYTES or more is copied 1/G per rank on the copy stream and the ranks
    all-gather over NVLink.  HGX boards pair GPUs on a PCIe switch uplink: eight ranks copying
    80 MB each at once got ~25 GB/s apiece, 3.4 ms against K0's 2.1 ms at G = 8 (and a pageable
    source is staged by the host at ~10 GB/s: 1 ms for a tenth instead of 8 ms for the whole).
    Returns (tensor, event recorded after the gather), or None when the input is something else."""
    torch = require_cuda()
    rank, ranks = _dist.world()
    if ranks < max(2, SHARD_COPY_MIN_RANKS) or _dist.backend() != "nccl":
        return None
    if type(x) is np.ndarray:
        if not (x.ndim == 1 and x.dtype == np.float64 and x.flags.c_contiguous):
            return None
        n = x.shape[0]
    elif isinstance(x, torch.Tensor) and x.device.type == "cpu":
        if not (x.dim() == 1 and x.dtype == torch.float64 and x.is_contiguous()):
            return None
        n = x.numel()
    else:
        return None
    if n * 8 < SHARD_COPY_BYTES:
        return None
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    side = _copy_stream(dev)
    width = -(-n // ranks)
    lo, hi = min(rank * width, n), min((rank + 1) * width, n)
    src = x[lo:hi]
    if type(src) is np.ndarray:
        if not src.flags.writeable:      # torch.from_numpy wants a writable buffer (it is only read)
            src = src.copy()
        src = torch.from_numpy(src)
    whole = torch.empty(ranks * width, dtype=torch.float64, device=dev)
    part = torch.empty(width, dtype=torch.float64, device=dev)
    with torch.cuda.stream(side):
        if hi > lo:
            part[: hi - lo].copy_(src, non_blocking=True)
        if hi - lo < width:
            part[hi - lo:].zero_()
        _dist.all_gather_equal(whole, part)
    for t in (whole, part):
        t.record_stream(side)
    event = torch.cuda.Event()
    event.record(side)
    TRAFFIC["h2d"] += (hi - lo) * 8
    return whole[:n], event


def to_device_f64_async(x, device=None):
    """Like to_device_f64, but a pinned host tensor is copied on a side stream: returns
    (tensor, event) where `event` (or None) is recorded after the copy.  Work that does
    not read the data can be enqueued before the consumer waits for the event.  Long host
    rows of a sharded run take the split copy (shard_copy_async)."""
    torch = require_cuda()
    split = shard_copy_async(x, device)
    if split is not None:
        return split
    if not (isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.is_pinned()
            and x.dtype == torch.float64 and x.is_contiguous()):
        return to_device_f64(x, device), None
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    side = _copy_stream(dev)
    dst = torch.empty(x.shape, dtype=torch.float64, device=dev)
    with torch.cuda.stream(side):
        dst.copy_(x, non_blocking=True)
    dst.record_stream(side)
    event = torch.cuda.Event()
    event.record(side)
    TRAFFIC["h2d"] += x.numel() * 8
    return dst, event


def to_device_f64(x, device=None):
    """array-like / torch tensor -> contiguous float64 CUDA tensor.  float64 CUDA
    tensors are used in place (no copy)."""
    torch = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        arr = np.asarray(x)          # no host copy: the transfer below is the copy
        if arr.dtype == object:
            raise TypeError("data must be numeric")
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:      # torch.from_numpy wants a writable buffer (it is only read)
            arr = arr.copy()
        t = torch.from_numpy(arr)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    if t.device.type != "cuda":
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        TRAFFIC["h2d"] += t.numel() * 8
        t = t.to(dev, non_blocking=t.is_pinned())
    elif t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()


def stream_ptr() -> int:
    """Raw handle of torch's current CUDA stream on the current device."""
    torch = _torch()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # ~1 us; the public path costs ~15 us
    if raw is not None:
        return raw(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class Workspace:
    """Grow-only uint8 scratch tensor handed to the library."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes: int):
        torch = _torch()
        nbytes = max(int(nbytes), 256)
        if self._buf is None or self._buf.numel() < nbytes or self._buf.device.index != torch.cuda.current_device():
            self._buf = None
            self._buf = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        ptr = self._buf.data_ptr()
        off = (-ptr) % 256
        return ptr + off, self._buf.numel() - off


class _PerThreadWorkspace:
    """One grow-only workspace per (Python thread, device, CUDA stream): calls that may run
    concurrently — other threads, or the same thread under another `torch.cuda.stream(...)` —
    never share scratch; work on one stream is ordered, so reuse there is safe."""

    _MAX_STREAMS = 8          # per thread; beyond that the least recently used scratch is dropped

    def __init__(self):
        import threading
        self._local = threading.local()

    def get(self, nbytes: int):
        torch = _torch()
        table = getattr(self._local, "table", None)
        if table is None:
            table = self._local.table = {}
        key = (torch.cuda.current_device(), stream_ptr())
        ws = table.pop(key, None)
        if ws is None:
            ws = Workspace()
            while len(table) >= self._MAX_STREAMS:
                table.pop(next(iter(table)))
        table[key] = ws                       # most recently used last
        return ws.get(nbytes)


_ws = _PerThreadWorkspace()


class _NoSwitch:
    """The job already runs on the right device: nothing to enter or leave."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def device_context(arrays):
    """Context manager selecting the CUDA device the job runs on: the device of the first
    CUDA tensor among the inputs, else the current device."""
    torch = require_cuda()
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.device.type == "cuda":
            if a.device.index == torch.cuda.current_device():
                return _NO_SWITCH
            return torch.cuda.device(a.device)
    return _NO_SWITCH


def _ptr_array(tensors: Sequence) -> C.Array:
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


@functools.lru_cache(maxsize=256)
def _workspace_bytes(stat_id: int, n_rows: int, n_arrays: int, n_cols: int, n_local: int, n_alphas: int) -> int:
    return _abi.lib().b200boot_workspace_bytes(stat_id, n_rows, n_arrays, n_cols, n_local, n_alphas)


def workspace_for(stat_id: int, n_rows: int, n_arrays: int, n_cols: int, n_local: int, n_alphas: int):
    return _ws.get(_workspace_bytes(stat_id, n_rows, n_arrays, n_cols, n_local, n_alphas))


def _ws_budget() -> int:
    """Upper bound on the scratch a single K1 call may use (bytes); larger jobs are
    processed as consecutive resample sub-ranges, which cannot change the result
    because every resample depends only on (seed, resample id)."""
    import os
    return int(os.environ.get("B200BOOT_WS_BYTES", 12 << 30))


def copy_stream(device_index: int):
    """The side stream host-to-device copies run on (one per device)."""
    torch = _torch()
    side = _copy_streams.get(device_index)
    if side is None:
        side = _copy_streams[device_index] = torch.cuda.Stream(device=torch.device("cuda", device_index))
    return side


def resample_stat(arrays: Sequence, stat_id: int, seed: int, lo: int, hi: int, out=None, ready=None,
                  host_rows=None):
    """K1 on contiguous float64[N] CUDA tensors -> float64[hi-lo] CUDA tensor
    (float64[n_out][hi-lo] for a statistic with several outputs).  host_rows: the numpy arrays
    whose rows `arrays` are still to receive (copied under the occupancy chains of the first
    sub-range, b200boot_resample_stat_host)."""
    torch = _torch()
    L = _abi.lib()
    n_rows = arrays[0].numel()
    n_local = hi - lo
    n_out = _abi.STAT_OUTPUTS.get(stat_id, 1)
    if out is None:
        shape = n_local if n_out == 1 else (n_out, n_local)
        out = torch.empty(shape, dtype=torch.float64, device=arrays[0].device)
    step = max(n_local, 1)
    budget = _ws_budget()
    while step > 32 and L.b200boot_workspace_bytes(stat_id, n_rows, len(arrays), 1, step, 1) > budget:
        step = (step + 1) // 2
    ptrs = _ptr_array(arrays)
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        ws, wsz = workspace_for(stat_id, n_rows, len(arrays), 1, b - a, 1)
        whole = n_out == 1 or (a == lo and b == hi)
        dst = out[a - lo:] if n_out == 1 else (out if whole else out.new_empty((n_out, b - a)))
        if host_rows is not None and a == lo:
            side = copy_stream(arrays[0].device.index)
            hp = (C.c_void_p * len(host_rows))(*[h.ctypes.data for h in host_rows])
            for t in arrays:
                t.record_stream(side)
            TRAFFIC["h2d"] += 8 * n_rows * len(host_rows)
            _abi.check(L.b200boot_resample_stat_host(hp, ptrs, len(arrays), n_rows, stat_id, seed, a, b,
                                                     dst.data_ptr(), ws, wsz, stream_ptr(), side.cuda_stream))
        else:
            # `ready`: event recorded after the arrays' host-to-device copy on another stream; the
            # first sub-range waits for it (after its occupancy chains), later ones follow in order
            ev = ready.cuda_event if (ready is not None and a == lo) else None
            _abi.check(L.b200boot_resample_stat_after(ptrs, len(arrays), n_rows, stat_id, seed, a, b,
                                                      dst.data_ptr(), ws, wsz, stream_ptr(), ev))
        if not whole:
            out[:, a - lo:b - lo] = dst
    return out


def resample_stat_moving_block(arrays: Sequence, stat_id: int, seed: int, block_length: int, wrap: bool,
                               lo: int, hi: int):
    """K1b: moment statistic of moving-block resamples -> float64[hi-lo] ([n_out][hi-lo])."""
    torch = _torch()
    n_rows = arrays[0].numel()
    n_out = _abi.STAT_OUTPUTS.get(stat_id, 1)
    out = torch.empty(hi - lo if n_out == 1 else (n_out, hi - lo), dtype=torch.float64, device=arrays[0].device)
    ws, wsz = workspace_for(stat_id, n_rows, len(arrays), 1, 0, 1)
    _abi.check(_abi.lib().b200boot_resample_stat_moving_block(_ptr_array(arrays), len(arrays), n_rows, stat_id, seed,
                                                              int(block_length), int(bool(wrap)), lo, hi,
                                                              out.data_ptr(), ws, wsz, stream_ptr()))
    return out


class _SmallCallBuffers:
    """Per (thread, device, stream): the statistic vector, the device result block (whose ticket
    counter behind the results starts at zero), its pinned host mirror, and — for host inputs —
    a pinned staging block and the device rows of b200boot_ci_small_host."""

    def __init__(self):
        import threading
        self._local = threading.local()

    def get(self, n_samples: int):
        torch = _torch()
        table = getattr(self._local, "table", None)
        if table is None:
            table = self._local.table = {}
        key = (torch.cuda.current_device(), stream_ptr())
        ent = table.get(key)
        if ent is None or ent.stat.numel() < n_samples:
            if len(table) >= 8:
                table.clear()
            ent = table[key] = _SmallEntry(torch, max(n_samples, 4096))
        return ent


class _SmallEntry:
    __slots__ = ("stat", "res_dev", "res_host", "res_np", "staging", "rows", "stat_ptr", "res_dev_ptr",
                 "res_host_ptr", "staging_ptrs", "rows_ptrs", "alphas", "alphas_ptr", "host_ptrs")

    def __init__(self, torch, n_stat: int):
        self.stat = torch.empty(n_stat, dtype=torch.float64, device="cuda")
        self.res_dev = torch.zeros(48, dtype=torch.float64, device="cuda")
        self.res_host = torch.zeros(32, dtype=torch.float64).pin_memory()
        self.res_np = self.res_host.numpy()
        self.staging = torch.empty((2, SMALL_MAX_CELL), dtype=torch.float64).pin_memory()
        self.rows = torch.empty((2, SMALL_MAX_CELL), dtype=torch.float64, device="cuda")
        self.stat_ptr = self.stat.data_ptr()
        self.res_dev_ptr = self.res_dev.data_ptr()
        self.res_host_ptr = self.res_host.data_ptr()
        self.staging_ptrs = (C.c_void_p * 2)(self.staging[0].data_ptr(), self.staging[1].data_ptr())
        self.rows_ptrs = (C.c_void_p * 2)(self.rows[0].data_ptr(), self.rows[1].data_ptr())
        self.host_ptrs = (C.c_void_p * 2)()
        self.alphas = (C.c_double * 8)()
        self.alphas_ptr = C.addressof(self.alphas)


SMALL_MAX_CELL = 28672          # rows x arrays of the fused small call (kFusedMaxCellDoubles)
_small = _SmallCallBuffers()


@functools.lru_cache(maxsize=256)
def _ci_small_supported(stat_id: int, n_arrays: int, n_rows: int, n_samples: int, n_alphas: int) -> bool:
    return bool(_abi.lib().b200boot_ci_small_supported(stat_id, n_arrays, n_rows, n_samples, n_alphas))


def ci_small_supported(stat_id: int, n_arrays: int, n_rows: int, n_samples: int, n_alphas: int) -> bool:
    return _ci_small_supported(stat_id, n_arrays, n_rows, n_samples, n_alphas)


def ci_small(arrays: Sequence, stat_id: int, seed: int, n_samples: int, alphas: np.ndarray, bca: bool,
             need_jack: bool):
    """One library call (one kernel) for a whole short ci(): returns (stat float64[n_samples] CUDA
    view, endpoints[A], ranks[A] int64, less, jack[8]) — ranks are the device's speculative ones.
    `arrays`: float64 CUDA tensors, or contiguous float64 numpy arrays (staged through pinned
    memory inside the call)."""
    L = _abi.lib()
    n_arrays = len(arrays)
    a = np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1)
    na = len(a)
    ent = _small.get(n_samples)
    for i in range(na):
        ent.alphas[i] = a[i]
    s = stream_ptr()
    if isinstance(arrays[0], np.ndarray):
        n_rows = arrays[0].shape[0]
        ws, wsz = _ws.get(4096)
        for i, x in enumerate(arrays):
            ent.host_ptrs[i] = x.ctypes.data
        TRAFFIC["h2d"] += 8 * n_rows * n_arrays
        rc = L.b200boot_ci_small_host(ent.host_ptrs, ent.staging_ptrs, ent.rows_ptrs, n_arrays, n_rows, stat_id, seed,
                                      n_samples, ent.alphas_ptr, na, int(bca), int(need_jack), ent.stat_ptr,
                                      ent.res_dev_ptr, ent.res_host_ptr, ws, wsz, s)
        if rc:
            _abi.check(rc)
    else:
        n_rows = arrays[0].numel()
        ws, wsz = _ws.get(4096)
        rc = L.b200boot_ci_small(_ptr_array(arrays), n_arrays, n_rows, stat_id, seed, n_samples,
                                 ent.alphas_ptr, na, int(bca), int(need_jack), ent.stat_ptr,
                                 ent.res_dev_ptr, ent.res_host_ptr, ws, wsz, s)
        if rc:
            _abi.check(rc)
        _sync_stream(s)
    TRAFFIC["d2h"] += 32 * 8
    r = ent.res_np
    return (ent.stat[:n_samples], r[:na].copy(), r[8:8 + na].astype(np.int64), int(r[16]), r[17:25].copy())


def pval_small(arrays: Sequence, stat_id: int, seed: int, n_samples: int, op: int, t0: float, t1: float) -> int:
    """One kernel for a short pval(): #{statistic op t0 [, t1]} over n_samples resamples.  `arrays`:
    float64 CUDA tensors or contiguous float64 numpy arrays."""
    L = _abi.lib()
    n_arrays = len(arrays)
    ent = _small.get(n_samples)
    ws, wsz = _ws.get(4096)
    if isinstance(arrays[0], np.ndarray):
        n_rows = arrays[0].shape[0]
        for i, x in enumerate(arrays):
            ent.host_ptrs[i] = x.ctypes.data
        TRAFFIC["h2d"] += 8 * n_rows * n_arrays
        rc = L.b200boot_pval_small(ent.rows_ptrs, ent.host_ptrs, ent.staging_ptrs, n_arrays, n_rows, stat_id, seed,
                                   n_samples, op, t0, t1, ent.stat_ptr, ent.res_dev_ptr, ent.res_host_ptr, ws, wsz,
                                   stream_ptr())
    else:
        n_rows = arrays[0].numel()
        rc = L.b200boot_pval_small(_ptr_array(arrays), None, None, n_arrays, n_rows, stat_id, seed, n_samples, op,
                                   t0, t1, ent.stat_ptr, ent.res_dev_ptr, ent.res_host_ptr, ws, wsz, stream_ptr())
    if rc:
        _abi.check(rc)
    TRAFFIC["d2h"] += 32 * 8
    return int(ent.res_np[16])


def ci_median_small_supported(n_rows: int, n_samples: int, n_alphas: int) -> bool:
    return 1 <= n_rows <= MEDIAN_RESIDENT_ROWS and ci_tail_supported(n_samples, n_alphas)


def ci_median_small(x, seed: int, n_samples: int, alphas: np.ndarray, bca: bool, need_jack: bool, rule=None):
    """One library call for ci() of np.median / a quantile of one resident column.  x: a contiguous
    float64 numpy array or CUDA tensor.  Returns (stat CUDA view, endpoints[A], ranks[A] int64, less,
    jack[8], has_nan)."""
    torch = _torch()
    L = _abi.lib()
    a = np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1)
    na = len(a)
    ent = _small.get(n_samples)
    for i in range(na):
        ent.alphas[i] = a[i]
    if isinstance(x, np.ndarray):
        n_rows = x.shape[0]
        TRAFFIC["h2d"] += 8 * n_rows
        data_ptr, host_ptr, staging_ptr = ent.rows_ptrs[0], x.ctypes.data, ent.staging_ptrs[0]
    else:
        n_rows = x.numel()
        data_ptr, host_ptr, staging_ptr = x.data_ptr(), None, None
    ws, wsz = workspace_for(_abi.STAT_MEDIAN, n_rows, 1, 1, n_samples, 1)
    rc = L.b200boot_ci_median_small(data_ptr, host_ptr, staging_ptr, n_rows, seed, n_samples, ent.alphas_ptr, na,
                                    int(bca), int(need_jack), _rule_ptr(rule), ent.stat_ptr, ent.res_dev_ptr,
                                    ent.res_host_ptr, ws, wsz, stream_ptr())
    if rc:
        _abi.check(rc)
    TRAFFIC["d2h"] += 32 * 8
    r = ent.res_np
    return (ent.stat[:n_samples], r[:na].copy(), r[8:8 + na].astype(np.int64), int(r[16]), r[17:25].copy(),
            bool(r[25] != 0.0))


def ci_tail_supported(n_samples: int, n_alphas: int) -> bool:
    return 1 <= n_samples <= (1 << 21) and 1 <= n_alphas <= 8


def ci_tail(stat, jack_dev, alphas: np.ndarray, bca: bool):
    """The single-CTA tail over a statistic vector already on the device (float64[n] CUDA,
    contiguous): returns (endpoints[A], ranks[A] int64, less, jack[8]) after one host wait;
    ranks are the device's speculative ones.  jack_dev: float64[8] CUDA (K3 output) or None."""
    L = _abi.lib()
    a = np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1)
    na = len(a)
    n = stat.numel()
    ent = _small.get(1)
    for i in range(na):
        ent.alphas[i] = a[i]
    s = stream_ptr()
    rc = L.b200boot_ci_tail(stat.data_ptr(), n, jack_dev.data_ptr() if jack_dev is not None else None,
                            ent.alphas_ptr, na, int(bca), ent.res_dev_ptr, ent.res_host_ptr, s)
    if rc:
        _abi.check(rc)
    _sync_stream(s)
    TRAFFIC["d2h"] += 32 * 8
    r = ent.res_np
    return r[:na].copy(), r[8:8 + na].astype(np.int64), int(r[16]), r[17:25].copy()


def _sync_stream(stream: int) -> None:
    """The one host wait of a small call: cudaStreamSynchronize through the library (a few
    microseconds cheaper than building torch's stream object)."""
    _abi.check(_abi.lib().b200boot_stream_synchronize(stream))


def resample_stat_indexed(arrays: Sequence, stat_id: int, indices):
    """K1i: indices int64[n_sets][N] CUDA tensor -> float64[n_sets] ([n_out][n_sets])."""
    torch = _torch()
    n_rows = arrays[0].numel()
    n_sets = indices.shape[0]
    n_out = _abi.STAT_OUTPUTS.get(stat_id, 1)
    out = torch.empty(n_sets if n_out == 1 else (n_out, n_sets), dtype=torch.float64, device=arrays[0].device)
    ws, wsz = workspace_for(stat_id, n_rows, len(arrays), 1, n_sets, 1)
    _abi.check(_abi.lib().b200boot_resample_stat_indexed(_ptr_array(arrays), len(arrays), n_rows, stat_id,
                                                         indices.data_ptr(), n_sets, out.data_ptr(), ws, wsz,
                                                         stream_ptr()))
    return out


def columns_block_width(stat_id: int, n_rows: int) -> int:
    return _abi.lib().b200boot_columns_block_width(stat_id, n_rows)


def resample_stat_columns(data2d, stat_id: int, seed: int, lo: int, hi: int):
    """K1c: data float64[N][D] -> float64[D][hi-lo] (all columns share the index set)."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    out = torch.empty((n_cols, hi - lo), dtype=torch.float64, device=data2d.device)
    ws, wsz = workspace_for(stat_id, n_rows, 1, max(n_cols, 2), hi - lo, 1)
    _abi.check(_abi.lib().b200boot_resample_stat_columns(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0),
                                                         stat_id, seed, lo, hi, out.data_ptr(), ws, wsz,
                                                         stream_ptr()))
    return out


GEMM_MIN_COLS = 64           # long data up to 2^18 rows: below this one K1 pass per column wins
GEMM_MIN_COLS_HUGE = 1024    # longer data: the byte count matrix no longer fits L2


def gemm_pays(n_rows: int, n_cols: int) -> bool:
    """Measured crossover of the GEMM form against the gather kernels (mean; B200).
    One-cell data (K6i builds the count matrix itself; K1c is the alternative): from 8 columns on
    once N * D >= 20 000 (N = 7000: D = 8 -> 0.73 vs 1.19 ms; N = 1000: D = 24 -> 0.20 vs
    0.26 ms, D = 16 -> 0.19 vs 0.17 ms).  Longer data, where the alternative is one K1 pass per
    column and the count matrix is built from the materialised indices of the tiled plan (an
    atomic per draw on an n x N byte matrix, plus the index arrays through HBM):
    tools/long_cols_crossover.py, n = 2000: N = 1e5: 0.2 ms per column against 10.1 ms + 0.02 ms
    per column (crossover ~50 columns); N = 1e6: 0.83 ms per column against 640 ms (~770 columns:
    the 2 GB count matrix is out of L2 and every draw is a DRAM read-modify-write)."""
    if n_rows <= MEDIAN_RESIDENT_ROWS:
        return n_cols >= 8 and n_rows * n_cols >= 20000
    return n_cols >= (GEMM_MIN_COLS if n_rows <= (1 << 18) else GEMM_MIN_COLS_HUGE)


_GEMM_COUNT_BYTES = 1 << 30  # count-matrix chunk


def gemm_supported(stat_id: int, n_rows: int) -> bool:
    """K6i covers mean / sum / var / std for N < 2^25 (int32 digit sums stay exact): one-cell
    data build the count matrix directly, longer data through the materialised indices of the
    tiled plan (K2)."""
    return stat_id in (_abi.STAT_MEAN, _abi.STAT_SUM, _abi.STAT_VAR, _abi.STAT_STD) and 1 <= n_rows < 2**25


def resample_stat_gemm(data2d, stat_id: int, seed: int, lo: int, hi: int):
    """K6i: data float64[N][D] -> float64[D][hi-lo].  Column moment sums of every resample as the
    resamples' count matrix times the data on the tensor cores, in exact integer arithmetic
    (tcgen05 kind::i8, csrc/k6i_mma.cuh); var / std use the column-centred data and its squares, as
    K1c does.  Data beyond one cell take their counts from the materialised indices of the tiled
    plan (K2), in chunks of resamples."""
    torch = _torch()
    L = _abi.lib()
    n_rows, n_cols = data2d.shape
    n_local = hi - lo
    out = torch.empty((n_cols, n_local), dtype=torch.float64, device=data2d.device)
    two = stat_id in (_abi.STAT_VAR, _abi.STAT_STD)
    one_cell = bool(L.b200boot_gemm_supported(stat_id, n_rows))
    import os
    budget = int(os.environ.get("B200BOOT_GEMM_COUNT_BYTES", _GEMM_COUNT_BYTES))
    per_resample = n_rows + (0 if one_cell else 8 * n_rows) + (8 * n_cols if two else 0)
    step = max(1, min(n_local, budget // per_resample))
    for a in range(lo, hi, step):
        b = min(hi, a + step)
        whole = a == lo and b == hi
        dst = out if whole else torch.empty((n_cols, b - a), dtype=torch.float64, device=data2d.device)
        second = torch.empty_like(dst) if two else None
        idx = None if one_cell else fill_indices(seed, n_rows, 1, a, b)
        need = L.b200boot_count_gemm_ws_bytes(n_rows, n_cols, b - a)
        ws, wsz = _ws.get(need)
        _abi.check(L.b200boot_resample_stat_columns_gemm(
            data2d.data_ptr(), n_rows, n_cols, data2d.stride(0), stat_id, seed, a, b,
            idx.data_ptr() if idx is not None else None, dst.data_ptr(),
            second.data_ptr() if two else None, ws, wsz, stream_ptr()))
        if not whole:
            out[:, a - lo:b - lo] = dst
        del idx, second
    return out


def jackknife_columns(data2d, stat_id: int):
    """K3c: float64[D][8] (ostat, jmean, d2, d3, accel, centre, ...) per column."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    out = torch.empty((n_cols, 8), dtype=torch.float64, device=data2d.device)
    _abi.check(_abi.lib().b200boot_jackknife_columns(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0),
                                                     stat_id, out.data_ptr(), stream_ptr()))
    return out


def rank_rule(n_rows: int, q):
    """The order statistics behind np.quantile(x, q) (method 'linear') for samples of n_rows
    and n_rows - 1 rows, with numpy's own rounding of the virtual index (n - 1) * q
    (numpy/lib/_function_base_impl.py: _QuantileMethods['linear'], _get_indexes, _get_gamma).
    None (np.median) -> NULL."""
    if q is None:
        return None
    import ctypes as C
    import numpy as np
    r = _abi.RankRule()
    r.mode = 1
    for n, pre in ((n_rows, ""), (max(n_rows - 1, 1), "loo_")):
        vi = (n - 1) * np.float64(q)
        lo = int(np.floor(vi))
        if vi >= n - 1:                      # at or above the last rank: the maximum itself
            k_lo = k_hi = n - 1
            gamma = 0.0
        else:
            k_lo, k_hi, gamma = lo, lo + 1, float(vi - lo)
        setattr(r, pre + "k_lo", k_lo)
        setattr(r, pre + "k_hi", k_hi)
        setattr(r, pre + "gamma", gamma)
    return r


def _rule_ptr(rule):
    import ctypes as C
    return None if rule is None else C.cast(C.pointer(rule), C.c_void_p)


def resample_median_columns(data2d, seed: int, lo: int, hi: int, rule=None, with_jackknife: bool = False):
    """K1m: data float64[N][D] -> float64[D][hi-lo]; with_jackknife: also the K3m summary
    float64[D][8] of the same (already sorted) columns, returned as a second tensor."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    out = torch.empty((n_cols, hi - lo), dtype=torch.float64, device=data2d.device)
    jack = torch.empty((n_cols, 8), dtype=torch.float64, device=data2d.device) if with_jackknife else None
    ws, wsz = workspace_for(_abi.STAT_MEDIAN, n_rows, 1, n_cols, hi - lo, 1)
    _abi.check(_abi.lib().b200boot_resample_median_columns(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0),
                                                           seed, lo, hi, out.data_ptr(), ws, wsz, stream_ptr(),
                                                           _rule_ptr(rule),
                                                           jack.data_ptr() if with_jackknife else None))
    return (out, jack) if with_jackknife else out


def resample_median_columns_indexed(data2d, indices, rule=None):
    """K1m parity mode: indices int64[n_sets][N] -> float64[D][n_sets]."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    n_sets = indices.shape[0]
    out = torch.empty((n_cols, n_sets), dtype=torch.float64, device=data2d.device)
    ws, wsz = workspace_for(_abi.STAT_MEDIAN, n_rows, 1, n_cols, n_sets, 1)
    _abi.check(_abi.lib().b200boot_resample_median_columns_indexed(
        data2d.data_ptr(), n_rows, n_cols, data2d.stride(0), indices.data_ptr(), n_sets, out.data_ptr(),
        ws, wsz, stream_ptr(), _rule_ptr(rule)))
    return out


MEDIAN_RESIDENT_ROWS = 28672     # csrc/k1m_median.cuh: kMedMaxRows


def resample_median_sorted(sorted_vals, seed: int, lo: int, hi: int, out=None, rule=None):
    """K1mt: median of sorted_vals[bootstrap indices] for resamples [lo, hi) -> float64[1][hi-lo]."""
    torch = _torch()
    n_rows = sorted_vals.numel()
    if out is None:
        out = torch.empty((1, hi - lo), dtype=torch.float64, device=sorted_vals.device)
    ws, wsz = workspace_for(_abi.STAT_MEDIAN, n_rows, 1, 1, hi - lo, 1)
    _abi.check(_abi.lib().b200boot_resample_median_sorted(sorted_vals.data_ptr(), n_rows, seed, lo, hi,
                                                          out.data_ptr(), ws, wsz, stream_ptr(),
                                                          _rule_ptr(rule)))
    return out


def jackknife_median_sorted(sorted_vals, rule=None):
    torch = _torch()
    out = torch.empty((1, 8), dtype=torch.float64, device=sorted_vals.device)
    _abi.check(_abi.lib().b200boot_jackknife_median_sorted(sorted_vals.data_ptr(), sorted_vals.numel(),
                                                           out.data_ptr(), stream_ptr(), _rule_ptr(rule)))
    return out


def nan_columns(data2d):
    """(flags int32[D + 1] CUDA tensor, host bool[D]): which columns of an (N, D) tensor hold NaNs."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    flags = torch.empty(n_cols + 1, dtype=torch.int32, device=data2d.device)
    _abi.check(_abi.lib().b200boot_nan_columns(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0), flags.data_ptr(),
                                               stream_ptr()))
    return flags, to_host(flags)[:n_cols].astype(bool)


def nan_fixup(src, stride: int, n_rows: int, indices, out_row):
    """out_row[set] = NaN wherever indices[set] draws a NaN of src[:: stride]."""
    _abi.check(_abi.lib().b200boot_nan_fixup(src.data_ptr(), stride, n_rows, indices.data_ptr(), indices.shape[0],
                                             out_row.data_ptr(), stream_ptr()))


def nan_poison_rows(table, flags):
    n_cols, width = table.shape
    _abi.check(_abi.lib().b200boot_nan_poison_rows(table.data_ptr(), n_cols, width, flags.data_ptr(), stream_ptr()))


MEDIAN_ROWS_MAX = 1 << 24        # csrc/k1mr_rows.cuh: kRowsMaxRows


def rank_columns(data2d):
    """K1mr preparation: (sorted float64[D][N], rank int32[N][D4]) of an (N, D) CUDA tensor; the
    rank rows are padded to a multiple of four columns (D4) so that the select reads a row's group
    of four ranks with one 128-bit load (the pad entries are never interpreted)."""
    torch = _torch()
    n_rows, n_cols = data2d.shape
    sorted_vals = torch.empty((n_cols, n_rows), dtype=torch.float64, device=data2d.device)
    rank = torch.empty((n_rows, (n_cols + 3) // 4 * 4), dtype=torch.int32, device=data2d.device)
    L = _abi.lib()
    need = L.b200boot_rank_columns_ws_bytes(n_rows, n_cols)
    ws, wsz = _ws.get(need)
    _abi.check(L.b200boot_rank_columns_ld(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0), sorted_vals.data_ptr(),
                                          rank.data_ptr(), rank.stride(0), ws, wsz, stream_ptr()))
    return sorted_vals, rank


def median_rows_indexed(sorted_vals, rank, indices, out, col_offset: int, rule=None):
    """K1mr: order statistics of every column for index sets int64[n_sets][N] ->
    out[:, col_offset : col_offset + n_sets] (out float64[D][n_total])."""
    n_cols, n_rows = sorted_vals.shape
    n_sets = indices.shape[0]
    _abi.check(_abi.lib().b200boot_median_rows_indexed_ld(sorted_vals.data_ptr(), rank.data_ptr(), rank.stride(0),
                                                          n_rows, n_cols, indices.data_ptr(), n_sets,
                                                          out.data_ptr() + 8 * col_offset, out.stride(0),
                                                          stream_ptr(), _rule_ptr(rule)))
    return out


def fill_indices(seed: int, n_rows: int, n_arrays: int, lo: int, hi: int):
    """K2 -> int64[hi-lo][N] CUDA tensor."""
    torch = require_cuda()
    out = torch.empty((hi - lo, n_rows), dtype=torch.int64, device="cuda")
    plan = _abi.make_plan(n_rows, n_arrays)
    ws, wsz = _ws.get(plan.n_tiles * (hi - lo) * 4 * 17 + 4096)
    _abi.check(_abi.lib().b200boot_fill_indices(seed, n_rows, n_arrays, lo, hi, out.data_ptr(), ws, wsz,
                                                stream_ptr()))
    return out


def fill_moving_block(seed: int, n_rows: int, block_length: int, wrap: bool, lo: int, hi: int):
    """K2b -> int64[hi-lo][N] CUDA tensor."""
    torch = require_cuda()
    out = torch.empty((hi - lo, n_rows), dtype=torch.int64, device="cuda")
    _abi.check(_abi.lib().b200boot_fill_moving_block(seed, n_rows, block_length, int(bool(wrap)), lo, hi,
                                                     out.data_ptr(), stream_ptr()))
    return out


def fill_subsample(seed: int, n_rows: int, size: int, lo: int, hi: int):
    """K2c -> int64[hi-lo][size] CUDA tensor."""
    torch = require_cuda()
    out = torch.empty((hi - lo, size), dtype=torch.int64, device="cuda")
    n_pad = 1
    while n_pad < n_rows:
        n_pad <<= 1
    ws, wsz = _ws.get(n_pad * (hi - lo) * 12 + 4096)
    _abi.check(_abi.lib().b200boot_fill_subsample(seed, n_rows, size, lo, hi, out.data_ptr(), ws, wsz,
                                                  stream_ptr()))
    return out


def tile_counts(seed: int, n_rows: int, n_arrays: int, lo: int, hi: int, with_classes: bool = False):
    """K0 occupancies: int32[n_tiles][hi-lo] (and int32[n_full][hi-lo][16] class counts)."""
    torch = require_cuda()
    plan = _abi.make_plan(n_rows, n_arrays)
    out = torch.empty((plan.n_tiles, hi - lo), dtype=torch.int32, device="cuda")
    cls = torch.empty((plan.n_full, hi - lo, 16), dtype=torch.int32, device="cuda") if with_classes else None
    _abi.check(_abi.lib().b200boot_tile_counts(seed, n_rows, n_arrays, lo, hi, out.data_ptr(),
                                               cls.data_ptr() if cls is not None else None, stream_ptr()))
    return (out, cls) if with_classes else out


def tile_counts_packed(seed: int, n_rows: int, n_arrays: int, lo: int, hi: int):
    """K0 as the resample path runs it: int32[n_tiles][hi-lo], uint16[n_full][hi-lo][16]."""
    torch = require_cuda()
    plan = _abi.make_plan(n_rows, n_arrays)
    L = _abi.lib()
    out = torch.empty((plan.n_tiles, hi - lo), dtype=torch.int32, device="cuda")
    cls = torch.empty((plan.n_full, hi - lo, 16), dtype=torch.uint16, device="cuda")
    ws = torch.empty(L.b200boot_tile_counts_packed_ws_bytes(n_rows, n_arrays, hi - lo), dtype=torch.uint8, device="cuda")
    _abi.check(L.b200boot_tile_counts_packed(seed, n_rows, n_arrays, lo, hi, out.data_ptr(), cls.data_ptr(),
                                             ws.data_ptr(), ws.numel(), stream_ptr()))
    return out, cls


def gather(data, idx):
    torch = _torch()
    out = torch.empty(idx.numel(), dtype=torch.float64, device=data.device)
    _abi.check(_abi.lib().b200boot_gather(data.data_ptr(), data.numel(), idx.data_ptr(), idx.numel(),
                                          out.data_ptr(), stream_ptr()))
    return out


def jackknife(arrays: Sequence, stat_id: int):
    """K3 -> float64[8] CUDA tensor (ostat, jmean, d2, d3, accel, ...); [n_out][8] for a
    statistic with several outputs."""
    torch = _torch()
    n_rows = arrays[0].numel()
    n_out = _abi.STAT_OUTPUTS.get(stat_id, 1)
    out = torch.empty(8 if n_out == 1 else (n_out, 8), dtype=torch.float64, device=arrays[0].device)
    ws, wsz = workspace_for(stat_id, n_rows, len(arrays), 1, 0, 1)
    _abi.check(_abi.lib().b200boot_jackknife(_ptr_array(arrays), len(arrays), n_rows, stat_id, out.data_ptr(),
                                             ws, wsz, stream_ptr()))
    return out


def jackknife_median_columns(data2d, rule=None):
    torch = _torch()
    n_rows, n_cols = data2d.shape
    out = torch.empty((n_cols, 8), dtype=torch.float64, device=data2d.device)
    ws, wsz = workspace_for(_abi.STAT_MEDIAN, n_rows, 1, n_cols, 0, 1)
    _abi.check(_abi.lib().b200boot_jackknife_median_columns(data2d.data_ptr(), n_rows, n_cols, data2d.stride(0),
                                                            out.data_ptr(), ws, wsz, stream_ptr(),
                                                            _rule_ptr(rule)))
    return out


def abc_identity_averages(x, ep: float):
    """method='abc': the 2N perturbed weighted means (numpy arithmetic) -> (tp, tm) numpy arrays."""
    torch = _torch()
    n = x.numel()
    out = torch.empty((2, n), dtype=torch.float64, device=x.device)
    _abi.check(_abi.lib().b200boot_abc_identity_averages(x.data_ptr(), n, float(ep), out[0].data_ptr(),
                                                         out[1].data_ptr(), stream_ptr()))
    h = to_host(out)
    return h[0], h[1]


def weighted_averages(x, weights) -> np.ndarray:
    """np.average(x, weights=w) for every row w of `weights` (numpy arithmetic, on the device);
    weights=None: np.average(x) itself, as a one-element array."""
    torch = _torch()
    w = None if weights is None else host_to_device(np.ascontiguousarray(weights, dtype=np.float64), x.device)
    n_vec = 1 if w is None else w.shape[0]
    out = torch.empty(n_vec, dtype=torch.float64, device=x.device)
    _abi.check(_abi.lib().b200boot_weighted_averages(x.data_ptr(), x.numel(), None if w is None else w.data_ptr(),
                                                     n_vec, out.data_ptr(), stream_ptr()))
    return to_host(out)


def combine(op: int, a, b):
    """f(a[i], b[i]) of two statistic vectors (multi='independent')."""
    torch = _torch()
    out = torch.empty_like(a)
    _abi.check(_abi.lib().b200boot_combine(op, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), stream_ptr()))
    return out


def jackknife_values(arrays: Sequence, stat_id: int):
    """Leave-one-out statistic of every row (moment statistics): (values float64[N], summary float64[8])."""
    torch = _torch()
    n_rows = arrays[0].numel()
    values = torch.empty(n_rows, dtype=torch.float64, device=arrays[0].device)
    summary = torch.empty(8, dtype=torch.float64, device=arrays[0].device)
    ws, wsz = workspace_for(stat_id, n_rows, len(arrays), 1, 0, 1)
    _abi.check(_abi.lib().b200boot_jackknife_values(_ptr_array(arrays), len(arrays), n_rows, stat_id,
                                                    values.data_ptr(), summary.data_ptr(), ws, wsz, stream_ptr()))
    return values, summary


def jackknife_values_sorted(sorted_vals, rule=None):
    """The same for a rank statistic of an ascending array (values listed by sorted position)."""
    torch = _torch()
    n_rows = sorted_vals.numel()
    values = torch.empty(n_rows, dtype=torch.float64, device=sorted_vals.device)
    summary = torch.empty(8, dtype=torch.float64, device=sorted_vals.device)
    _abi.check(_abi.lib().b200boot_jackknife_values_sorted(sorted_vals.data_ptr(), n_rows, values.data_ptr(),
                                                           summary.data_ptr(), stream_ptr(), _rule_ptr(rule)))
    return values, summary


def jackknife_pooled(op: int, loo_a, s_a, loo_b, s_b):
    """Pooled jackknife summary float64[1][8] of f(s_a, s_b) from the two leave-one-out vectors."""
    torch = _torch()
    out = torch.empty((1, 8), dtype=torch.float64, device=loo_a.device)
    ws, wsz = _ws.get(64 << 10)
    _abi.check(_abi.lib().b200boot_jackknife_pooled(op, loo_a.data_ptr(), loo_a.numel(), s_a.data_ptr(),
                                                    loo_b.data_ptr(), loo_b.numel(), s_b.data_ptr(),
                                                    out.data_ptr(), ws, wsz, stream_ptr()))
    return out


def count(stat, op: int, t0, t1=None):
    """K4 on stat float64[D][n] -> int64[D] CUDA tensor."""
    torch = _torch()
    n_cols, n_local = stat.shape
    out = torch.empty(n_cols, dtype=torch.int64, device=stat.device)
    _abi.check(_abi.lib().b200boot_count(stat.data_ptr(), n_local, n_cols, op, t0.data_ptr(),
                                         t1.data_ptr() if t1 is not None else None, out.data_ptr(),
                                         stream_ptr()))
    return out


class Select:
    """K5 radix-descent select with an optional hook between the histogram and
    pick steps (the all-reduce of the sharded path)."""

    def __init__(self, stat, ranks):
        torch = _torch()
        self.stat = stat
        self.n_cols, self.n_local = stat.shape
        self.ranks = ranks.contiguous()              # int64[A][D] on device
        self.n_alphas = ranks.shape[0]
        need = _abi.lib().b200boot_workspace_bytes(_abi.STAT_MEAN, 1, 1, self.n_cols, 1, self.n_alphas)
        self._buf = torch.empty(need + 256, dtype=torch.uint8, device=stat.device)
        ptr = self._buf.data_ptr()
        off = (-ptr) % 256
        self.ws, self.wsz = ptr + off, self._buf.numel() - off

    def hist_tensor(self):
        torch = _torch()
        p = C.c_void_p()
        n = C.c_int64()
        _abi.check(_abi.lib().b200boot_select_hist_buffer(self.n_alphas, self.n_cols, self.ws, self.wsz,
                                                          C.byref(p), C.byref(n)))
        off = p.value - self._buf.data_ptr()
        return self._buf[off:off + n.value * 8].view(torch.int64)

    def run(self, sharded: bool = False):
        """sharded=False: one library call.  sharded=True: the same passes with the
        histogram all-reduced across ranks between hist and pick (_dist)."""
        from . import _dist
        torch = _torch()
        L, s = _abi.lib(), stream_ptr()
        out = torch.empty((self.n_alphas, self.n_cols), dtype=torch.float64, device=self.stat.device)
        if not sharded:
            _abi.check(L.b200boot_select(self.stat.data_ptr(), self.n_local, self.n_cols, self.ranks.data_ptr(),
                                         self.n_alphas, out.data_ptr(), self.ws, self.wsz, s))
            return out
        a, d, ws, wsz = self.n_alphas, self.n_cols, self.ws, self.wsz

        def end():
            _abi.check(L.b200boot_select_end(a, d, out.data_ptr(), ws, wsz, s))
            return out

        return _dist.radix_select_passes(
            begin=lambda: _abi.check(L.b200boot_select_begin(self.ranks.data_ptr(), a, d, ws, wsz, s)),
            hist=lambda p: _abi.check(L.b200boot_select_hist(self.stat.data_ptr(), self.n_local, d, a, p, ws, wsz, s)),
            pick=lambda p: _abi.check(L.b200boot_select_pick(a, d, p, ws, wsz, s)),
            end=end, hist_buffer=self.hist_tensor())


def sort_columns(stat):
    """In-place ascending sort of every row of stat float64[D][n]."""
    n_cols, n = stat.shape
    ws, wsz = workspace_for(_abi.STAT_MEAN, 1, 1, n_cols, n, 1)
    _abi.check(_abi.lib().b200boot_sort_columns(stat.data_ptr(), n, n_cols, ws, wsz, stream_ptr()))
    return stat
