"""Multi-GPU sharding of the resample range (one process per GPU).

Resample ids are independent units: rank g of G owns the contiguous id range
[g*n/G, (g+1)*n/G) — a contiguous Philox counter range — and the data is
replicated.  A small statistic vector (the usual case) is all-gathered once and
every rank finishes on the whole vector, which is the unsharded run's by
construction; a large one stays sharded and the only exchanges are integer
all-reduces (SUM): the z0 / pval counts and the per-pass radix histograms of the
order-statistic select, plus an all-gather of the shards when the caller asks for
the distribution.  Integer sums are order-independent, so either way results are
bit-identical for any G."""
from __future__ import annotations

from typing import Optional, Tuple

_group = None
_enabled = False


def enable(group=None) -> None:
    """Shard subsequent ci()/pval() calls over the (default) process group.
    Every rank must make the same calls with the same data and seed."""
    import torch.distributed as dist
    global _group, _enabled
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable() -> None:
    global _group, _enabled
    _group, _enabled = None, False


def active() -> bool:
    return _enabled


def world() -> Tuple[int, int]:
    """(rank, world_size) of the sharding group; (0, 1) when not sharded."""
    if not _enabled:
        return 0, 1
    import torch.distributed as dist
    return dist.get_rank(_group), dist.get_world_size(_group)


def backend() -> str:
    """Backend of the sharding group ('nccl', 'gloo'); '' when not sharded."""
    if not _enabled:
        return ""
    import torch.distributed as dist
    return dist.get_backend(_group)


def all_gather_equal(out, part) -> None:
    """out[r * len(part) : (r + 1) * len(part)] = rank r's `part`, on the current stream."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(out, part, group=_group)


def shard_range(n_samples: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous resample-id range of `rank`: [rank*n/G, (rank+1)*n/G)."""
    return (rank * n_samples) // world_size, ((rank + 1) * n_samples) // world_size


def all_reduce_sum(t) -> None:
    if _enabled:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_group)


def radix_select_passes(begin, hist, pick, end, hist_buffer=None):
    """Drives the 8-pass MSB radix descent of K5: begin(); then per pass hist(p) on the
    local shard, SUM all-reduce of the histogram buffer across ranks, pick(p); end().
    The callables wrap the C-ABI steps (b200boot_select_begin/hist/pick/end)."""
    begin()
    for p in range(8):
        hist(p)
        if hist_buffer is not None:
            all_reduce_sum(hist_buffer)
        pick(p)
    return end()


def all_gather_ragged(t, n_total: int, dim: int = -1):
    """Concatenates the ranks' shards (contiguous id ranges) along the last dim."""
    if not _enabled:
        return t
    import torch
    import torch.distributed as dist
    rank, ws = world()
    sizes = [shard_range(n_total, r, ws)[1] - shard_range(n_total, r, ws)[0] for r in range(ws)]
    width = max(sizes)
    if (min(sizes) == width and t.numel() == width and t.is_contiguous()
            and dist.get_backend(_group) == "nccl"):
        # equal shards of one row (the usual case: a scalar statistic, n divisible by the ranks):
        # the ranks' shards land side by side in the result, no padding, no list of parts, no cat
        out = torch.empty(t.shape[:-1] + (n_total,), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.view(-1), group=_group)
        return out
    pad = torch.zeros(t.shape[:-1] + (width,), dtype=t.dtype, device=t.device)
    pad[..., : t.shape[-1]] = t
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad, group=_group)
    return torch.cat([p[..., :s] for p, s in zip(parts, sizes)], dim=-1)


def broadcast_seed(key: int) -> int:
    """Makes rank 0's 64-bit key the key of every rank (used when seed=None)."""
    if not _enabled:
        return key
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(_group) == "nccl" else "cpu"
    t = torch.tensor([key - (1 << 63)], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(_group, 0) if _group is not None else 0, group=_group)
    return int(t.item()) + (1 << 63)
