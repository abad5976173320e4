"""Worker for tests/test_dist_gloo.py: one of `world` CPU ranks over gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import bootstrap_oracle as orc  # noqa: E402
import scikits_bootstrap_b200 as boot  # noqa: E402
from scikits_bootstrap_b200 import _dist  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    boot.distributed.enable()
    assert _dist.world() == (rank, world)

    # seed agreement when seed=None
    key = _dist.broadcast_seed(1000 + rank * 17 + (1 << 63))
    assert key == 1000 + (1 << 63)
    key = boot.bootstrap._seed_key(None)
    keys = [None] * world
    dist.all_gather_object(keys, key)
    assert len(set(keys)) == 1

    # shards partition the resample range; ragged all-gather restores id order
    n = 1003
    lo, hi = _dist.shard_range(n, rank, world)
    full = torch.arange(2 * n, dtype=torch.float64).reshape(2, n)
    got = _dist.all_gather_ragged(full[:, lo:hi].contiguous(), n)
    assert torch.equal(got, full)

    # z0 / pval count: integer SUM all-reduce
    rng = np.random.default_rng(7)
    stat = rng.standard_normal(n)
    stat[::50] = 0.25            # ties
    c = torch.tensor([int((stat[lo:hi] < 0.25).sum())])
    _dist.all_reduce_sum(c)
    assert int(c) == int((stat < 0.25).sum())

    # sharded order-statistic select: per-pass histogram all-reduce
    ranks = [0, 1, 25, 501, n - 1]
    st = orc.RadixSelectState(stat[lo:hi], ranks)
    hist_t = torch.from_numpy(st.hist)          # shares memory with st.hist
    out = _dist.radix_select_passes(begin=lambda: None, hist=st.do_hist, pick=st.do_pick,
                                    end=st.result, hist_buffer=hist_t)
    np.testing.assert_array_equal(out, np.sort(stat)[ranks])

    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
