"""Worker for tests/test_gpu_multi.py: one of `world` GPU ranks over NCCL.  Every case is
run sharded over the ranks and again unsharded on this rank; the results must be identical."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scikits_bootstrap_b200 as boot  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(5)
    x = rng.standard_normal(50001) + 1.0
    a, b = rng.gamma(2, 1, 20000), rng.gamma(3, 1, 20000)
    table = rng.standard_normal((500, 37))
    tall = rng.standard_normal((29001, 3))                  # beyond the resident rows: row-space medians (K1mr)
    long_rows = rng.gamma(2.0, 1.0, 600_000)                # numpy rows copied under the occupancy chains
    wide = rng.standard_normal((300, 80))                   # many short columns: prefix counts on the tensor cores (K1mg)
    # pinned rows above _device.SHARD_COPY_BYTES: every rank copies 1/G of them, all-gather over NVLink
    # (odd length: the last rank's part is padded)
    pinned = torch.from_numpy(rng.standard_normal(4_500_001) + 2.0).pin_memory()
    holes = x[:3000].copy()
    holes[17] = np.nan                                      # numpy's NaN rule for rank statistics
    cases = [
        lambda: boot.ci(x, np.mean, n_samples=1003, seed=3),
        lambda: boot.ci(x, np.std, n_samples=1003, seed=3, alpha=(0.05, 0.5, 0.95), method="pi"),
        lambda: boot.ci((a, b), boot.pearson_r, n_samples=777, seed=4),
        lambda: boot.ci(table, boot.median_axis0, n_samples=501, seed=5),
        lambda: boot.ci(table, boot.mean_axis0, n_samples=501, seed=5, output="errorbar"),
        lambda: boot.ci(x, np.median, n_samples=333, seed=6),
        lambda: boot.ci((a[:60], b[:45]), boot.diff_of_means, n_samples=999, seed=7, multi="independent"),
        lambda: boot.pval(table, boot.mean_axis0, boot.greater_than(0.01), n_samples=501, seed=8),
        lambda: boot.pval(x, np.mean, lambda s: s > 1.0, n_samples=501, seed=8),
        lambda: boot.ci(x[:2000], np.mean, n_samples=801, seed=9, return_dist=True)[1],
        lambda: boot.ci((a, b), boot.linear_fit, n_samples=555, seed=10),
        lambda: boot.ci(x, boot.percentile(90), n_samples=333, seed=11),
        lambda: boot.ci(table, boot.quantile_axis0(0.25), n_samples=301, seed=12),
        lambda: boot.ci((a[:6000], x), boot.diff_of(np.median), n_samples=401, seed=13, multi="independent"),
        lambda: boot.ci(tall, boot.median_axis0, n_samples=90, seed=14, method="pi"),
        lambda: boot.ci(tall, boot.quantile_axis0(0.7), n_samples=64, seed=15, return_dist=True, method="pi")[1],
        lambda: boot.ci(long_rows, np.mean, n_samples=40000, seed=16),       # single-launch tail over 4e4 values
        lambda: boot.ci(long_rows, np.std, n_samples=301, seed=17, method="pi"),
        lambda: boot.ci(holes, np.median, n_samples=201, seed=18, method="pi", return_dist=True)[1],
        lambda: boot.ci(x[:1000], np.mean, n_samples=2000, seed=19),         # small-call shape (unsharded: one kernel)
        lambda: boot.ci(wide, boot.median_axis0, n_samples=403, seed=20),
        lambda: boot.pval(wide, boot.quantile_axis0(0.4), boot.less_than(-0.2), n_samples=403, seed=21),
        lambda: boot.ci(pinned, np.mean, n_samples=302, seed=22),
        lambda: boot.ci(pinned, np.median, n_samples=101, seed=23, method="pi"),
        lambda: boot.ci(pinned.numpy(), np.std, n_samples=203, seed=24),                  # numpy rows, same split copy
        lambda: boot.ci((pinned.numpy(), pinned.numpy()[::-1].copy()), boot.ratio_of_means, n_samples=99, seed=25),
        lambda: boot.ci(x, np.mean, n_samples=64, seed=None),          # seed agreed through a broadcast
    ]
    boot._device.SHARD_COPY_MIN_RANKS = 2         # (8 by default: exercise the split copy on any box)
    boot.distributed.enable()
    sharded = [np.asarray(c()) for c in cases[:-1]]
    # small statistic vectors are gathered and finished locally; with the limit at zero the same
    # calls go through the all-reduced counts and radix histograms instead
    limit = boot.bootstrap._GATHER_LIMIT_BYTES
    boot.bootstrap._GATHER_LIMIT_BYTES = 0
    reduced = [np.asarray(c()) for c in cases[:-1]]
    boot.bootstrap._GATHER_LIMIT_BYTES = limit
    for i, (got_gathered, got_reduced) in enumerate(zip(sharded, reduced)):      # (not a, b: the cases close over those)
        np.testing.assert_array_equal(got_gathered, got_reduced, err_msg=f"gather vs reduce, case {i}")
    free = np.asarray(cases[-1]())
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, free.tolist())
    assert all(g == gathered[0] for g in gathered), gathered
    boot.distributed.disable()
    for i, c in enumerate(cases[:-1]):
        alone = np.asarray(c())
        np.testing.assert_array_equal(sharded[i], alone, err_msg=f"case {i}")
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {local} ok")


if __name__ == "__main__":
    main()
