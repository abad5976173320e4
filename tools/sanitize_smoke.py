"""Small invocation of every kernel family, for compute-sanitizer runs (development aid)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
x = rng.standard_normal(40001)            # 2 full tiles + odd remainder
a, b = rng.gamma(2, 1, 20011), rng.gamma(3, 1, 20011)
print(boot.ci(x, np.mean, n_samples=70, seed=1))
print(boot.ci(x, np.std, n_samples=70, seed=1, return_dist=True)[0])
print(boot.ci((a, b), boot.pearson_r, n_samples=70, seed=1))
print(boot.ci((a, b), boot.ratio_of_means, n_samples=70, seed=1, method="pi"))
print(boot.ci(x[:999], np.median, n_samples=70, seed=1))
print(boot.ci(x, np.median, n_samples=70, seed=1))
t = rng.standard_normal((301, 37))
print(boot.ci(t, boot.median_axis0, n_samples=70, seed=1)[:, :2])
print(boot.ci(t, boot.std_axis0, n_samples=70, seed=1)[:, :2])
print(boot.pval(t, boot.mean_axis0, n_samples=70, seed=1)[:3])
print(boot.ci((a[:60], b[:45]), boot.diff_of_means, n_samples=70, seed=1, multi="independent"))
print(len(list(boot.bootstrap_indices(x, 3, seed=1))), boot.subsample_indices(x[:50], 5, 10, seed=1).shape,
      len(list(boot.bootstrap_indices_moving_block(x[:50], 3, seed=1))))
idx = np.array(list(boot.bootstrap_indices(x[:500], 40, seed=2)))
print(boot.ci_indexed(x[:500], np.var, idx), boot.ci_indexed(x[:500], np.median, idx))
print(boot.ci(x[:20], method="abc"))
# round-2 paths: fused small call (numpy and CUDA input), single-launch tail, row-space long
# medians, K6i pipeline (resident and streamed count tile), moving-block plan, generic independent
import torch
print(boot.ci(x[:1000], np.mean, n_samples=2000, seed=3), boot.ci(torch.from_numpy(x[:1000]).cuda(), np.std, n_samples=300, seed=3))
print(boot.ci((a[:3000], b[:3000]), boot.pearson_r, n_samples=500, seed=3))
print(boot.ci(x, np.mean, n_samples=40000, seed=3, alpha=(0.01, 0.5, 0.99)))
big = rng.standard_normal((29000, 5))
print(boot.ci(big, boot.median_axis0, n_samples=40, seed=1, method="pi")[:, :2])
print(boot.ci(big, boot.quantile_axis0(0.3), n_samples=40, seed=1, method="pi")[:, :2])
wide = rng.standard_normal((1000, 300))
print(boot.ci(wide, boot.mean_axis0, n_samples=300, seed=1, method="pi")[:, :2])
tall = rng.standard_normal((1500, 40))
print(boot.ci(tall, boot.std_axis0, n_samples=200, seed=1, method="pi")[:, :2])
print(boot.ci_moving_block(x[:5000], np.mean, n_samples=100, block_length=7, seed=1))
print(boot.ci((a[:600], b[:450]), boot.independent("div", np.mean, np.median), n_samples=200, seed=1, multi="independent"))
