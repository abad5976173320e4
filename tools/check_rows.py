"""The kernels changed last (K2's transposed stores, K1mr's 128-bit rank loads and sized histograms)
on small shapes against numpy on the materialised indices (development aid; compute-sanitizer is
closed on this pool)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
rng = np.random.default_rng(0)
x = rng.standard_normal(40001)            # 2 full tiles + odd remainder
idx = np.array(list(boot.bootstrap_indices(x, 3, seed=1)))
print(idx.shape, int(idx.min()), int(idx.max()))
for d in (5, 8):
    big = rng.standard_normal((29000, d))
    got = boot.ci(big, boot.median_axis0, n_samples=6, seed=1, method="pi")
    rows = np.array(list(boot.bootstrap_indices(big[:, :1], 6, seed=1)))
    want = np.sort(np.median(big[rows], axis=1), axis=0)[np.round(5 * np.array([0.025, 0.975])).astype(int)]
    print(d, bool(np.array_equal(got, want)))
