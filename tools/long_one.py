"""One long-table median call for launch lists (development aid)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
n, d, m = (int(v) for v in sys.argv[1:4])
x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, d))).cuda()
for it in range(2):
    boot.ci(x, boot.median_axis0, n_samples=m, seed=it, method="pi")
torch.cuda.synchronize()
