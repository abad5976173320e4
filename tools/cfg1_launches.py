"""Three cfg1 calls (for an ncu launch list of the small-call path; development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
x = np.random.default_rng(0).standard_normal(1000)
xd = torch.from_numpy(x).cuda()
for i in range(3):
    print(boot.ci(xd, np.mean, n_samples=10000, seed=i))
torch.cuda.synchronize()
