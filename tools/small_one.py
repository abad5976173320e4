import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import scikits_bootstrap_b200 as boot
x = torch.from_numpy(np.random.default_rng(0).standard_normal(1000)).cuda()
for i in range(6): boot.ci(x, np.mean, n_samples=10000, seed=i)
torch.cuda.synchronize()
