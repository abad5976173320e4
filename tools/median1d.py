"""1-D median calls for launch lists (development aid)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
x = np.random.default_rng(0).standard_normal(n)
for i in range(3):
    boot.ci(x, np.median, n_samples=10000, seed=i)
torch.cuda.synchronize()
