"""Times the five BASELINE.json configs through the public API (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scikits_bootstrap_b200 as boot


def run(name, points, fn, reps=3):
    fn(0)
    best = 1e9
    for i in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = fn(i + 1); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    print(f"{name:5s} {best*1e3:9.2f} ms  {points/best:.3e} points/s  {np.asarray(out).ravel()[:4]}")


def main():
    dev = lambda a: torch.from_numpy(a).cuda()
    x1 = dev(np.random.default_rng(0).standard_normal(1000))
    x2 = dev(np.random.default_rng(1).standard_normal(100000))
    a = dev(np.random.default_rng(2).gamma(2, 1, 10**6)); b = dev(np.random.default_rng(3).gamma(3, 1, 10**6))
    x4 = dev(np.random.default_rng(4).standard_normal((1000, 10000)))
    x5 = dev(np.random.default_rng(5).standard_normal(10**7) + 1.0)
    run("cfg1", 1e7, lambda s: boot.ci(x1, np.mean, method="bca", n_samples=10000, seed=s))
    run("cfg2", 1e10, lambda s: boot.ci(x2, np.average, method="pi", n_samples=100000, seed=s))
    run("cfg3", 1e11, lambda s: boot.ci((a, b), boot.ratio_of_means, method="bca", n_samples=100000, seed=s, multi="paired"))
    run("cfg4", 1e11, lambda s: boot.ci(x4, boot.median_axis0, method="bca", n_samples=10000, seed=s))
    run("cfg4p", 1e11, lambda s: boot.pval(x4, boot.median_axis0, n_samples=10000, seed=s))
    run("cfg5", 1e12, lambda s: boot.ci(x5, np.mean, method="bca", n_samples=100000, seed=s), reps=2)


if __name__ == "__main__":
    main()
