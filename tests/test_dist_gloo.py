"""Multi-rank host logic on CPU: two and three gloo ranks run tests/_gloo_worker.py
(shard partition, seed broadcast, ragged all-gather, count all-reduce and the
sharded radix-select protocol with the oracle's numpy steps)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import bootstrap_oracle as orc
from scikits_bootstrap_b200 import _dist

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "_gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert p.stdout.count("ok") == world


def test_shard_ranges_partition():
    for n in (0, 1, 7, 100000, 99999):
        for world in (1, 2, 3, 4, 8):
            edges = [_dist.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_radix_select_restatement_single_rank():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(5000)
    x[:40] = [np.nan, -0.0, 0.0, np.inf, -np.inf] * 8
    ranks = [0, 3, 20, 2500, 4999 - 8, 4999]
    st = orc.RadixSelectState(x, ranks)
    out = _dist.radix_select_passes(lambda: None, st.do_hist, st.do_pick, st.result)
    want = np.sort(x)[ranks]
    np.testing.assert_array_equal(np.isnan(out), np.isnan(want))
    np.testing.assert_array_equal(out[~np.isnan(want)], want[~np.isnan(want)])
