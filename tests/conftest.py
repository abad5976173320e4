"""pytest configuration: registers the `gpu` marker and puts the repo root,
`oracle/` and `tests/golden/` on sys.path (oracle imports are allowed in tests)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the oracle's restatement of the device streams follows the round count of the built library
    try:
        import bootstrap_oracle as orc
        from scikits_bootstrap_b200 import _abi
        orc.PHILOX_ROUNDS = int(_abi.lib().b200boot_philox_rounds())
    except Exception:  # library not built: the tests that need it fail on their own
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_cases.npz"))
