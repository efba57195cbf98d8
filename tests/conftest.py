import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def krr_test_data(seed=0, n=2000, p=100, d=7, n_valid=50):
    """The reference's KRR test distribution (/root/reference/tests/test_krr_approx.py:22-44), seeded."""
    g = torch.Generator().manual_seed(seed)
    X = torch.normal(0, 1, size=(n, p), generator=g)
    Xv = torch.normal(0, 1, size=(n_valid, p), generator=g)
    W = torch.normal(0, 1, size=(p, d), generator=g)
    return X, Xv, W, X.matmul(W), Xv.matmul(W)


def relerr(x, ref):
    x = torch.as_tensor(np.asarray(x)).double() if not torch.is_tensor(x) else x.detach().cpu().double()
    ref = torch.as_tensor(np.asarray(ref)).double() if not torch.is_tensor(ref) else ref.detach().cpu().double()
    den = float(ref.abs().max())
    return float((x - ref).abs().max() / (den if den > 0 else 1.0))


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need CUDA and the built library: skip them (instead of erroring) where either is missing."""
    from sobolev_alignment_b200 import _native as nat

    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    elif not os.path.exists(nat.library_path()):
        reason = "libsobolev_b200.so not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def sab_option():
    """Set library options (SAB_*) for one test; restored to the environment's values afterwards."""
    from sobolev_alignment_b200 import _native as nat

    touched = []

    def _set(name, value):
        touched.append(name)
        nat.set_option(name, value)

    yield _set
    for name in touched:
        nat.set_option(name, None)


@pytest.fixture(scope="session")
def cuda_ops():
    from sobolev_alignment_b200.engine import CudaOps

    return CudaOps()
