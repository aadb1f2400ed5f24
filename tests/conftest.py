import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with `-m gpu` on the B200 box")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is present, CPU tests never touch CUDA."""
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def rel_err(a, b):
    """max-norm relative error  ||a-b||_inf / (||b||_inf + eps)  (SURVEY.md 8d parity metric)."""
    import torch
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.is_cuda and b.is_cuda and a.device == b.device:   # large tensors: reduce on the device
        d = float((a.double() - b.double()).abs().max()) if a.dtype == torch.float64 or a.numel() < (1 << 24) \
            else float((a.float() - b.float()).abs().max())
        return d / (float(b.abs().max()) + 1e-30)
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def l2_err(a, b):
    import torch
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
