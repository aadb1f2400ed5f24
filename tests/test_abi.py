"""CPU: the C-ABI library loads and exports every symbol include/uda_b200.h declares; argument
validation rejects bad calls before any launch; the C oracle builds and matches the numpy oracle."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT, GOLDEN
from uda_aerial_semantic_segmentation_research_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `make` (or __graft_entry__.build()) first"
    lib = _lib.lib()
    names = _lib.header_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/uda_b200.h but not exported: {missing}"
    assert lib.uda_abi_version() == 1


def test_argument_validation_happens_before_launch():
    # null pointers / bad shapes must fail with a message and never reach a kernel launch (works without a GPU)
    with pytest.raises(_lib.UdaError, match="null"):
        _lib.call("seg_loss_fwd_bwd", None, _lib.ci(0), None, None, None, None, None, None, _lib.ci(1), _lib.ci(1),
                  _lib.ll(1), _lib.ci(1), _lib.ci(0), 0.25, 2.0, _lib.ci(1), _lib.ll(-100), 1.0, 1.0, 1.0, 1.0, None)
    with pytest.raises(_lib.UdaError):
        _lib.call("conv2d_direct_fwd", None, _lib.ci(0), None, _lib.ci(0), None, None, None, _lib.ci(1), _lib.ci(8),
                  _lib.ci(8), _lib.ci(3), _lib.ci(4), _lib.ci(3), _lib.ci(3), _lib.ci(1), _lib.ci(1), None)
    assert _lib.lib().uda_last_error()


def test_no_cpu_fallback():
    import torch
    from uda_aerial_semantic_segmentation_research_b200 import Unet
    from uda_aerial_semantic_segmentation_research_b200.losses import DiceLoss
    m = Unet("resnet34", classes=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DiceLoss()(torch.zeros(1, 4, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long))


def test_c_oracle_matches_numpy_oracle_and_golden():
    so = os.path.join(ROOT, "oracle", "_build", "libuda_oracle.so")
    assert os.path.exists(so), "run `make oracle`"
    lib = ctypes.CDLL(so)
    d = np.load(os.path.join(GOLDEN, "losses_c24_blocky.npz"))
    z = np.ascontiguousarray(d["z1"], dtype=np.float32)
    B, C, H, W = z.shape
    mask = np.zeros((B, H, W), dtype=np.int64)
    lib.oracle_argmax(z.ctypes.data_as(ctypes.c_void_p), mask.ctypes.data_as(ctypes.c_void_p), B, C,
                      ctypes.c_int64(H * W))
    assert np.array_equal(mask, d["argmax"])
    hist = np.zeros((C, C), dtype=np.int64)
    t = np.ascontiguousarray(d["target"], dtype=np.int64)
    lib.oracle_fast_hist(mask.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p),
                         hist.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(mask.size), C, ctypes.c_int64(0), 0)
    assert np.array_equal(hist, d["hist"])
    lib.oracle_fast_hist(mask.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p),
                         hist.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(mask.size), C, ctypes.c_int64(0), 1)
    assert np.array_equal(hist, d["hist_ignore0"])
