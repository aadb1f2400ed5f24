"""Prediction entry points (``src/models/predict.py:113-130``) and the sliding-window evaluation driver
of BASELINE config 5 (window = stride = 512, no overlap, no blending)."""
import numpy as np
import torch

from . import ops
from .metrics import SegmentationMetrics


def argmax_mask(logits, mask_dtype=torch.int64):
    """``outputs.argmax(dim=1)`` on the fused kernel (first maximal index, NaN treated as maximal)."""
    if not logits.is_cuda:
        raise RuntimeError("argmax_mask: CUDA tensors only (no CPU fallback)")
    z = logits.contiguous()
    if z.dtype not in (torch.float32, torch.bfloat16):
        z = z.float()
    mask, _ = ops.argmax_confmat(z, None, want_mask=True, mask_dtype=mask_dtype)
    return mask


def predict_batch(model, images, device="cuda"):
    """Drop-in for ``predict_batch`` (``src/models/predict.py:113-130``): eval-mode forward, argmax over
    classes, int64 numpy array ``[B,H,W]``."""
    model.eval()
    with torch.no_grad():
        images = images.to(device)
        outputs = model(images)
        return argmax_mask(outputs).cpu().numpy()


def tile_windows(tile, window=512, stride=None):
    """Split ``[C,H,W]`` (or ``[1,C,H,W]``) into non-overlapping ``window``² crops, row-major: [N,C,window,window]."""
    stride = stride or window
    if stride != window:
        raise NotImplementedError("overlapping windows / blending are not defined by the reference (SURVEY T5)")
    if tile.dim() == 4:
        tile = tile[0]
    C, H, W = tile.shape
    if H % window or W % window:
        raise ValueError("tile size must be a multiple of the window")
    t = tile.reshape(C, H // window, window, W // window, window).permute(1, 3, 0, 2, 4)
    return t.reshape(-1, C, window, window)


@torch.no_grad()
def sliding_window_evaluate(model, tile, target, num_classes, window=512, batch=16, ignore_index=None,
                            return_mask=False):
    """Config 5: predict a large tile window by window and accumulate the confusion matrix on device.

    Returns ``{'hist': int64 [C,C] tensor, 'mean_iou', 'class_iou'[, 'mask': int64 [H,W]]}``; the only
    host synchronisation is the final histogram read."""
    model.eval()
    wins = tile_windows(tile, window)
    tw = tile_windows(target.reshape(1, *target.shape[-2:]), window).reshape(-1, window, window)
    hist = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=wins.device)
    masks = []
    for i in range(0, wins.shape[0], batch):
        logits = model(wins[i:i + batch].contiguous())
        m, _ = ops.argmax_confmat(logits.contiguous(), tw[i:i + batch].contiguous().long(), num_classes=num_classes,
                                  ignore_index=ignore_index, want_mask=return_mask, hist=hist)
        if return_mask:
            masks.append(m)
    out = SegmentationMetrics.iou_from_hist(hist.cpu().numpy())
    out["hist"] = hist
    if return_mask:
        H, W = target.shape[-2:]
        m = torch.cat(masks).reshape(H // window, W // window, window, window).permute(0, 2, 1, 3).reshape(H, W)
        out["mask"] = m
    return out
