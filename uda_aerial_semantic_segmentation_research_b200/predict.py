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


class GraphedForward:
    """Eval-mode forward of a uda_b200 network captured in a CUDA graph (inference is ~70 launches of 10-40 us each:
    issued one by one from Python the host, not the B200, bounds the window rate).  ``fwd = GraphedForward(model,
    example); logits = fwd(x)`` — ``x`` must have the example's shape; the returned logits buffer is reused by the next
    call.  The folded BatchNorm weights are baked in: build a new object after the model's weights change."""

    def __init__(self, model, example, warmup=2):
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input")
        model.eval()
        self.model = model
        self.x = example.clone()
        s = torch.cuda.Stream(device=example.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):          # fills the folded-weight cache, sizes the workspaces
                model(self.x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._keep = dict(model._store._fold_cache)     # the captured launches read these tensors
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = model(self.x)

    def __call__(self, x=None):
        """``x`` None: the caller has already written the input into ``self.x`` (e.g. a gather kernel's output)."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y


#: ImageNet statistics — the usual ``Config.NORMALIZE_MEAN / NORMALIZE_STD`` (the reference's config module is missing)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@torch.no_grad()
def sliding_window_evaluate_u8(model, tile_u8, target, num_classes, window=512, batch=16, ignore_index=None,
                               mean=IMAGENET_MEAN, std=IMAGENET_STD, return_mask=False, graphed=None):
    """Config 5 from the RAW tile: ``tile_u8`` is the uint8 [H,W,3] image as it comes from disk, ``target`` an int64 or
    uint8 [H,W] label map.  Per batch of windows: one gather + ToTensor + Normalize kernel (reference
    ``src/models/predict.py:93-97``), the conv-only eval forward, one fused argmax + confusion-matrix kernel (uint8
    masks scattered back into the tile mask).  The metrics are derived on the device from the resident histogram
    (``SegmentationMetrics.device_metrics``); nothing synchronises with the host.  ``graphed``: a ``GraphedForward`` of
    ``model`` for ``[batch,3,window,window]`` inputs — the windows are then gathered straight into its input buffer and
    the forward is one graph replay.
    Returns ``{'hist', 'metrics' (device tensors)[, 'mask': uint8 [H,W]]}``."""
    model.eval()
    if not (tile_u8.is_cuda and tile_u8.dtype == torch.uint8 and tile_u8.dim() == 3 and tile_u8.shape[2] == 3):
        raise ValueError("sliding_window_evaluate_u8: tile must be a CUDA uint8 [H,W,3] tensor")
    H, W = tile_u8.shape[:2]
    if H % window or W % window:
        raise ValueError("tile size must be a multiple of the window")
    n_win = (H // window) * (W // window)
    hist = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=tile_u8.device)
    tile_mask = torch.empty((H, W), dtype=torch.uint8, device=tile_u8.device) if return_mask else None
    target = target.contiguous()
    for first in range(0, n_win, batch):
        n = min(batch, n_win - first)
        t = ops.gather_label_windows(target, window, first, n)
        if graphed is not None and n == graphed.x.shape[0]:
            ops.gather_windows_u8(tile_u8, window, first, n, mean, std, out=graphed.x)
            logits = graphed()
        else:
            logits = model(ops.gather_windows_u8(tile_u8, window, first, n, mean, std))
        m, _ = ops.argmax_confmat(logits.contiguous(), t, num_classes=num_classes, ignore_index=ignore_index,
                                  want_mask=return_mask, mask_dtype=torch.uint8, hist=hist)
        if return_mask:
            ops.scatter_window_masks(m, tile_mask, window, first)
    out = {"hist": hist, "metrics": SegmentationMetrics.device_metrics(hist)}
    if return_mask:
        out["mask"] = tile_mask
    return out


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
