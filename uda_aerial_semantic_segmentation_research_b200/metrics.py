"""Evaluation entry points of the reference (``src/analysis/metrics.py``) on the CUDA histogram kernel.

``SegmentationMetrics(num_classes, ignore_index)`` keeps the reference's methods and return types
(numpy confusion matrix, dict of IoUs, floats); the confusion matrix comes from the shared-memory
privatised histogram kernel and is bit-exact (integer arithmetic).
"""
from typing import List, Optional, Union

import numpy as np
import torch

from . import ops


def _idx(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: uda_b200 metrics run on CUDA tensors only (no CPU fallback)")
    t = t.reshape(-1)
    if t.dtype not in (torch.int64, torch.uint8):
        t = t.long()
    return t.contiguous()


class SegmentationMetrics:
    """Drop-in for ``src/analysis/metrics.py:5-67``."""

    def __init__(self, num_classes: int, ignore_index: Optional[int] = None):
        self.num_classes = num_classes
        self.ignore_index = ignore_index

    def hist_tensor(self, pred: torch.Tensor, true: torch.Tensor, hist: Optional[torch.Tensor] = None,
                    return_bad: bool = False):
        """Device-resident int64 [C,C] confusion matrix (rows = true, cols = pred); accumulates into
        ``hist`` when given — no host synchronisation.  ``return_bad``: also return the device counter of pixels
        whose PREDICTION is outside [0, C) (the reference's ``bincount(...).reshape(C, C)`` raises on those)."""
        p = _idx(pred, "SegmentationMetrics")
        t = _idx(true, "SegmentationMetrics").long()
        h, bad = ops.confmat(p, t, self.num_classes, self.ignore_index, hist=hist)
        return (h, bad) if return_bad else h

    def _fast_hist(self, pred: torch.Tensor, true: torch.Tensor) -> np.ndarray:
        """``metrics.py:17-27``: bincount(C*true+pred) over pixels with 0 <= true < C (and != ignore)."""
        h, bad = self.hist_tensor(pred, true, return_bad=True)
        h = h.cpu().numpy()
        nbad = int(bad.item())
        if nbad:
            raise ValueError(f"SegmentationMetrics: {nbad} predictions lie outside [0, {self.num_classes}) "
                             "(the reference's bincount/reshape fails on such input)")
        return h

    def batch_iou(self, predictions: torch.Tensor, targets: torch.Tensor) -> dict:
        """``metrics.py:29-42``."""
        return self.iou_from_hist(self._fast_hist(predictions.flatten(), targets.flatten()))

    @staticmethod
    def device_metrics(hist: torch.Tensor) -> dict:
        """Per-step metrics of ``SegmentationTrainer.calculate_metrics`` (``src/models/train.py:225-243``: macro
        Jaccard, accuracy, per-class binary Jaccard — 26 ``.item()`` synchronisations per step in the reference) derived
        ON THE DEVICE from the resident int64 [C,C] confusion matrix: one tiny launch, no host synchronisation; read the
        returned tensors once per epoch.  ``mean_iou`` / ``class_iou`` follow the in-tree metric
        (``src/analysis/metrics.py:29-42``); ``iou`` follows ``torchmetrics.JaccardIndex(task='multiclass')`` (macro
        average over the classes that occur in the target or the prediction; third-party, absent here: restated from
        its published algorithm, torchmetrics >= 1.0 — parity unpinned, checked against oracle.ref_metrics)."""
        out = ops.metrics_from_hist(hist)
        C = hist.shape[0]
        return {"mean_iou": out[0], "accuracy": out[1], "iou": out[2], "pixels": out[3],
                "class_iou": out[4:4 + C], "iou_per_class": out[4 + C:4 + 2 * C]}

    @staticmethod
    def iou_from_hist(hist: np.ndarray) -> dict:
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist) + 1e-7)
        return {"mean_iou": np.nanmean(iu), "class_iou": {i: iou for i, iou in enumerate(iu)}}

    def pixel_accuracy(self, predictions: torch.Tensor, targets: torch.Tensor) -> float:
        """``metrics.py:44-49`` (two scalar reductions; plain torch ops on the device)."""
        mask = targets != self.ignore_index if self.ignore_index is not None else torch.ones_like(targets, dtype=torch.bool)
        correct = torch.sum((predictions == targets) & mask).item()
        total = torch.sum(mask).item()
        return correct / (total + 1e-7)

    def f1_score(self, predictions: torch.Tensor, targets: torch.Tensor,
                 class_index: Optional[int] = None) -> Union[float, List[float]]:
        """``metrics.py:51-67``."""
        hist = self._fast_hist(predictions.flatten(), targets.flatten())
        if class_index is not None:
            tp = hist[class_index, class_index]
            fp = hist[:, class_index].sum() - tp
            fn = hist[class_index, :].sum() - tp
            return 2 * tp / (2 * tp + fp + fn + 1e-7)
        tp = np.diag(hist)
        fp = hist.sum(axis=0) - tp
        fn = hist.sum(axis=1) - tp
        return (2 * tp / (2 * tp + fp + fn + 1e-7)).tolist()


def logits_confusion_matrix(logits, targets, ignore_index=None, hist=None):
    """Fused ``outputs.argmax(dim=1)`` + confusion matrix (``src/models/train.py:227`` + metrics): one pass
    over the logits, no int64 mask round-trip.  Returns (mask int64 [B,H,W], hist int64 [C,C])."""
    if not logits.is_cuda:
        raise RuntimeError("logits_confusion_matrix: CUDA tensors only (no CPU fallback)")
    z = logits.contiguous()
    if z.dtype not in (torch.float32, torch.bfloat16):
        z = z.float()
    return ops.argmax_confmat(z, targets.long().contiguous(), ignore_index=ignore_index, hist=hist)
