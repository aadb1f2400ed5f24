"""Oracle restatement of the evaluation path — TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy integer arithmetic, bit-exact by construction.  Follows
``src/analysis/metrics.py:17-42`` (confusion matrix via bincount(C*true+pred)) and
``src/models/predict.py:113-130`` (argmax over dim 1, first maximal index; NaN is
treated as maximal, as torch.argmax does).
"""
import numpy as np


def argmax_mask(logits):
    """predict_batch's ``outputs.argmax(dim=1)`` — src/models/predict.py:129. logits [B,C,H,W] -> int64 [B,H,W]."""
    z = np.asarray(logits, dtype=np.float32)
    B, C, H, W = z.shape
    best = z[:, 0].copy()
    idx = np.zeros((B, H, W), dtype=np.int64)
    for c in range(1, C):
        v = z[:, c]
        # strictly-greater keeps the first max; a NaN beats any non-NaN and the first NaN wins
        take = (v > best) | (np.isnan(v) & ~np.isnan(best))
        best = np.where(take, v, best)
        idx = np.where(take, c, idx)
    return idx


def fast_hist(pred, true, num_classes, ignore_index=None):
    """SegmentationMetrics._fast_hist — src/analysis/metrics.py:17-27. rows=true, cols=pred, int64."""
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    true = np.asarray(true).reshape(-1).astype(np.int64)
    mask = (true >= 0) & (true < num_classes)
    if ignore_index is not None:
        mask &= true != ignore_index
    return np.bincount(num_classes * true[mask] + pred[mask],
                       minlength=num_classes ** 2).reshape(num_classes, num_classes)


def batch_iou(pred, true, num_classes, ignore_index=None):
    """SegmentationMetrics.batch_iou — src/analysis/metrics.py:29-42."""
    hist = fast_hist(pred, true, num_classes, ignore_index)
    d = np.diag(hist)
    iu = d / (hist.sum(axis=1) + hist.sum(axis=0) - d + 1e-7)
    return {"mean_iou": np.nanmean(iu), "class_iou": {i: v for i, v in enumerate(iu)}}


def pixel_accuracy(pred, true, ignore_index=None):
    """SegmentationMetrics.pixel_accuracy — src/analysis/metrics.py:44-49."""
    pred, true = np.asarray(pred), np.asarray(true)
    mask = (true != ignore_index) if ignore_index is not None else np.ones_like(true, dtype=bool)
    return int(((pred == true) & mask).sum()) / (int(mask.sum()) + 1e-7)


def f1_scores(pred, true, num_classes, ignore_index=None):
    """SegmentationMetrics.f1_score (all classes) — src/analysis/metrics.py:51-67."""
    hist = fast_hist(pred, true, num_classes, ignore_index)
    tp = np.diag(hist)
    fp = hist.sum(axis=0) - tp
    fn = hist.sum(axis=1) - tp
    return 2 * tp / (2 * tp + fp + fn + 1e-7)


def trainer_metrics_from_hist(hist):
    """The per-step metrics of ``SegmentationTrainer.calculate_metrics`` (``src/models/train.py:225-243``) restated on
    the confusion matrix (rows = true, cols = pred).  ``iou``: ``torchmetrics.JaccardIndex(task='multiclass',
    num_classes=C)`` — third-party, unpinned in requirements.txt and absent here (PARITY UNPINNED); its published
    algorithm (torchmetrics >= 1.0, ``_jaccard_index_reduce`` with average='macro'): per-class diag / (row + col -
    diag) with 0 for 0/0, averaged over the classes whose row + column is non-zero.  ``accuracy``: mean(pred == mask).
    ``iou_class_c``: binary JaccardIndex of (pred == c, mask == c) = tp / (tp + fp + fn), 0 for 0/0."""
    hist = np.asarray(hist, dtype=np.float64)
    d = np.diag(hist)
    row, col = hist.sum(axis=1), hist.sum(axis=0)
    union = row + col - d
    jac = np.where(union > 0, d / np.where(union > 0, union, 1.0), 0.0)
    support = (row + col) > 0
    return {"iou": float(jac[support].mean()) if support.any() else 0.0,
            "accuracy": float(d.sum() / hist.sum()) if hist.sum() > 0 else 0.0,
            "iou_per_class": jac,
            "mean_iou": float(np.nanmean(d / (union + 1e-7))), "class_iou": d / (union + 1e-7)}
