"""GPU parity of the prediction / evaluation path: argmax masks and confusion matrices are BIT-EXACT
(north-star) against the golden vectors from the reference's own SegmentationMetrics, the numpy oracle,
and property checks at BASELINE cfg5 size (one 4096x4096 tile = 16.8 M pixels)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import ref_metrics as M

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(GOLDEN, "losses_*.npz")))


def _mods():
    from uda_aerial_semantic_segmentation_research_b200 import metrics, predict, ops
    return metrics, predict, ops


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(c) for c in CASES])
def test_golden_vectors(path):
    metrics, predict, ops = _mods()
    d = np.load(path)
    dev = torch.device("cuda:0")
    z = torch.from_numpy(d["z1"]).to(dev)
    t = torch.from_numpy(d["target"]).to(dev)
    C = z.shape[1]
    mask = predict.argmax_mask(z)
    assert mask.dtype == torch.int64 and np.array_equal(mask.cpu().numpy(), d["argmax"])
    sm = metrics.SegmentationMetrics(C)
    assert np.array_equal(sm._fast_hist(mask.flatten(), t.flatten()), d["hist"])
    assert np.array_equal(metrics.SegmentationMetrics(C, ignore_index=0)._fast_hist(mask.flatten(), t.flatten()), d["hist_ignore0"])
    m2, h2 = metrics.logits_confusion_matrix(z, t)
    assert np.array_equal(h2.cpu().numpy(), d["hist"]) and torch.equal(m2, mask)
    iou = sm.batch_iou(mask, t)
    assert abs(iou["mean_iou"] - float(d["mean_iou"])) < 1e-12
    assert np.allclose([iou["class_iou"][i] for i in range(C)], d["class_iou"], atol=1e-12)
    assert abs(sm.pixel_accuracy(mask, t) - float(d["pixel_acc"])) < 1e-9
    assert np.allclose(sm.f1_score(mask, t), d["f1"], atol=1e-12)
    # bf16 logits: bit-exact against the oracle evaluated on the same (rounded) logits
    zb = z.bfloat16()
    assert np.array_equal(predict.argmax_mask(zb).cpu().numpy(), M.argmax_mask(zb.float().cpu().numpy()))
    m8 = predict.argmax_mask(z, mask_dtype=torch.uint8)
    assert m8.dtype == torch.uint8 and np.array_equal(m8.cpu().numpy().astype(np.int64), d["argmax"])


def test_ties_nan_and_invalid_targets():
    metrics, predict, ops = _mods()
    dev = torch.device("cuda:0")
    z = torch.zeros(1, 6, 5, 7)                      # ragged (scalar path); all ties -> index 0
    z[0, 3, 1, 1] = 1.0; z[0, 5, 1, 1] = 1.0         # tie between 3 and 5 -> first (3)
    z[0, 2, 2, 2] = float("nan"); z[0, 4, 2, 2] = float("nan")  # NaN is maximal, first NaN wins
    z[0, :, 3, 3] = float("-inf")
    ref = M.argmax_mask(z.numpy())
    assert np.array_equal(ref, z.argmax(1).numpy())   # oracle == torch semantics
    assert np.array_equal(predict.argmax_mask(z.to(dev)).cpu().numpy(), ref)
    pred = torch.tensor([0, 1, 2, 3, 4, 5, 0, 1], device=dev)
    true = torch.tensor([0, 1, -1, 6, 255, 5, 2, 1], device=dev)   # out-of-range targets are skipped
    h = metrics.SegmentationMetrics(6)._fast_hist(pred, true)
    assert np.array_equal(h, M.fast_hist(pred.cpu().numpy(), true.cpu().numpy(), 6)) and h.sum() == 5
    # empty input
    e = torch.zeros(0, dtype=torch.int64, device=dev)
    assert metrics.SegmentationMetrics(6)._fast_hist(e, e).sum() == 0


@pytest.mark.parametrize("C", [2, 24, 23, 64, 150])
def test_random_against_oracle(C):
    metrics, predict, ops = _mods()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C)
    z = torch.randn(2, C, 48, 64, generator=g).bfloat16().float()   # bf16-valued: ties are common
    t = torch.randint(0, C, (2, 48, 64), generator=g)
    mask, hist = metrics.logits_confusion_matrix(z.to(dev), t.to(dev))
    ref = M.argmax_mask(z.numpy())
    assert np.array_equal(mask.cpu().numpy(), ref)
    assert np.array_equal(hist.cpu().numpy(), M.fast_hist(ref, t.numpy(), C))
    # blocky labels (warp-uniform fast path)
    tb = torch.randint(0, C, (2, 3, 4), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    _, hb = metrics.logits_confusion_matrix(z.to(dev), tb.to(dev))
    assert np.array_equal(hb.cpu().numpy(), M.fast_hist(ref, tb.numpy(), C))


def test_full_tile_properties():
    """cfg5: 64 windows of 512x512, 24 classes.  Checksum properties instead of a CPU oracle run:
    the histogram's total equals the pixel count, its column sums equal the mask's class counts, its row
    sums the targets', the fused kernel equals the two-step (argmax, then confmat) path, and accumulating
    window batches equals one pass over everything."""
    metrics, predict, ops = _mods()
    dev = torch.device("cuda:0")
    C = 24
    g = torch.Generator(device=dev).manual_seed(5)
    hist_acc = torch.zeros(C, C, dtype=torch.int64, device=dev)
    col = torch.zeros(C, dtype=torch.int64, device=dev)
    row = torch.zeros(C, dtype=torch.int64, device=dev)
    for i in range(4):
        z = torch.randn(16, C, 512, 512, device=dev, generator=g)
        t = torch.randint(0, C, (16, 512, 512), device=dev, generator=g)
        mask, _ = metrics.logits_confusion_matrix(z, t, hist=hist_acc)
        assert torch.equal(mask, z.argmax(1))   # torch on the same device, identical logits
        h2 = metrics.SegmentationMetrics(C).hist_tensor(mask, t)
        _, h1 = metrics.logits_confusion_matrix(z, t)
        assert torch.equal(h1, h2)
        col += torch.bincount(mask.flatten(), minlength=C)
        row += torch.bincount(t.flatten(), minlength=C)
    assert hist_acc.sum().item() == 64 * 512 * 512
    assert torch.equal(hist_acc.sum(0), col) and torch.equal(hist_acc.sum(1), row)


@pytest.mark.parametrize("C", [2, 24, 150])
def test_device_metrics_from_resident_histogram(C):
    """f2: IoU / accuracy / macro Jaccard derived on the device from the int64 [C,C] histogram (no host numpy) against
    the numpy restatements (in-tree SegmentationMetrics formulas and the torchmetrics-semantics macro Jaccard)."""
    metrics, predict, ops = _mods()
    rng = np.random.default_rng(C)
    h = rng.integers(0, 1000, size=(C, C)).astype(np.int64)
    if C > 2:
        h[3, :] = 0; h[:, 3] = 0                 # a class that occurs nowhere: ignored by the macro Jaccard
        h[5, :] = 0                              # a class that is only predicted: union > 0, tp = 0
    for hist in (h, np.zeros((C, C), np.int64)):
        out = metrics.SegmentationMetrics.device_metrics(torch.from_numpy(hist).to("cuda:0"))
        ref = M.trainer_metrics_from_hist(hist)
        assert abs(float(out["iou"]) - ref["iou"]) < 1e-12
        assert abs(float(out["accuracy"]) - ref["accuracy"]) < 1e-12
        assert np.allclose(out["iou_per_class"].cpu().numpy(), ref["iou_per_class"], rtol=0, atol=1e-12)
        assert np.allclose(out["class_iou"].cpu().numpy(), ref["class_iou"], rtol=0, atol=1e-12)
        assert abs(float(out["mean_iou"]) - ref["mean_iou"]) < 1e-12
        assert float(out["pixels"]) == float(hist.sum())
    # and the in-tree host formulas on the same matrix
    host = metrics.SegmentationMetrics.iou_from_hist(h)
    assert abs(float(metrics.SegmentationMetrics.device_metrics(torch.from_numpy(h).to("cuda:0"))["mean_iou"]) - host["mean_iou"]) < 1e-12


def test_u8_window_gather_normalise_and_mask_scatter():
    """f3: raw uint8 tile -> normalised fp32 NCHW windows (ToTensor + Normalize, reference predict.py:93-97), int64
    target windows, uint8 masks scattered back — against torch indexing; ragged last batch; bad windows rejected."""
    metrics, predict, ops = _mods()
    from uda_aerial_semantic_segmentation_research_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    H, W, win = 192, 320, 64
    tile = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8)
    target = torch.randint(0, 7, (H, W), generator=g)
    mean, std = predict.IMAGENET_MEAN, predict.IMAGENET_STD
    ref = ((tile.float() / 255 - torch.tensor(mean)) / torch.tensor(std)).permute(2, 0, 1)      # ToTensor + Normalize
    ref_w = predict.tile_windows(ref, win)
    ref_t = predict.tile_windows(target.reshape(1, H, W), win).reshape(-1, win, win)
    n_win = (H // win) * (W // win)
    tile_mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    fake = torch.randint(0, 7, (n_win, win, win), generator=g, dtype=torch.uint8)
    for first in range(0, n_win, 4):
        n = min(4, n_win - first)
        x = ops.gather_windows_u8(tile.to(dev), win, first, n, mean, std)
        assert rel_err(x.cpu(), ref_w[first:first + n]) < 1e-6
        for tt in (target, target.to(torch.uint8)):
            t = ops.gather_label_windows(tt.to(dev), win, first, n)
            assert torch.equal(t.cpu(), ref_t[first:first + n])
        ops.scatter_window_masks(fake[first:first + n].contiguous().to(dev), tile_mask, win, first)
    back = predict.tile_windows(tile_mask.cpu().reshape(1, H, W), win).reshape(-1, win, win)
    assert torch.equal(back, fake)
    with pytest.raises(_lib.UdaError):
        ops.gather_windows_u8(tile.to(dev), win, n_win - 1, 2, mean, std)


def test_sliding_window_from_raw_tile_is_bit_exact():
    """Config 5 from the raw uint8 tile: masks and confusion matrix equal torch.argmax + bincount on the same logits."""
    metrics, predict, ops = _mods()
    import uda_aerial_semantic_segmentation_research_b200 as U
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    model = U.Unet("resnet18", classes=9).to(dev).eval()
    g = torch.Generator().manual_seed(6)
    H, W, win = 128, 192, 64
    tile = torch.randint(0, 256, (H, W, 3), generator=g, dtype=torch.uint8).to(dev)
    target = torch.randint(0, 9, (H, W), generator=g, dtype=torch.uint8).to(dev)
    out = predict.sliding_window_evaluate_u8(model, tile, target, 9, window=win, batch=4, return_mask=True)
    n_win = (H // win) * (W // win)
    with torch.no_grad():
        x = ops.gather_windows_u8(tile, win, 0, n_win, predict.IMAGENET_MEAN, predict.IMAGENET_STD)
        full = torch.cat([model(x[i:i + 4]) for i in range(0, n_win, 4)]).argmax(1)
    full = full.reshape(H // win, W // win, win, win).permute(0, 2, 1, 3).reshape(H, W)
    assert torch.equal(out["mask"].long(), full)
    assert np.array_equal(out["hist"].cpu().numpy(), M.fast_hist(full.cpu().numpy(), target.cpu().numpy(), 9))
    ref = M.trainer_metrics_from_hist(out["hist"].cpu().numpy())
    assert abs(float(out["metrics"]["iou"]) - ref["iou"]) < 1e-12 and abs(float(out["metrics"]["accuracy"]) - ref["accuracy"]) < 1e-12
