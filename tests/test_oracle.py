"""CPU: the oracle restatements against the golden vectors generated from the reference's own modules
(oracle/gen_golden.py), against the reference imported live when /root/reference is present, and
against the survey's closed-form known-answer values."""
import glob
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, rel_err
from oracle import ref_losses as R
from oracle import ref_metrics as M

CASES = sorted(glob.glob(os.path.join(GOLDEN, "losses_*.npz")))


def _load(path):
    d = np.load(path)
    return d, torch.from_numpy(d["z1"]), torch.from_numpy(d["z2"]), torch.from_numpy(d["target"]), \
        torch.from_numpy(d["class_weights"])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(c) for c in CASES])
def test_losses_match_golden(path):
    d, z1, z2, t, w = _load(path)
    C = z1.shape[1]

    def check(key, fn, wrt):
        leaves = [x.clone().requires_grad_() for x in wrt]
        loss = fn(*leaves)
        grads = torch.autograd.grad(loss, leaves)
        assert abs(loss.item() - float(d[key])) <= 2e-6 * max(1.0, abs(float(d[key]))), key
        for i, g in enumerate(grads):
            assert rel_err(g, d[f"{key}_grad{i}"]) < 1e-5, (key, i)

    check("ce", lambda z: R.cross_entropy(z, t), [z1])
    check("dice", lambda z: R.dice_loss(z, t), [z1])
    check("ce_plus_dice", lambda z: R.cross_entropy(z, t) + R.dice_loss(z, t), [z1])
    check("weighted", lambda z: R.weighted_segmentation_loss(z, t, w, domain_weight=0.7), [z1])
    check("weighted_noweights_sum", lambda z: R.weighted_segmentation_loss(z, t, None, reduction="sum"), [z1])
    check("consistency", lambda a, b: R.consistency_loss(a, b, 0.5), [z1, z2])
    check("consistency_T1", lambda a, b: R.consistency_loss(a, b, 1.0), [z1, z2])
    s, dd = torch.from_numpy(d["d_src"]), torch.from_numpy(d["d_tgt"])
    check("disc_loss", lambda a, b: R.discriminator_loss(a, b), [s, dd])
    check("gen_loss", lambda a: R.generator_loss(a), [dd])
    for ep in (0, 20, 60):
        r = R.fine_tuning_loss(z1, z2, dd, ep, supervised_pred=z1, supervised_target=t)
        assert abs(r["total"].item() - float(d[f"ft_total_ep{ep}"])) <= 2e-6 * max(1, abs(float(d[f"ft_total_ep{ep}"])))
        assert r["rampup_weight"].item() == float(d[f"ft_ramp_ep{ep}"])
    # evaluation path: integer results are bit-exact
    pred = M.argmax_mask(d["z1"])
    assert np.array_equal(pred, d["argmax"])
    assert np.array_equal(M.fast_hist(pred, d["target"], C), d["hist"])
    assert np.array_equal(M.fast_hist(pred, d["target"], C, ignore_index=0), d["hist_ignore0"])
    iou = M.batch_iou(pred, d["target"], C)
    assert abs(iou["mean_iou"] - float(d["mean_iou"])) < 1e-12
    assert np.allclose([iou["class_iou"][i] for i in range(C)], d["class_iou"], atol=1e-12)
    assert abs(M.pixel_accuracy(pred, d["target"]) - float(d["pixel_acc"])) < 1e-12
    assert np.allclose(M.f1_scores(pred, d["target"], C), d["f1"], atol=1e-12)


def test_closed_form_kats():
    # SURVEY.md 8c known answers
    assert abs(R.dice_loss(torch.zeros(1, 24, 8, 8), torch.zeros(1, 8, 8, dtype=torch.long)).item() - 0.734736) < 1e-6
    z = torch.zeros(4, 1)
    assert abs(R.discriminator_loss(z, z).item() - math.log(2)) < 1e-6
    assert abs(R.generator_loss(z).item() - 1e-3 * math.log(2)) < 1e-9
    x = torch.randn(2, 5, 4, 4)
    assert abs(R.consistency_loss(x, x).item()) < 1e-7
    assert R.rampup(0) == 0 and R.rampup(20) == 0.5 and R.rampup(40) == 1 and R.rampup(99) == 1
    g = torch.ones(3, requires_grad=True)
    R.gradient_reverse_layer(g, 0.3).sum().backward()
    assert torch.allclose(g.grad, torch.full((3,), -0.3))


def test_reference_range_assertions():
    """The reference's own (range / shape) assertions: src/test_system.py:110-124,141-148,313-319,527-580."""
    torch.manual_seed(0)
    C = 24
    pred = torch.rand(4, C, 32, 32)
    t = torch.randint(0, C, (4, 32, 32))
    onehot = torch.nn.functional.one_hot(t, C).permute(0, 3, 1, 2).float()
    l = R.dice_loss(pred, onehot)
    assert l.shape == torch.Size([]) and 0 <= l <= 1
    assert R.weighted_segmentation_loss(torch.randn(4, C, 32, 32), t) >= 0
    r0 = R.fine_tuning_loss(pred, pred.flip(0), torch.rand(4, 1), 0)
    r40 = R.fine_tuning_loss(pred, pred.flip(0), torch.rand(4, 1), 40, pred, t)
    assert r0["rampup_weight"] == 0 and r40["rampup_weight"] == 1 and r40["supervised"] > 0 and r40["total"] >= 0


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference tree not present")
def test_oracle_matches_reference_live():
    sys.path.insert(0, REFERENCE)
    try:
        from src.models import losses as L
        from src.models.discriminator import DomainDiscriminator
        from src.analysis.metrics import SegmentationMetrics
    finally:
        sys.path.remove(REFERENCE)
    from oracle.ref_discriminator import RefDomainDiscriminator
    torch.manual_seed(3)
    z1, z2 = torch.randn(2, 7, 12, 20) * 4, torch.randn(2, 7, 12, 20) * 4
    t = torch.randint(0, 7, (2, 12, 20))
    w = torch.rand(7) + 0.5
    assert abs(R.dice_loss(z1, t) - L.DiceLoss()(z1, t)) < 1e-6
    assert abs(R.weighted_segmentation_loss(z1, t, w) - L.WeightedSegmentationLoss(7, w)(z1, t)) < 1e-6
    assert abs(R.consistency_loss(z1, z2) - L.ConsistencyLoss()(z1, z2)) < 1e-3  # value ~1e3, fp32
    d = torch.rand(2, 1)
    f, g = L.FineTuningLoss()(z1, z2, d, 13, z1, t), R.fine_tuning_loss(z1, z2, d, 13, z1, t)
    assert abs(f["total"] - g["total"]) < 1e-3
    D, Dr = DomainDiscriminator(), RefDomainDiscriminator()
    Dr.load_state_dict(D.state_dict())
    x = torch.randn(2, 3, 32, 32)
    assert torch.allclose(D(x), Dr(x), atol=1e-6)
    p = z1.argmax(1)
    assert np.array_equal(SegmentationMetrics(7)._fast_hist(p.flatten(), t.flatten()), M.fast_hist(p.numpy(), t.numpy(), 7))


def test_unet_oracle_structure():
    """Structure pinned by the reference's logged graph: op counts, params, keys."""
    from collections import Counter
    from oracle.ref_unet import RefUnet
    m34, m50 = RefUnet("resnet34", classes=24), RefUnet("resnet50", classes=23)
    assert sum(p.numel() for p in m34.parameters()) == 24_439_704
    c = Counter(type(x).__name__ for x in m50.modules())
    assert c["Conv2d"] == 64 and c["BatchNorm2d"] == 63
    c = Counter(type(x).__name__ for x in m34.modules())
    assert c["Conv2d"] == 47 and c["BatchNorm2d"] == 46
    keys = list(m34.state_dict().keys())
    assert keys[0] == "encoder.conv1.weight" and keys[-1] == "segmentation_head.0.bias"
    assert "decoder.blocks.0.conv1.0.weight" in keys and "decoder.blocks.4.conv2.1.running_var" in keys
    assert "decoder.blocks.0.conv1.0.bias" not in keys  # decoder convs have no bias, the head has one
    m50.eval()
    with torch.no_grad():
        assert m50(torch.zeros(1, 3, 64, 64)).shape == (1, 23, 64, 64)


@pytest.mark.parametrize("name", ["resnet18", "resnet34", "resnet50"])
def test_encoder_oracle_equals_torchvision_resnet(name):
    """Pins the U-Net oracle's encoder to the real third-party dependency: smp's ResNet encoders ARE torchvision
    ResNets without avgpool/fc.  The restatement loads torchvision's state_dict unchanged and must
    reproduce its feature maps bit for bit (train mode, so BatchNorm batch statistics and the running-stat updates are
    covered too)."""
    tv = pytest.importorskip("torchvision")
    from oracle.ref_unet import RefResNetEncoder
    torch.manual_seed(0)
    net = getattr(tv.models, name)(weights=None)
    enc = RefResNetEncoder(name)
    sd = {k: v for k, v in net.state_dict().items() if not k.startswith("fc.")}
    assert sorted(sd.keys()) == sorted(enc.state_dict().keys())
    enc.load_state_dict(sd)
    x = torch.randn(2, 3, 64, 64)
    net.train(); enc.train()
    f = enc(x)
    y = net.relu(net.bn1(net.conv1(x)))
    assert torch.equal(f[0], x) and torch.equal(f[1], y)
    y = net.maxpool(y)
    for li in range(1, 5):
        y = getattr(net, f"layer{li}")(y)
        assert torch.equal(f[li + 1], y), (name, li)
    for k, v in net.state_dict().items():            # running statistics advanced identically
        if "running" in k and not k.startswith("fc."):
            assert torch.equal(v, enc.state_dict()[k]), k
    assert enc.out_channels == ((3, 64, 64, 128, 256, 512) if name != "resnet50" else (3, 64, 256, 512, 1024, 2048))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference tree not present")
def test_domain_model_restatement_matches_reference():
    """oracle.ref_domain_model.RefDomainAdaptationModel == the reference's DomainAdaptationModel
    (src/models/domain_model.py:4-83) on the same members: forward with and without domain adaptation,
    get_features, train / eval propagation, parameters()."""
    sys.path.insert(0, REFERENCE)
    try:
        from src.models.domain_model import DomainAdaptationModel
    finally:
        sys.path.remove(REFERENCE)
    from oracle.ref_domain_model import RefDomainAdaptationModel
    from oracle.ref_unet import RefUnet
    from oracle.ref_discriminator import RefDomainDiscriminator
    torch.manual_seed(1)
    seg, disc = RefUnet("resnet18", classes=5).eval(), RefDomainDiscriminator().eval()
    a, b = DomainAdaptationModel(seg, disc), RefDomainAdaptationModel(seg, disc)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        assert torch.equal(a(x), b(x))
        (s1, d1), (s2, d2) = a(x, domain_adaptation=True), b(x, domain_adaptation=True)
        assert torch.equal(s1, s2) and torch.equal(d1, d2)
        assert all(torch.equal(u, v) for u, v in zip(a.get_features(x), b.get_features(x)))
        assert torch.equal(DomainAdaptationModel(seg)(x, domain_adaptation=True),
                           RefDomainAdaptationModel(seg)(x, domain_adaptation=True))
    assert [id(p) for p in a.parameters()] == [id(p) for p in b.parameters()]
    assert b.train() is b and seg.training and disc.training
    assert b.eval() is b and not seg.training and not disc.training
    assert RefDomainAdaptationModel(torch.nn.Identity()).get_features(x) is None


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="reference tree not present")
def test_reference_domain_model_wraps_uda_b200_networks(monkeypatch):
    """The reference's OWN DomainAdaptationModel around the uda_b200 Unet + DomainDiscriminator (engine executed by the
    oracle ops on the CPU — the GPU twin of this test is tests/test_gpu_unet.py::test_domain_adaptation_model_wrap)."""
    sys.path.insert(0, REFERENCE)
    try:
        from src.models.domain_model import DomainAdaptationModel
    finally:
        sys.path.remove(REFERENCE)
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200 import unet, engine, discriminator
    from oracle import ref_ops
    from oracle.ref_unet import RefUnet
    from oracle.ref_discriminator import RefDomainDiscriminator
    for mod in (unet, engine, discriminator):
        monkeypatch.setattr(mod, "ops", ref_ops)
    monkeypatch.setattr(unet.Unet, "_prepare", lambda self, device: self._store.ensure_flat(device))
    monkeypatch.setattr(discriminator.DomainDiscriminator, "_prepare", lambda self, device: self._store.ensure_flat(device))
    torch.manual_seed(2)
    rseg, rdisc = RefUnet("resnet18", classes=4), RefDomainDiscriminator()
    seg = U.Unet("resnet18", classes=4, compute_dtype=torch.float32)
    disc = discriminator.DomainDiscriminator(compute_dtype=torch.float32)
    seg.load_state_dict(rseg.state_dict()); disc.load_state_dict(rdisc.state_dict())
    ours, ref = DomainAdaptationModel(seg, disc).eval(), DomainAdaptationModel(rseg, rdisc).eval()
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        (s1, d1), (s2, d2) = ours(x, domain_adaptation=True), ref(x, domain_adaptation=True)
        assert torch.allclose(s1, s2, atol=1e-4) and torch.allclose(d1, d2, atol=1e-5)
        f1, f2 = ours.get_features(x), ref.get_features(x)
        assert len(f1) == 6 and all(torch.allclose(u, v, atol=1e-4) for u, v in zip(f1, f2))
    assert len(ours.parameters()) == len(ref.parameters())
    ours.train()
    assert seg.training and disc.training


def test_trainer_metrics_known_answers():
    """oracle.ref_metrics.trainer_metrics_from_hist (restated torchmetrics macro Jaccard / binary Jaccard / accuracy of
    src/models/train.py:225-243) on hand-computed confusion matrices."""
    r = M.trainer_metrics_from_hist(np.array([[2, 1], [0, 3]]))
    assert abs(r["iou"] - (2 / 3 + 3 / 4) / 2) < 1e-12 and abs(r["accuracy"] - 5 / 6) < 1e-12
    assert np.allclose(r["iou_per_class"], [2 / 3, 3 / 4])
    # class 2 occurs nowhere: ignored by the macro average; class 1 only predicted: counts with IoU 0
    r = M.trainer_metrics_from_hist(np.array([[4, 2, 0], [0, 0, 0], [0, 0, 0]]))
    assert np.allclose(r["iou_per_class"], [4 / 6, 0.0, 0.0]) and abs(r["iou"] - (4 / 6 + 0.0) / 2) < 1e-12
    assert M.trainer_metrics_from_hist(np.zeros((3, 3)))["iou"] == 0.0
    # consistency with the in-tree metric on the same pixels
    rng = np.random.default_rng(0)
    p, t = rng.integers(0, 5, 1000), rng.integers(0, 5, 1000)
    h = M.fast_hist(p, t, 5)
    assert abs(M.trainer_metrics_from_hist(h)["mean_iou"] - M.batch_iou(p, t, 5)["mean_iou"]) < 1e-12
    assert abs(M.trainer_metrics_from_hist(h)["accuracy"] - float((p == t).mean())) < 1e-12
