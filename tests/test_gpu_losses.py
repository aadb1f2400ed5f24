"""GPU parity of the fused loss kernels (through the C ABI) against the golden vectors generated from the
reference's own modules, the oracle restatement at larger sizes, and size-independent properties at
BASELINE sizes.  Tolerance (north-star): losses and gradients within 1e-3 relative (fp32 kernels are
held to 1e-4; bf16 inputs are compared against the oracle evaluated on the same bf16-rounded logits)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import ref_losses as R

pytestmark = pytest.mark.gpu
CASES = sorted(glob.glob(os.path.join(GOLDEN, "losses_*.npz")))
TOL = 1e-4


def _dev():
    return torch.device("cuda:0")


def _L():
    from uda_aerial_semantic_segmentation_research_b200 import losses
    return losses


def _check(loss_t, ref_val, grads, ref_grads, tol=TOL, what=""):
    assert loss_t.dim() == 0 and loss_t.is_cuda
    rv = float(ref_val)
    assert abs(loss_t.item() - rv) <= tol * max(abs(rv), 1e-6), (what, loss_t.item(), rv)
    for i, (g, rg) in enumerate(zip(grads, ref_grads)):
        assert rel_err(g.float(), rg) < tol, (what, i, rel_err(g.float(), rg))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(c) for c in CASES])
def test_golden_vectors(path):
    L = _L()
    d = np.load(path)
    dev = _dev()
    z1 = torch.from_numpy(d["z1"]).to(dev).requires_grad_()
    z2 = torch.from_numpy(d["z2"]).to(dev).requires_grad_()
    t = torch.from_numpy(d["target"]).to(dev)
    w = torch.from_numpy(d["class_weights"]).to(dev)
    C = z1.shape[1]

    def run(key, fn, wrt):
        for x in wrt:
            x.grad = None
        loss = fn()
        loss.backward()
        _check(loss, d[key], [x.grad for x in wrt], [d[f"{key}_grad{i}"] for i in range(len(wrt))], what=key)

    run("ce", lambda: L.CrossEntropyLoss()(z1, t), [z1])
    run("dice", lambda: L.DiceLoss()(z1, t), [z1])
    run("ce_plus_dice", lambda: L.CombinedCEDiceLoss()(z1, t), [z1])
    run("weighted", lambda: L.WeightedSegmentationLoss(C, w)(z1, t, domain_weight=0.7), [z1])
    run("weighted_noweights_sum", lambda: L.WeightedSegmentationLoss(C, reduction="sum")(z1, t), [z1])
    run("consistency", lambda: L.ConsistencyLoss(0.5)(z1, z2), [z1, z2])
    run("consistency_T1", lambda: L.ConsistencyLoss(1.0)(z1, z2), [z1, z2])
    s = torch.from_numpy(d["d_src"]).to(dev).requires_grad_()
    dd = torch.from_numpy(d["d_tgt"]).to(dev).requires_grad_()
    adv = L.AdversarialLoss(0.001)
    run("disc_loss", lambda: adv.discriminator_loss(s, dd), [s, dd])
    run("gen_loss", lambda: adv.generator_loss(dd), [dd])
    # one-hot float targets, as the reference's loss_functions_suite passes them (src/test_system.py:117-119)
    onehot = torch.nn.functional.one_hot(t, C).permute(0, 3, 1, 2).float().contiguous()
    run("dice", lambda: L.DiceLoss()(z1, onehot), [z1])
    ft = L.FineTuningLoss()
    for ep in (0, 20, 60):
        for x in (z1, z2, dd):
            x.grad = None
        r = ft(z1, z2, dd, ep, supervised_pred=z1, supervised_target=t)
        assert set(r) == {"total", "consistency", "domain_confusion", "supervised", "rampup_weight"}
        assert abs(r["total"].item() - float(d[f"ft_total_ep{ep}"])) <= TOL * max(1, abs(float(d[f"ft_total_ep{ep}"])))
        assert r["rampup_weight"].item() == float(d[f"ft_ramp_ep{ep}"])
        if ep == 20:
            r["total"].backward()
            assert rel_err(z1.grad, d["ft_ep20_grad_z1"]) < TOL
            assert rel_err(z2.grad, d["ft_ep20_grad_z2"]) < TOL
            assert rel_err(dd.grad, d["ft_ep20_grad_d"]) < TOL
            assert abs(r["consistency"].item() - float(d["ft_ep20_consistency"])) <= TOL * float(d["ft_ep20_consistency"])
            assert abs(r["supervised"].item() - float(d["ft_ep20_supervised"])) <= TOL


def test_closed_form_kats():
    L = _L()
    dev = _dev()
    assert abs(L.DiceLoss()(torch.zeros(1, 24, 8, 8, device=dev), torch.zeros(1, 8, 8, dtype=torch.long, device=dev)).item() - 0.734736) < 1e-5
    z = torch.zeros(4, 1, device=dev)
    adv = L.AdversarialLoss(0.001)
    assert abs(adv.discriminator_loss(z, z).item() - 0.693147) < 1e-5
    assert abs(adv.generator_loss(z).item() - 6.93147e-4) < 1e-8
    x = torch.randn(2, 24, 16, 16, device=dev)
    assert abs(L.ConsistencyLoss()(x, x.clone()).item()) < 1e-4
    g = torch.ones(5, device=dev, requires_grad=True)
    L.gradient_reverse_layer(g, 0.3).sum().backward()
    assert torch.allclose(g.grad, torch.full((5,), -0.3, device=dev))


# the last shapes are large enough (>= 1 MiB of logits) for the streaming kernels, with CTAs that cross image
# boundaries; (2, 19, ...) exercises the padded-class instance; 144 x 112 = 63 x 256 pixels per image: the tensor-map
# path with a PARTIAL last 512-pixel tile (out-of-range half zero-filled on load, clipped on store); 136 x 120 is not a
# multiple of 256 pixels: the row-copy path
@pytest.mark.parametrize("shape", [(2, 24, 64, 64), (1, 23, 37, 53), (3, 5, 8, 8), (2, 33, 16, 16),
                                   (3, 24, 128, 128), (2, 19, 128, 96), (2, 24, 144, 112), (2, 24, 136, 120)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_against_oracle(shape, dtype):
    """Odd class counts / ragged sizes (scalar path), C > 32 (64-wide instance), bf16 logits."""
    L = _L()
    dev = _dev()
    B, C, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + C)
    z1c = (torch.randn(shape, generator=g) * 3).to(dtype)
    z2c = (torch.randn(shape, generator=g) * 3).to(dtype)
    t = torch.randint(0, C, (B, H, W), generator=g)
    t_ign = t.clone()
    t_ign[:, ::3, ::2] = -100
    w = torch.rand(C, generator=g) + 0.5
    tol = 1e-4 if dtype == torch.float32 else 1e-2  # bf16 gradient storage: 2^-8 relative per element
    ltol = 1e-4 if dtype == torch.float32 else 1e-3

    def both(fn_gpu, fn_ref, n_in=1):
        zs = [z1c, z2c][:n_in]
        gl = [z.detach().clone().to(dev).requires_grad_() for z in zs]
        rl = [z.detach().clone().float().requires_grad_() for z in zs]
        lg, lr = fn_gpu(*gl), fn_ref(*rl)
        lg.backward(); lr.backward()
        assert abs(lg.item() - lr.item()) <= ltol * max(abs(lr.item()), 1e-6), (lg.item(), lr.item())
        for a, b in zip(gl, rl):
            assert a.grad.dtype == dtype
            assert rel_err(a.grad.float(), b.grad) < tol

    td, wd = t.to(dev), w.to(dev)
    both(lambda z: L.CrossEntropyLoss()(z, td), lambda z: R.cross_entropy(z, t))
    both(lambda z: L.CrossEntropyLoss()(z, t_ign.to(dev)), lambda z: R.cross_entropy(z, t_ign))
    both(lambda z: L.CrossEntropyLoss(weight=wd)(z, td),
         lambda z: torch.nn.functional.cross_entropy(z, t, weight=w))
    both(lambda z: L.DiceLoss(smooth=0.5)(z, td), lambda z: R.dice_loss(z, t, 0.5))
    both(lambda z: L.CombinedCEDiceLoss(0.7, 1.3)(z, td), lambda z: 0.7 * R.cross_entropy(z, t) + 1.3 * R.dice_loss(z, t))
    both(lambda z: L.WeightedSegmentationLoss(C, wd, alpha=0.5, gamma=1.5)(z, td),
         lambda z: R.weighted_segmentation_loss(z, t, w, alpha=0.5, gamma=1.5))
    both(lambda a, b: L.ConsistencyLoss(0.7)(a, b), lambda a, b: R.consistency_loss(a, b, 0.7), 2)
    both(lambda z: L.EntropyMinimizationLoss()(z), lambda z: R.entropy_loss(z))
    # a non-unit upstream gradient exercises the device-scalar rescale path
    both(lambda z: 3.0 * L.CrossEntropyLoss()(z, td), lambda z: 3.0 * R.cross_entropy(z, t))


def test_full_size_properties():
    """BASELINE cfg2 size (B=16, C=24, 512x512): properties that need no CPU oracle run."""
    L = _L()
    dev = _dev()
    B, C, H, W = 16, 24, 512, 512
    g = torch.Generator(device=dev).manual_seed(7)
    z = (torch.randn(B, C, H, W, device=dev, generator=g) * 3).requires_grad_()
    t = torch.randint(0, C, (B, H, W), device=dev, generator=g)
    ce = L.CrossEntropyLoss()(z, t)
    ce.backward()
    gce = z.grad.clone()
    # softmax-CE gradient sums to zero over classes at every pixel and to (p - onehot)/N overall
    assert gce.sum(1).abs().max().item() < 1e-9
    picked = gce.gather(1, t.unsqueeze(1))
    assert (picked <= 0).all() and (gce.sum(1, keepdim=True) - picked).min() >= -1e-12
    # linearity of the fused CE+Dice kernel in its weights
    z.grad = None
    L.DiceLoss()(z, t).backward()
    gd = z.grad.clone()
    z.grad = None
    comb = L.CombinedCEDiceLoss(0.5, 2.0)(z, t)
    comb.backward()
    assert rel_err(z.grad, 0.5 * gce + 2.0 * gd) < 1e-4
    assert abs(comb.item() - (0.5 * ce.item() + 2.0 * L.DiceLoss()(z.detach(), t).item())) < 1e-4
    # CE against torch on device (same fp32 logits)
    zr = z.detach().clone().requires_grad_()
    lr = torch.nn.functional.cross_entropy(zr, t)
    lr.backward()
    assert abs(ce.item() - lr.item()) < 1e-4 * lr.item() and rel_err(gce, zr.grad) < 1e-4
    # consistency: symmetric in its arguments, zero for identical views, entropy in [0, log C]
    z2 = torch.randn(B, C, H, W, device=dev, generator=g)
    c12 = L.ConsistencyLoss()(z.detach(), z2).item()
    c21 = L.ConsistencyLoss()(z2, z.detach()).item()
    assert abs(c12 - c21) <= 1e-4 * abs(c12) and c12 > 0
    e = L.EntropyMinimizationLoss()(z.detach()).item()
    assert 0 <= e <= np.log(C) + 1e-5
