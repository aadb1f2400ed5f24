"""GPU parity of the output-space adversarial path (north-star config 3: output-space discriminator behind a
gradient-reversal layer).  The reference DEFINES both building blocks — ``gradient_reverse_layer``
(src/models/uda.py:99-112) and ``DomainDiscriminator`` (src/models/discriminator.py:4-55) — and never wires them
together ("beyond reference"); the oracle is therefore the PyTorch composition of the oracle's pinned
restatements of exactly those two blocks:  D(gradient_reverse_layer(softmax(logits, 1), alpha))."""
import copy

import pytest
import torch

from conftest import rel_err, l2_err
from oracle import ref_losses as R
from oracle.ref_discriminator import RefDomainDiscriminator

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from uda_aerial_semantic_segmentation_research_b200 import ops
    return ops


@pytest.mark.parametrize("shape", [(2, 24, 32, 48), (1, 16, 16, 16), (3, 5, 8, 24), (2, 32, 8, 8)])
def test_softmax_pack_and_reversed_backward_kernels(shape):
    ops = _ops()
    B, C, H, W = shape
    cpad = 8 if C <= 8 else (16 if C <= 16 else 32)
    g = torch.Generator().manual_seed(C)
    z = (torch.randn(shape, generator=g) * 3).to(DEV)
    p = ops.softmax_nhwc(z, cpad)
    pr = torch.softmax(z, 1).permute(0, 2, 3, 1)
    assert p.shape == (B, H, W, cpad) and p.dtype == torch.bfloat16
    assert rel_err(p[..., :C].float(), pr) < 2 ** -8           # bf16 storage of the probabilities
    assert float(p[..., C:].abs().max()) == 0.0 if cpad > C else True
    # backward on the SAME bf16 probabilities: scale * p * (dp - sum p dp), exact in fp32
    dp = torch.randn(B, H, W, cpad, generator=g).bfloat16().to(DEV)
    pf, dpf = p.float()[..., :C], dp.float()[..., :C]
    ref = -0.7 * pf * (dpf - (pf * dpf).sum(-1, keepdim=True))
    got = ops.softmax_bwd_grl(p, dp, -0.7, C)
    assert got.shape == (B, C, H, W) and rel_err(got, ref.permute(0, 3, 1, 2)) < 1e-5
    acc = torch.randn(B, C, H, W, generator=g).to(DEV)
    base = acc.clone()
    out = ops.softmax_bwd_grl(p, dp, -0.7, C, out=acc)           # accumulate into an existing logit gradient
    assert out.data_ptr() == acc.data_ptr() and rel_err(acc, base + ref.permute(0, 3, 1, 2)) < 1e-5


def test_gradient_reverse_layer_kernel():
    """a10: identity forward, -alpha * grad backward in one kernel pass (fp32 and bf16, ragged sizes)."""
    from uda_aerial_semantic_segmentation_research_b200 import losses as L
    for n, dtype in ((5, torch.float32), (4099, torch.float32), (1 << 16, torch.bfloat16), (37, torch.bfloat16)):
        x = torch.randn(n, device=DEV).to(dtype).requires_grad_()
        w = torch.randn(n, device=DEV).to(dtype)
        y = L.gradient_reverse_layer(x, 0.3)
        assert torch.equal(y, x)
        (y * w).sum().backward()
        xr = x.detach().clone().cpu().float().requires_grad_()
        (R.gradient_reverse_layer(xr, 0.3) * w.cpu().float()).sum().backward()
        assert rel_err(x.grad.float().cpu(), xr.grad) < (1e-6 if dtype == torch.float32 else 2 ** -8)


def test_output_space_adversary_against_reference_composition():
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator, OutputSpaceAdversary
    from uda_aerial_semantic_segmentation_research_b200.losses import AdversarialLoss
    C, alpha = 24, 0.5
    torch.manual_seed(3)
    ref = RefDomainDiscriminator(C)
    D = DomainDiscriminator(C)
    D.load_state_dict(ref.state_dict())
    D = D.to(DEV).train()
    adv = OutputSpaceAdversary(D, alpha)
    # the oracle composition with bf16 storage at the same points (probabilities, conv weights / outputs, activations)
    r16 = copy.deepcopy(ref).train()
    with torch.no_grad():
        for mod in r16.modules():
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.copy_(mod.weight.bfloat16().float())
    rnd = lambda _m, _i, out: out.bfloat16().float()
    for mod in r16.features:
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.LeakyReLU)):
            mod.register_forward_hook(rnd)
    g = torch.Generator().manual_seed(4)
    z = (torch.randn(4, C, 128, 128, generator=g) * 2)
    zg = z.to(DEV).requires_grad_()
    y = adv(zg)
    zr = z.clone().requires_grad_()
    pr = torch.softmax(zr, 1)
    pr = pr + (pr.bfloat16().float() - pr).detach()               # bf16-stored probabilities, straight-through
    yr = r16(R.gradient_reverse_layer(pr, alpha))
    assert y.shape == (4, 1) and rel_err(y.detach().cpu(), yr.detach()) < 1e-2
    loss = AdversarialLoss().discriminator_loss(y[:2], y[2:])
    lr = R.adversarial_discriminator_loss(yr[:2], yr[2:]) if hasattr(R, "adversarial_discriminator_loss") else None
    loss.backward()
    if lr is None:
        bce = torch.nn.BCEWithLogitsLoss()
        lr = (bce(yr[:2], torch.ones_like(yr[:2])) + bce(yr[2:], torch.zeros_like(yr[2:]))) * 0.5
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-3 * abs(lr.item())
    e_z = l2_err(zg.grad.cpu(), zr.grad)
    worst = 0.0
    for (n, p), (_, p2) in zip(D.named_parameters(), r16.named_parameters()):
        if n in ("features.2.bias", "features.5.bias", "features.8.bias"):
            continue
        worst = max(worst, l2_err(p.grad.cpu(), p2.grad))
    print(f"output-space adversary: y {rel_err(y.detach().cpu(), yr.detach()):.2e}  dlogits L2 {e_z:.2e}  worst dparam L2 {worst:.2e}")
    assert e_z < 5e-2 and worst < 5e-2
    # the reversal itself: the logit gradient is -alpha times the gradient of the un-reversed composition
    z2 = z.clone().requires_grad_()
    y2 = r16(torch.softmax(z2, 1))
    bce = torch.nn.BCEWithLogitsLoss()
    ((bce(y2[:2], torch.ones_like(y2[:2])) + bce(y2[2:], torch.zeros_like(y2[2:]))) * 0.5).backward()
    cos = torch.nn.functional.cosine_similarity(zg.grad.cpu().flatten(), z2.grad.flatten(), dim=0)
    assert float(cos) < -0.95
    assert abs(float(zg.grad.norm().cpu() / z2.grad.norm()) - alpha) < 0.1 * alpha


def test_adversarial_grl_training_step_runs_and_is_capturable():
    """One optimizer step of the GRL form: CE(source) + BCE(D(GRL(softmax(logits_src))), 1) + BCE(D(GRL(softmax(
    logits_tgt))), 0), one backward, both networks stepped."""
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator, OutputSpaceAdversary
    from uda_aerial_semantic_segmentation_research_b200.losses import AdversarialLoss, CrossEntropyLoss
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    torch.manual_seed(0)
    model = U.Unet("resnet18", classes=24).to(DEV).train()
    disc = DomainDiscriminator(24).to(DEV).train()
    adv, bce, ce = OutputSpaceAdversary(disc, 0.1), AdversarialLoss(), CrossEntropyLoss()
    opt = FusedAdam([model, disc], lr=1e-4)
    xs, xt = torch.randn(2, 3, 64, 64, device=DEV), torch.randn(2, 3, 64, 64, device=DEV)
    ts = torch.randint(0, 24, (2, 64, 64), device=DEV)
    w0, d0 = model._store.flat.clone() if model._store.flat is not None else None, None
    losses = []
    for _ in range(3):
        opt.zero_grad()
        ls, lt = model(xs), model(xt)
        total = ce(ls, ts) + bce.discriminator_loss(adv(ls), adv(lt))
        total.backward()
        opt.step()
        losses.append(float(total))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert disc._store.grad is not None and float(disc._store.grad.abs().sum()) > 0
