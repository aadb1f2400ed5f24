"""GPU parity of the whole hot path behind the reference's entry points (model creation, losses,
optimizer step, prediction) against the fp32 oracle on identical synthetic inputs and weights.

Tolerances (north-star): logits 1e-4 relative in fp32 mode / 2e-2 in bf16; losses 1e-3.  Gradients: the
logit-gradient and every per-kernel gradient are held to 1e-3 in tests/test_gpu_layers.py and
tests/test_gpu_losses.py; for whole-network WEIGHT gradients at random initialisation the max-norm is
dominated by ReLU-mask flips caused by 1e-6-level forward differences (the fp32 oracle disagrees with
its own fp64 run by up to 5e-2 there — see DESIGN.md "Gradient parity"), so they are compared in the
L2 norm with the tolerance stated next to each assertion."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err, l2_err
from oracle.ref_unet import RefUnet
from oracle.ref_discriminator import RefDomainDiscriminator
from oracle import ref_losses as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pkg():
    import uda_aerial_semantic_segmentation_research_b200 as U
    return U


def _pair(enc, classes, dtype, seed=0):
    U = _pkg()
    torch.manual_seed(seed)
    ref = RefUnet(enc, classes=classes)
    m = U.Unet(encoder_name=enc, encoder_weights=None, in_channels=3, classes=classes, compute_dtype=dtype)
    m.load_state_dict(ref.state_dict())
    return m.to(DEV), ref


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_fp32_mode_logits_and_gradients(enc):
    m, ref = _pair(enc, 24, torch.float32)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 64, 64, generator=g)
    t = torch.randint(0, 24, (2, 64, 64), generator=g)
    y = m(x.to(DEV))
    yr = ref(x)
    assert y.shape == (2, 24, 64, 64) and y.dtype == torch.float32
    assert rel_err(y.detach().cpu(), yr.detach()) < 1e-4            # north-star: 1e-4 in fp32 mode
    from uda_aerial_semantic_segmentation_research_b200.losses import CombinedCEDiceLoss
    loss = CombinedCEDiceLoss()(y, t.to(DEV))
    lr = R.cross_entropy(yr, t) + R.dice_loss(yr, t)
    assert abs(loss.item() - lr.item()) < 1e-3 * lr.item()          # north-star: 1e-3
    loss.backward(); lr.backward()
    worst = 0.0
    for (n, p), (_, p2) in zip(m.named_parameters(), ref.named_parameters()):
        worst = max(worst, l2_err(p.grad.cpu(), p2.grad))
    assert worst < 8e-2, worst     # mask-flip dominated (see module docstring); logic is exact in fp64 on CPU
    # head and last decoder BN are upstream of any flip amplification: tight
    pg = dict(m.named_parameters()); rg = dict(ref.named_parameters())
    for n in ("segmentation_head.0.weight", "segmentation_head.0.bias", "decoder.blocks.4.conv2.1.weight"):
        assert rel_err(pg[n].grad.cpu(), rg[n].grad) < 1e-3, n
    for (n, b), (_, b2) in zip(m.named_buffers(), ref.named_buffers()):
        assert rel_err(b.cpu(), b2) < 1e-3, n


def test_bf16_mode_logits_cfg1():
    """BASELINE config 1: U-Net r34, batch 2 @256x256, 24 classes, fwd/bwd + CE/Dice, bf16 tensor-core path.

    bf16 parity is checked against the reference PyTorch path evaluated "in bf16" (the oracle with bf16
    storage rounding at the same points, oracle.ref_unet.emulate_bf16):
      * 2e-2 (north-star) on the shallow parts: encoder stages 1-2 and the whole decoder+head run on
        identical features;
      * for the full 47-conv network at random initialisation in train mode the comparison is chaotic:
        two bf16 evaluations that differ only by 1e-6 relative accumulation noise end up ~1e-1 apart
        (measured on the oracle, DESIGN.md "bf16 parity"), and bf16 vs fp32 is ~1.5e-1.  There the
        kernel must stay within 1.5x that natural spread (+ the 2e-2 gate) of both references; every STAGE
        is held to the 2e-2 gate teacher-forced in tests/test_gpu_stages.py."""
    from oracle.ref_unet import emulate_bf16
    m, ref = _pair("resnet34", 24, torch.bfloat16)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 256, 256, generator=g)
    t = torch.randint(0, 24, (2, 16, 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    ref16 = emulate_bf16(ref)
    with torch.no_grad():
        yr = ref(x)
        f16 = ref16.encoder(x)
        yr16 = ref16.segmentation_head(ref16.decoder(*f16))
    nat = rel_err(yr16, yr)
    # shallow chains: tight gate
    feats = m.encoder(x.to(DEV))
    e1, e2 = rel_err(feats[1].detach().cpu(), f16[1]), rel_err(feats[2].detach().cpu(), f16[2])
    fin = [f.to(DEV) for f in f16]
    yd = m.segmentation_head(m.decoder(*fin))
    ed = rel_err(yd.detach().cpu(), yr16)
    # full network
    y = m(x.to(DEV))
    err16, err32 = rel_err(y.detach().cpu(), yr16), rel_err(y.detach().cpu(), yr)
    print(f"bf16 parity: enc1 {e1:.2e} enc2 {e2:.2e} decoder+head {ed:.2e} | full vs bf16 ref {err16:.2e}, "
          f"vs fp32 {err32:.2e}, bf16 ref vs fp32 {nat:.2e}")
    assert e1 < 2e-2 and e2 < 2e-2 and ed < 2e-2, (e1, e2, ed)      # north-star: 2e-2 in bf16
    assert err16 < 1.5 * nat + 2e-2 and err32 < 1.5 * nat + 2e-2, (err16, err32, nat)
    from uda_aerial_semantic_segmentation_research_b200.losses import CombinedCEDiceLoss
    loss = CombinedCEDiceLoss()(y, t.to(DEV))
    loss.backward()
    assert torch.isfinite(loss).item()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # bf16 tensor-core path and FP32-pipe direct path on the same weights: same natural-spread bound
    from uda_aerial_semantic_segmentation_research_b200 import ops
    ops.USE_TC = False
    try:
        y_direct = m(x.to(DEV))
    finally:
        ops.USE_TC = True
    assert rel_err(y.detach(), y_direct.detach()) < 1.5 * nat + 2e-2


def test_bf16_eval_mode_logits():
    """Eval mode (running statistics, as in prediction)."""
    from oracle.ref_unet import emulate_bf16
    m, ref = _pair("resnet34", 24, torch.bfloat16, seed=2)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 128, 128, generator=g)
    ref.train()
    with torch.no_grad():   # a few batches' worth of running statistics so eval activations are well scaled
        for _ in range(3):
            ref(x + 0.1 * torch.randn(x.shape, generator=g))
    m.load_state_dict(ref.state_dict())
    m.eval(); ref.eval()
    with torch.no_grad():
        y = m(x.to(DEV)).cpu()
        yr, yr16 = ref(x), emulate_bf16(ref).eval()(x)
    nat = rel_err(yr16, yr)
    print(f"eval bf16 logits: vs bf16 reference {rel_err(y, yr16):.3e}, vs fp32 oracle {rel_err(y, yr):.3e}, "
          f"bf16 reference vs fp32 {nat:.3e}")
    # inference folds BatchNorm into the weights (no bf16 rounding of the pre-normalisation tensor): a different, more
    # accurate rounding sequence than the bf16 reference's — compared with both references at the natural-spread level
    assert rel_err(y, yr16) < 1.5 * nat + 2e-2
    assert rel_err(y, yr) < 1.5 * nat + 2e-2


def test_training_reduces_loss_and_matches_oracle_trend():
    """Reference step semantics (src/models/train.py:336-346): zero_grad, forward, CE, backward, Adam step."""
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    for dtype in (torch.float32, torch.bfloat16):
        m, ref = _pair("resnet34", 6, dtype, seed=3)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(4, 3, 64, 64, generator=g)
        t = torch.randint(0, 6, (4, 4, 4), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
        opt, opt_r = FusedAdam(m, lr=1e-3), torch.optim.Adam(ref.parameters(), lr=1e-3)
        crit = CrossEntropyLoss()
        ls, lrs = [], []
        for _ in range(8):
            opt.zero_grad(); opt_r.zero_grad()
            l = crit(m(x.to(DEV)), t.to(DEV)); l.backward(); opt.step(); ls.append(l.item())
            lr = F.cross_entropy(ref(x), t); lr.backward(); opt_r.step(); lrs.append(lr.item())
        assert ls[-1] < 0.7 * ls[0], ls
        assert abs(ls[0] - lrs[0]) < 2e-2 * lrs[0]
        assert abs(ls[-1] - lrs[-1]) < 0.25 * lrs[0], (ls, lrs)     # same trajectory, chaotic in the details
    # the stock torch optimizer also works on the same parameters (drop-in: train.py:461)
    m, _ = _pair("resnet34", 6, torch.bfloat16, seed=4)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    l0 = None
    for _ in range(4):
        opt.zero_grad()
        l = crit(m(x.to(DEV)), t.to(DEV)); l.backward(); opt.step()
        l0 = l0 or l.item()
    assert l.item() < l0


def test_discriminator_golden_and_adversarial_step():
    """DomainDiscriminator against the golden vectors of the reference's own module (fp32 mode), then one
    reference adversarial iteration (src/models/adversarial_trainer.py:76-114) end to end in bf16."""
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    from uda_aerial_semantic_segmentation_research_b200.losses import AdversarialLoss, CrossEntropyLoss
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    d = np.load(os.path.join(GOLDEN, "discriminator_small.npz"))
    torch.manual_seed(11)
    ref = RefDomainDiscriminator(3)        # same construction order as the reference module -> same init
    ok = all(abs(float(v.double().abs().sum()) - float(d["sd_abs_sum/" + k])) <= 1e-6 * max(1.0, float(d["sd_abs_sum/" + k]))
             for k, v in ref.state_dict().items() if "running" not in k and "num_batches" not in k)
    if not ok:
        pytest.skip("torch RNG stream differs from the fixture's; weights cannot be regenerated")
    D = DomainDiscriminator(3, compute_dtype=torch.float32)
    D.load_state_dict(ref.state_dict())
    D = D.to(DEV).train()
    x = torch.from_numpy(d["x"]).to(DEV).requires_grad_()
    y = D(x)
    assert y.shape == (2, 1) and rel_err(y.detach().cpu(), d["y_train"]) < 1e-4
    loss = AdversarialLoss().discriminator_loss(y[:1], y[1:])
    assert abs(loss.item() - float(d["loss"])) < 1e-4 * float(d["loss"])
    loss.backward()
    assert rel_err(x.grad.cpu(), d["grad_x"]) < 1e-3
    for n, p in D.named_parameters():
        if "grad/" + n in d.files:
            gref = d["grad/" + n]
            if np.abs(gref).max() < 1e-7:   # conv bias in front of a BatchNorm: analytically zero gradient
                assert p.grad.abs().max().item() < 1e-6, n
            else:
                assert rel_err(p.grad.cpu(), gref) < 1e-3, n
    for k in d.files:
        if k.startswith("sd_after/") and "running" in k:
            assert rel_err(D.state_dict()[k[len("sd_after/"):]].cpu(), d[k]) < 1e-4, k
    D.eval()
    with torch.no_grad():
        assert rel_err(D(x.detach()).cpu(), d["y_eval"]) < 1e-4
    # --- one adversarial iteration, bf16, reference semantics ---
    U = _pkg()
    torch.manual_seed(0)
    model = U.Unet("resnet34", classes=24).to(DEV)
    disc = DomainDiscriminator(3).to(DEV)
    opt, dopt = FusedAdam(model, lr=1e-4), FusedAdam(disc, lr=1e-4)
    adv, crit = AdversarialLoss(0.001), CrossEntropyLoss()
    src, tgt = torch.randn(2, 3, 64, 64, device=DEV), torch.randn(2, 3, 64, 64, device=DEV)
    masks = torch.randint(0, 24, (2, 64, 64), device=DEV)
    model.train(); disc.train()
    dopt.zero_grad()
    sp, tp = disc(src), disc(tgt)
    assert sp.shape == (2, 1) and bool(((sp >= 0) & (sp <= 1)).all())
    d_loss = adv.discriminator_loss(sp, tp)
    d_loss.backward(); dopt.step()
    opt.zero_grad()
    seg_loss = crit(model(src), masks)
    adv_loss = adv.generator_loss(disc(tgt))
    total = seg_loss + adv_loss
    total.backward(); opt.step()
    assert torch.isfinite(total).item() and d_loss.item() > 0


def test_predict_batch_and_sliding_window():
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.predict import predict_batch, sliding_window_evaluate, tile_windows
    m, ref = _pair("resnet34", 24, torch.float32, seed=9)
    ref.eval()
    x = torch.randn(2, 3, 64, 64)
    pm = predict_batch(m, x, DEV)
    assert pm.dtype == np.int64 and pm.shape == (2, 64, 64)
    with torch.no_grad():
        logits = m(x.to(DEV))                     # eval mode (predict_batch switched it)
        assert rel_err(logits.cpu(), ref(x)) < 1e-4
    assert np.array_equal(pm, logits.argmax(1).cpu().numpy())   # bit-exact given identical logits
    tile = torch.randn(3, 128, 192, device=DEV)
    target = torch.randint(0, 24, (128, 192), device=DEV)
    out = sliding_window_evaluate(m, tile, target, 24, window=64, batch=4, return_mask=True)
    wins = tile_windows(tile, 64)
    with torch.no_grad():
        full = torch.cat([m(wins[i:i + 4]) for i in range(0, wins.shape[0], 4)]).argmax(1)
    full = full.reshape(2, 3, 64, 64).permute(0, 2, 1, 3).reshape(128, 192)
    assert torch.equal(out["mask"], full)
    from oracle import ref_metrics as M
    assert np.array_equal(out["hist"].cpu().numpy(), M.fast_hist(full.cpu().numpy(), target.cpu().numpy(), 24))
    assert out["hist"].sum().item() == 128 * 192


def test_graphed_step_matches_eager_steps():
    """graph.GraphedStep (CUDA-graph replay of zero_grad+fwd+loss+bwd, then fused Adam) walks the same loss
    trajectory as the eager step sequence from the same initial state (fp32 mode, 1e-3 on each loss)."""
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.graph import GraphedStep
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(2, 3, 64, 64, generator=g) for _ in range(4)]
    ts = [torch.randint(0, 6, (2, 64, 64), generator=g) for _ in range(4)]
    warm = 2

    def make():
        torch.manual_seed(3)
        m = U.Unet("resnet18", encoder_weights=None, classes=6, compute_dtype=torch.float32).to(DEV).train()
        return m, FusedAdam(m, lr=1e-3), CrossEntropyLoss()

    m1, o1, c1 = make()
    eager = []
    for i in [0] * warm + [0, 1, 2, 3]:          # GraphedStep's warm-up steps run on the example batch
        o1.zero_grad()
        loss = c1(m1(xs[i].to(DEV)), ts[i].to(DEV))
        loss.backward()
        o1.step()
        eager.append(float(loss))
    m2, o2, c2 = make()
    step = GraphedStep(m2, c2, o2, xs[0], ts[0], warmup=warm)
    got = [float(step(xs[0].pin_memory(), ts[0].pin_memory()))]
    for i in (1, 2, 3):
        step.stage(xs[i].pin_memory(), ts[i].pin_memory())
        got.append(float(step()))
    assert step.launches_per_step > 50
    for i, (a, b) in enumerate(zip(eager[warm:], got)):   # same state at the first step; noise compounds afterwards
        assert abs(a - b) <= (1e-3 if i == 0 else 5e-3) * abs(a), (eager, got)
    # parameters after the trajectory agree too — loosely: Adam's m/sqrt(v) turns the atomics-order noise of a
    # near-zero gradient into a full +-lr step per iteration for those weights, so two runs of the SAME eager code
    # already differ by 2e-3..1e-2 in relative L2; a skipped or doubled update would show up as >= 1e-1
    assert l2_err(m2._store.flat, m1._store.flat) < 3e-2
    assert float((m2._store.flat - m1._store.flat).abs().max()) <= 2 * (warm + 4) * 3.2e-3   # |Adam step| <= 3.2 lr
    assert int(m2.encoder.bn1.num_batches_tracked) == int(m1.encoder.bn1.num_batches_tracked)


def test_graphed_phases_match_eager_adversarial_steps():
    """graph.GraphedPhases (D-step graph -> optimizer -> G-step graph -> optimizer: the data-parallel form of the
    adversarial step, here without the all-reduce) walks the same loss trajectory as the eager D/G steps of
    src/models/adversarial_trainer.py:84-114 (fp32 mode)."""
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.graph import GraphedPhases
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss, AdversarialLoss
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    g = torch.Generator().manual_seed(9)
    xs = torch.randn(2, 3, 64, 64, generator=g).to(DEV)
    xt = torch.randn(2, 3, 64, 64, generator=g).to(DEV)
    ts = torch.randint(0, 6, (2, 64, 64), generator=g).to(DEV)
    warm, steps = 2, 3

    def make():
        torch.manual_seed(4)
        m = U.Unet("resnet18", encoder_weights=None, classes=6, compute_dtype=torch.float32).to(DEV).train()
        d = DomainDiscriminator(3, compute_dtype=torch.float32).to(DEV).train()
        return m, d, FusedAdam(m, lr=1e-3), FusedAdam(d, lr=1e-4), CrossEntropyLoss(), AdversarialLoss(0.001)

    def phases(m, d, o, do, c, adv):
        def d_compute(a, t, b):
            do.zero_grad()
            loss = adv.discriminator_loss(d(a), d(b))
            loss.backward()
            return loss.detach()

        def g_compute(a, t, b):
            o.zero_grad()
            total = c(m(a), t) + adv.generator_loss(d(b))
            total.backward()
            return total.detach()
        return [(d_compute, do.step), (g_compute, o.step)]

    ph = phases(*make())
    eager = []
    for _ in range(warm + 1 + steps):            # GraphedPhases: warm-up steps + one trajectory step while capturing
        for compute, finish in ph:
            out = compute(xs, ts, xt)
            finish()
        eager.append(float(out))
    step = GraphedPhases(phases(*make()), [xs, ts, xt], [], warmup=warm)
    got = [float(step(xs, ts, xt)) for _ in range(steps)]
    assert step.launches_per_step > 60
    for a, b in zip(eager[warm + 1:], got):
        assert abs(a - b) <= 5e-3 * abs(a), (eager, got)


def test_shadow_weights_follow_parameter_writes():
    """in bf16 mode the tensor-core convolutions read a bf16 shadow copy of the weights; it must be
    refreshed after (a) ``load_state_dict`` on a model that has already run (reference phase_manager.py:140,
    test_system.py:260) and (b) a stock ``torch.optim.Adam`` step (reference train.py:461)."""
    U = _pkg()
    torch.manual_seed(21)
    m = U.Unet("resnet18", classes=5).to(DEV).eval()
    x = torch.randn(2, 3, 64, 64, device=DEV)
    with torch.no_grad():
        y0 = m(x).clone()
    # (a) load different weights into the used model: must equal a FRESH model built from the same state
    torch.manual_seed(22)
    other = U.Unet("resnet18", classes=5)
    sd = {k: v.clone() for k, v in other.state_dict().items()}
    m.load_state_dict(sd)
    fresh = U.Unet("resnet18", classes=5)
    fresh.load_state_dict(sd)
    fresh = fresh.to(DEV).eval()
    with torch.no_grad():
        y1, yf = m(x), fresh(x)
    assert torch.equal(y1, yf) and not torch.allclose(y1, y0)
    # (b) torch.optim.Adam writes through the nn.Parameters: conv outputs must move, and equal a fresh model's
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    t = torch.randint(0, 5, (2, 64, 64), device=DEV)
    F.cross_entropy(m(x), t).backward()
    opt.step()
    m.eval()
    fresh2 = U.Unet("resnet18", classes=5)
    fresh2.load_state_dict({k: v.detach().cpu().clone() for k, v in m.state_dict().items()})
    fresh2 = fresh2.to(DEV).eval()
    with torch.no_grad():
        y2, yf2 = m(x), fresh2(x)
    assert torch.equal(y2, yf2)
    assert rel_err(y2, y1) > 1e-3          # the convolution weights really changed (not only BatchNorm affine / biases)


def test_fused_adam_checkpoint_and_graph_replay_after_zero_grad():
    """FusedAdam exposes param_groups / state_dict / load_state_dict (reference train.py:361,496);
    a GraphedStep replay after a user-side zero_grad() steps normally."""
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
    from uda_aerial_semantic_segmentation_research_b200.graph import GraphedStep
    torch.manual_seed(5)
    m = U.Unet("resnet18", classes=4, compute_dtype=torch.float32).to(DEV).train()
    opt = FusedAdam(m, lr=1e-3)
    assert opt.param_groups[0]["lr"] == 1e-3
    crit = CrossEntropyLoss()
    x, t = torch.randn(2, 3, 64, 64, device=DEV), torch.randint(0, 4, (2, 64, 64), device=DEV)
    for _ in range(2):
        opt.zero_grad(); crit(m(x), t).backward(); opt.step()
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 2
    w = m._store.flat.clone()
    rs = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
    opt.zero_grad(); crit(m(x), t).backward(); opt.step()
    w_next = m._store.flat.clone()
    with torch.no_grad():
        m._store.flat.copy_(w)
    m.load_state_dict(rs, strict=False)
    opt2 = FusedAdam(m, lr=5.0)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 1e-3
    opt2.zero_grad(); crit(m(x), t).backward(); opt2.step()
    assert l2_err(m._store.flat, w_next) < 1e-5      # same update up to the atomics order of the gradient sums
    step = GraphedStep(m, crit, opt2, x, t, warmup=1)
    l1 = float(step(x, t))
    opt2.zero_grad()                                  # user-side zero_grad between replays
    l2 = float(step(x, t))
    assert l2 < l1 * 1.5 and np.isfinite(l2)


def test_eval_mode_folded_batchnorm_matches_unfolded_path():
    """Inference (reference predict.py:113-130): BatchNorm folded into the convolution weights / bias, residual add and
    ReLU in the conv epilogue — against the unfolded path (normalise pass per layer), the bf16 reference and fp32."""
    from oracle.ref_unet import emulate_bf16
    from uda_aerial_semantic_segmentation_research_b200 import ops
    m, ref = _pair("resnet34", 24, torch.bfloat16, seed=6)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 3, 128, 128, generator=g)
    ref.train()
    with torch.no_grad():
        for _ in range(3):
            ref(x + 0.1 * torch.randn(x.shape, generator=g))
    m.load_state_dict(ref.state_dict())
    m.eval(); ref.eval()
    with torch.no_grad():
        l0 = ops.LAUNCHES
        y_fold = m(x.to(DEV)).cpu()
        n_fold = ops.LAUNCHES - l0
        ops.FOLD_BN_EVAL = False
        try:
            l0 = ops.LAUNCHES
            y_plain = m(x.to(DEV)).cpu()
            n_plain = ops.LAUNCHES - l0
        finally:
            ops.FOLD_BN_EVAL = True
        yr, yr16 = ref(x), emulate_bf16(ref).eval()(x)
        l0 = ops.LAUNCHES
        m(x.to(DEV))
        n_cached = ops.LAUNCHES - l0
    nat = rel_err(yr16, yr)
    e_fold, e_plain = rel_err(y_fold, yr), rel_err(y_plain, yr)
    print(f"eval logits vs fp32 oracle: folded {e_fold:.3e}  unfolded {e_plain:.3e}  bf16 reference {nat:.3e}; "
          f"launches folded {n_fold} (cached weights {n_cached}) vs unfolded {n_plain}")
    assert e_fold < 1.5 * nat + 2e-2 and rel_err(y_fold, yr16) < max(2 * nat, 2e-2)
    assert rel_err(y_fold, y_plain) < max(2 * nat, 2e-2)
    assert n_cached <= n_plain - 46         # the 46 BatchNorm normalise passes are gone (147 -> 69 launches at r34)
    # folded weights follow the running statistics: a training forward must invalidate the cache — checked against the
    # oracle evaluated with the network's NEW running statistics
    m.train()
    with torch.no_grad():
        m(x.to(DEV))
    m.eval()
    ref.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()})
    ref.eval()
    with torch.no_grad():
        y_after = m(x.to(DEV)).cpu()
        yr2, yr16_2 = ref(x), emulate_bf16(ref).eval()(x)
    nat2 = rel_err(yr16_2, yr2)
    assert not torch.equal(y_after, y_fold)
    assert rel_err(y_after, yr2) < 1.5 * nat2 + 2e-2, (rel_err(y_after, yr2), nat2)


def test_bf16_discriminator_matches_bf16_reference():
    """the bf16 discriminator path against the oracle discriminator with bf16 storage rounding at the
    same points (input, conv weights, every conv output, every activation) — forward 1e-2, gradients in L2."""
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    import copy
    torch.manual_seed(12)
    ref = RefDomainDiscriminator(3)
    D = DomainDiscriminator(3)
    D.load_state_dict(ref.state_dict())
    D = D.to(DEV).train()
    r16 = copy.deepcopy(ref).train()
    with torch.no_grad():
        for mod in r16.modules():
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.copy_(mod.weight.bfloat16().float())
    rnd = lambda _m, _i, out: out.bfloat16().float()
    for mod in r16.features:
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.LeakyReLU)):
            mod.register_forward_hook(rnd)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(4, 3, 128, 128, generator=g).bfloat16().float()
    y = D(x.to(DEV))
    yr = r16(x)
    assert rel_err(y.detach().cpu(), yr.detach()) < 1e-2
    w = torch.tensor([[1.0], [-1.0], [0.5], [2.0]])
    (y * w.to(DEV)).sum().backward()
    (yr * w).sum().backward()
    worst, who = 0.0, None
    for (n, p), (_, p2) in zip(D.named_parameters(), r16.named_parameters()):
        if n in ("features.2.bias", "features.5.bias", "features.8.bias"):
            continue      # conv bias in front of a BatchNorm: analytically zero gradient (round-off noise on both sides)
        e = l2_err(p.grad.cpu(), p2.grad)
        if e > worst:
            worst, who = e, n
    print(f"bf16 discriminator: y {rel_err(y.detach().cpu(), yr.detach()):.2e}, worst parameter-gradient L2 error {worst:.2e} ({who})")
    assert worst < 5e-2      # three train-mode BatchNorm + LeakyReLU layers: mask flips at bf16 ties (see test_gpu_stages)


def test_domain_adaptation_model_wrap():
    """the reference's DomainAdaptationModel (src/models/domain_model.py:4-83; restated and pinned in
    oracle/ref_domain_model.py because the GPU box has no reference tree) around U.Unet + DomainDiscriminator."""
    U = _pkg()
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    from oracle.ref_domain_model import RefDomainAdaptationModel
    torch.manual_seed(14)
    rseg, rdisc = RefUnet("resnet18", classes=4), RefDomainDiscriminator()
    seg = U.Unet("resnet18", classes=4, compute_dtype=torch.float32)
    disc = DomainDiscriminator(compute_dtype=torch.float32)
    seg.load_state_dict(rseg.state_dict()); disc.load_state_dict(rdisc.state_dict())
    ours = RefDomainAdaptationModel(seg, disc).to(DEV).eval()
    ref = RefDomainAdaptationModel(rseg, rdisc).eval()
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        (s1, d1), (s2, d2) = ours(x.to(DEV), domain_adaptation=True), ref(x, domain_adaptation=True)
        assert rel_err(s1.cpu(), s2) < 1e-4 and rel_err(d1.cpu(), d2) < 1e-4
        f1, f2 = ours.get_features(x.to(DEV)), ref.get_features(x)
        assert len(f1) == 6 and all(rel_err(a.cpu(), b) < 1e-4 for a, b in zip(f1, f2))
        assert rel_err(ours(x.to(DEV)).cpu(), ref(x)) < 1e-4
    assert len(ours.parameters()) == len(ref.parameters()) and all(p.is_cuda for p in ours.parameters())
    ours.train()
    assert seg.training and disc.training
    opt = torch.optim.Adam(ours.parameters(), lr=1e-4)          # the reference's optimizer over both members
    seg_pred, dom = ours(x.to(DEV), domain_adaptation=True)
    (seg_pred.mean() + dom.mean()).backward()
    opt.step()
