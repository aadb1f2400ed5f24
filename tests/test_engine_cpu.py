"""CPU: host logic.  The engine's hand-written backward (tape, residual / skip gradient folding, flat
parameter store) is checked against autograd of the oracle U-Net by swapping the CUDA op layer for the
oracle's torch restatements (test-only monkeypatch; the product has no such switch), in float64 so that
ReLU-mask flips from fp32 round-off cannot hide logic errors.  Also: state_dict compatibility and the
gloo world_size-2 gradient synchronisation."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, rel_err
from oracle import ref_ops
from oracle.ref_unet import RefUnet
from oracle.ref_discriminator import RefDomainDiscriminator


@pytest.fixture()
def cpu_engine(monkeypatch):
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200 import unet, engine, discriminator
    for mod in (unet, engine, discriminator):
        monkeypatch.setattr(mod, "ops", ref_ops)
    monkeypatch.setattr(unet.Unet, "_prepare", lambda self, device: self._store.ensure_flat(device))
    monkeypatch.setattr(discriminator.DomainDiscriminator, "_prepare",
                        lambda self, device: self._store.ensure_flat(device))
    ref_ops.set_precision(torch.float64)
    yield U
    ref_ops.set_precision(torch.float32)


@pytest.mark.parametrize("enc", ["resnet34", "resnet50"])
def test_unet_forward_backward_logic(cpu_engine, enc):
    U = cpu_engine
    torch.manual_seed(0)
    ref = RefUnet(enc, classes=5)
    m = U.Unet(enc, classes=5, compute_dtype=torch.float64)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    m.load_state_dict(ref.state_dict())
    ref = ref.double()
    x = torch.randn(2, 3, 64, 64, requires_grad=True)
    x2 = x.detach().double().requires_grad_()
    t = torch.randint(0, 5, (2, 64, 64))
    y, yr = m(x), ref(x2)
    assert y.shape == (2, 5, 64, 64) and rel_err(y.detach(), yr.detach()) < 1e-10
    F.cross_entropy(y.double(), t).backward()
    F.cross_entropy(yr, t).backward()
    for (n, p), (_, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, p2.grad) < 1e-5, n
    assert rel_err(x.grad, x2.grad) < 1e-5
    for (n, b), (_, b2) in zip(m.named_buffers(), ref.named_buffers()):
        assert rel_err(b, b2) < 1e-5, n  # running statistics and num_batches_tracked
    # state_dict round trip into the oracle (checkpoint ABI)
    ref2 = RefUnet(enc, classes=5)
    ref2.load_state_dict(m.state_dict())


def test_encoder_decoder_head_standalone(cpu_engine):
    """model.encoder(x) / model.decoder(*features) / model.segmentation_head(d): reference uda.py:64-68,
    domain_model.py:52-53."""
    U = cpu_engine
    torch.manual_seed(1)
    ref = RefUnet("resnet34", classes=3)
    m = U.Unet("resnet34", classes=3, compute_dtype=torch.float64)
    m.load_state_dict(ref.state_dict())
    ref = ref.double()
    x = torch.randn(1, 3, 64, 64)
    feats = m.encoder(x)
    assert len(feats) == 6 and m.encoder.out_channels == (3, 64, 64, 128, 256, 512)
    assert [tuple(f.shape[1:]) for f in feats] == [(3, 64, 64), (64, 32, 32), (64, 16, 16), (128, 8, 8), (256, 4, 4), (512, 2, 2)]
    z = m.segmentation_head(m.decoder(*feats))
    zr = ref(x.double())
    assert rel_err(z.detach(), zr.detach()) < 1e-6
    (z.double() ** 2).sum().backward()
    (zr ** 2).sum().backward()
    for (n, p), (_, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, p2.grad) < 1e-5, n


def test_eval_mode_uses_running_stats(cpu_engine):
    U = cpu_engine
    torch.manual_seed(2)
    ref = RefUnet("resnet34", classes=4)
    for mod in ref.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    m = U.Unet("resnet34", classes=4, compute_dtype=torch.float64)
    m.load_state_dict(ref.state_dict())
    ref = ref.double().eval()
    m.eval()
    x = torch.randn(1, 3, 32, 32)
    with torch.no_grad():
        assert rel_err(m(x), ref(x.double())) < 1e-6


def test_discriminator_logic(cpu_engine):
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    torch.manual_seed(3)
    ref = RefDomainDiscriminator()
    d = DomainDiscriminator(compute_dtype=torch.float64)
    assert list(d.state_dict().keys()) == list(ref.state_dict().keys())
    d.load_state_dict(ref.state_dict())
    ref = ref.double()
    x = torch.randn(3, 3, 64, 64, requires_grad=True)
    x2 = x.detach().double().requires_grad_()
    y, yr = d(x), ref(x2)
    assert y.shape == (3, 1) and rel_err(y.detach(), yr.detach()) < 1e-6
    (y.double() * torch.tensor([[1.0], [-2.0], [0.5]])).sum().backward()
    (yr * torch.tensor([[1.0], [-2.0], [0.5]])).sum().backward()
    for (n, p), (_, p2) in zip(d.named_parameters(), ref.named_parameters()):
        # conv biases in front of a BatchNorm have an analytically zero gradient: compare absolutely
        if p2.grad.abs().max() < 1e-12:
            assert p.grad.abs().max() < 1e-6, n
        else:
            assert rel_err(p.grad, p2.grad) < 1e-4, n
    assert rel_err(x.grad, x2.grad) < 1e-4


def test_fused_adam_matches_torch_adam(cpu_engine, monkeypatch):
    """Optimizer semantics in isolation: both optimizers are fed the SAME gradients (Adam normalises
    gradients, so round-off-level differences in tiny gradients would otherwise be amplified to O(lr))."""
    U = cpu_engine
    from uda_aerial_semantic_segmentation_research_b200 import optim
    monkeypatch.setattr(optim, "ops", ref_ops)
    torch.manual_seed(4)
    ref = RefUnet("resnet18", classes=3)
    m = U.Unet("resnet18", classes=3, compute_dtype=torch.float64)
    m.load_state_dict(ref.state_dict())
    opt = optim.FusedAdam(m, lr=1e-3, weight_decay=1e-2)
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    x, t = torch.randn(2, 3, 32, 32), torch.randint(0, 3, (2, 32, 32))
    for step in range(3):
        opt.zero_grad()
        F.cross_entropy(m(x).double() * (1 + step), t).backward()
        stolen = all(p.grad.data_ptr() == m._store.grad.data_ptr() + 4 * m._store.offsets[id(p)]
                     for p in m.parameters())
        assert stolen, "autograd should adopt the flat-buffer views as .grad (zero-copy)"
        for p, p2 in zip(m.parameters(), ref.parameters()):
            p2.grad = p.grad.detach().clone().contiguous()
        opt.step(); opt_r.step()
        for (n, p), (_, p2) in zip(m.named_parameters(), ref.named_parameters()):
            assert rel_err(p.detach(), p2.detach()) < 2e-5, (step, n)  # fp32 round-off of the update


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200 import unet, engine, ddp
    unet.ops = engine.ops = ref_ops
    unet.Unet._prepare = lambda self, device: self._store.ensure_flat(device)
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must make them identical
    m = U.Unet("resnet18", classes=3, compute_dtype=torch.float32)
    sync = ddp.GradSync(m, bucket_mb=1.0)
    torch.manual_seed(7)
    x = torch.randn(2 * world, 3, 32, 32)
    t = torch.randint(0, 3, (2 * world, 32, 32))
    xs, ts = x[2 * rank:2 * rank + 2], t[2 * rank:2 * rank + 2]
    F.cross_entropy(m(xs), ts).backward()
    flat = m._store.grad.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    w0 = m._store.flat.clone()
    wl = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(wl, w0)
    if rank == 0:
        q.put((all(torch.equal(g, gathered[0]) for g in gathered), all(torch.equal(w, wl[0]) for w in wl),
               float(flat.abs().sum()), len(sync._plan(m._store)[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gloo_world2():
    """N>1 path on CPU: parameters are broadcast, bucketed all-reduce leaves identical averaged gradients."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_grads, same_weights, gsum, nb = q.get(timeout=300)
    for p in procs:
        p.join(60)
    assert same_grads and same_weights and gsum > 0 and nb > 1


def test_shadow_freshness_key_tracks_parameter_writes(cpu_engine):
    """parameters are views of the flat buffer with their OWN version counters — the key that
    decides whether the bf16 shadow weights are stale must move on in-place writes through the nn.Parameter
    (torch.optim.Adam as at reference train.py:461, load_state_dict as at phase_manager.py:140, p.copy_)."""
    U = cpu_engine
    torch.manual_seed(5)
    m = U.Unet("resnet18", classes=3, compute_dtype=torch.float64)
    st = m._store
    st.ensure_flat(torch.device("cpu"))
    k0 = st.param_version()
    with torch.no_grad():
        next(m.parameters()).add_(1.0)
    k1 = st.param_version()
    assert k1 != k0
    m.load_state_dict(RefUnet("resnet18", classes=3).state_dict())
    k2 = st.param_version()
    assert k2 != k1
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    F.cross_entropy(m(torch.randn(1, 3, 32, 32)).double(), torch.randint(0, 3, (1, 32, 32))).backward()
    opt.step()
    assert st.param_version() != k2
    # the parameters still live in the flat buffer after all of that (no silent re-allocation)
    base = st.flat.data_ptr()
    assert all(p.data_ptr() == base + st.flat.element_size() * st.offsets[id(p)] for p in st.params)


def test_decoder_standalone_with_detached_features(cpu_engine):
    """model.decoder(*features) on features that need no gradient (frozen / no_grad encoder),
    followed by backward, must train the decoder instead of raising in upcat's backward."""
    U = cpu_engine
    torch.manual_seed(6)
    ref = RefUnet("resnet18", classes=3)
    m = U.Unet("resnet18", classes=3, compute_dtype=torch.float64)
    m.load_state_dict(ref.state_dict())
    ref = ref.double()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        feats = m.encoder(x)
        feats_r = ref.encoder(x.double())
    assert not any(f.requires_grad for f in feats)
    z = m.segmentation_head(m.decoder(*feats))
    zr = ref.segmentation_head(ref.decoder(*feats_r))
    assert rel_err(z.detach(), zr.detach()) < 1e-6
    (z.double() ** 2).sum().backward()
    (zr ** 2).sum().backward()
    got = dict(m.named_parameters())
    for n, p2 in ref.named_parameters():
        if n.startswith("encoder."):
            continue
        assert rel_err(got[n].grad, p2.grad) < 1e-5, n


def test_fused_adam_is_a_torch_optimizer(cpu_engine, monkeypatch):
    """the reference trainers read optimizer.param_groups[0]['lr'] (train.py:361,
    adversarial_trainer.py:58), LR schedulers write it, and checkpoints hold optimizer.state_dict()
    (train.py:496,678) — FusedAdam must honour all three, and a restored optimizer must continue identically."""
    U = cpu_engine
    from uda_aerial_semantic_segmentation_research_b200 import optim
    monkeypatch.setattr(optim, "ops", ref_ops)
    torch.manual_seed(8)
    m = U.Unet("resnet18", classes=3, compute_dtype=torch.float64)
    opt = optim.FusedAdam(m, lr=1e-3)
    assert isinstance(opt, torch.optim.Optimizer)
    assert opt.param_groups[0]["lr"] == 1e-3 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    x, t = torch.randn(1, 3, 32, 32), torch.randint(0, 3, (1, 32, 32))

    def one_step(o):
        o.zero_grad()
        F.cross_entropy(m(x).double(), t).backward()
        o.step()

    one_step(opt)
    sched.step()
    assert opt.param_groups[0]["lr"] == 5e-4 and opt.lr == 5e-4
    one_step(opt)
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 2 and sd["param_groups"][0]["lr"] == 5e-4
    w = m._store.flat.clone()
    one_step(opt)
    w_next = m._store.flat.clone()
    # restore parameters + optimizer state into a fresh optimizer: the third step must repeat exactly
    with torch.no_grad():
        m._store.flat.copy_(w)
    opt2 = optim.FusedAdam(m, lr=123.0)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 5e-4 and opt2.step_count == 2
    one_step(opt2)
    assert torch.equal(m._store.flat, w_next)
    # step() without a backward after zero_grad() is refused (no stale gradients)
    opt2.zero_grad()
    with pytest.raises(RuntimeError):
        opt2.step()
    with pytest.raises(TypeError):
        optim.FusedAdam(list(m.parameters()))
