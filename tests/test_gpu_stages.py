"""Teacher-forced bf16 parity of EVERY stage of the U-Net (the 2e-2 gate must hold on layer2, layer3,
layer4 and each decoder block, not only on the shallow chains).

The whole 47-conv network at random initialisation is chaotic in bf16 (DESIGN.md "bf16 parity"), so the
north-star's 2e-2 logit gate is enforced stage by stage: every stage of the CUDA path is fed the bf16 reference's
own input features ("teacher forcing") and compared with the bf16 reference's output of that stage — forward, hard
2e-2 gate — and with the reference's autograd gradients for an identical upstream gradient — backward.

Backward gate.  Train-mode BatchNorm + ReLU make the gradients of even ONE residual block discontinuous in the
forward rounding: on the oracle alone, the bf16-storage forward vs the fp32 forward of the same block (same inputs,
same upstream gradient, fp32 autograd in both) moves dX / dW / dgamma by 4-9e-2 in L2 and up to 5e-1 in max-norm
(ReLU masks flip where a rounded pre-activation changes sign), and a whole stage by 1.0-1.8e-1 (measured, see
DESIGN.md "bf16 parity").  No bf16 implementation can therefore meet a fixed 1e-2 gradient gate at stage level; the
gate asserted here is that the CUDA path is no further from the bf16 reference than 1.5x the distance of the bf16
reference from the fp32 reference ("natural spread", computed in the test), in L2, and the per-kernel gradient gates
(1e-3 / 1e-2 on identical inputs) stay in test_gpu_layers.py / test_gpu_conv_tc.py / test_gpu_bench_shapes.py.
Stages whose rounding SEQUENCE equals the reference's (encoder layers, materialised decoder block 0) sit well inside the
spread (0.3-0.8x); the decoder blocks that run conv1 as conv_transpose4x4(x) + conv3x3(skip) round differently (summed
4x4 weights, bf16 partial sum) and therefore sit AT the spread (0.9-1.1x): a different, equally valid bf16 evaluation.
BASELINE configs[0] shape: batch 2 @ 256 x 256."""
import pytest
import torch

from conftest import rel_err, l2_err
from oracle.ref_unet import RefUnet, emulate_bf16

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _models(seed=0, classes=24, with_fp32=False):
    import uda_aerial_semantic_segmentation_research_b200 as U
    torch.manual_seed(seed)
    ref = RefUnet("resnet34", classes=classes)
    m = U.Unet("resnet34", classes=classes, compute_dtype=torch.bfloat16)
    m.load_state_dict(ref.state_dict())
    if with_fp32:
        return m.to(DEV).train(), emulate_bf16(ref).train(), ref.train()
    return m.to(DEV).train(), emulate_bf16(ref).train()


def _run_stage(m, run, inputs, gout):
    """Run ``run(ctx, *vars)`` of the CUDA engine on NCHW fp32 ``inputs`` with the tape on, back-propagate the NCHW
    upstream gradient ``gout``; returns (output NCHW fp32, [dX NCHW], {param name: grad}) — all on the CPU."""
    from uda_aerial_semantic_segmentation_research_b200 import ops, engine as E
    st = m._store
    m._prepare(torch.device(DEV))
    tape = E.Tape()
    ctx = E.Ctx(st, torch.bfloat16, True, tape)
    vs = [E.Var(ops.nchw_to_nhwc(t.to(DEV).contiguous(), torch.bfloat16)) for t in inputs]
    out = run(ctx, *vs)
    ctx.finish_forward()
    y = ops.nhwc_to_nchw(out.t).cpu()
    st.new_grad()
    out.g = ops.nchw_to_nhwc(gout.to(DEV).contiguous(), torch.bfloat16)
    tape.backward()
    dxs = [ops.nhwc_to_nchw(v.g).cpu() for v in vs]
    names = {id(p): n for n, p in m.named_parameters()}
    grads = {names[id(p)]: g.detach().cpu().clone() for p, g in zip(st.params, st.grad_views())}
    return y, dxs, grads


def _ref_stage(fn, inputs, gout):
    xs = [t.clone().requires_grad_() for t in inputs]
    y = fn(*xs)
    y.backward(gout)
    return y.detach(), [x.grad for x in xs]


def _ref_grads(model, prefix):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if n.startswith(prefix) and p.grad is not None}


def _check(tag, y, yr, dxs, dxrs, grads, rgrads, dx32, g32):
    """Errors of the CUDA path vs the bf16 reference, next to the natural spread (bf16 reference vs fp32 reference)."""
    e_y = rel_err(y, yr)
    e_dx, n_dx = max(l2_err(a, b) for a, b in zip(dxs, dxrs)), max(l2_err(a, b) for a, b in zip(dxrs, dx32))
    e_dx_max = max(rel_err(a, b) for a, b in zip(dxs, dxrs))
    e_w = max(l2_err(grads[n], g) for n, g in rgrads.items() if g.dim() == 4)
    n_w = max(l2_err(g, g32[n]) for n, g in rgrads.items() if g.dim() == 4)
    e_bn = max(l2_err(grads[n], g) for n, g in rgrads.items() if g.dim() == 1)
    n_bn = max(l2_err(g, g32[n]) for n, g in rgrads.items() if g.dim() == 1)
    print(f"{tag:8s} fwd {e_y:.2e} | L2 vs bf16 reference (natural spread): dX {e_dx:.2e} ({n_dx:.2e})  "
          f"dW {e_w:.2e} ({n_w:.2e})  dgamma/dbeta {e_bn:.2e} ({n_bn:.2e}) | dX max-norm {e_dx_max:.2e}")
    return e_y, (e_dx, n_dx), (e_w, n_w), (e_bn, n_bn)


def test_every_stage_teacher_forced():
    m, ref16, ref32 = _models(with_fp32=True)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 256, 256, generator=g)
    with torch.no_grad():
        f = ref16.encoder(x)                       # bf16-valued features of the reference path
        dec_in = [f[5]]
        for i, blk in enumerate(ref16.decoder.blocks):
            dec_in.append(blk(dec_in[-1], f[4 - i] if i < 4 else None))
    from uda_aerial_semantic_segmentation_research_b200 import engine as E
    results = {}

    def grad_like(t, seed):
        return torch.randn(t.shape, generator=torch.Generator().manual_seed(seed)).bfloat16().float()

    # ---- encoder stages: layer1 (behind the max-pool), layer2, layer3, layer4 --------------------------------
    for li in range(1, 5):
        layer = getattr(m.encoder, f"layer{li}")
        rlayer = getattr(ref16.encoder, f"layer{li}")

        def run(ctx, v, layer=layer, li=li):
            if li == 1:
                v = E.maxpool(ctx, v)
            for blk in layer:
                v = blk.run(ctx, v)
            return v

        def rfn(t, rlayer=rlayer, li=li):
            return rlayer(ref16.encoder.maxpool(t) if li == 1 else t)

        def rfn32(t, li=li):
            return getattr(ref32.encoder, f"layer{li}")(ref32.encoder.maxpool(t) if li == 1 else t)

        ref16.zero_grad(); ref32.zero_grad()
        go = grad_like(f[li + 1], 100 + li)
        yr, dxr = _ref_stage(rfn, [f[li]], go)
        _, dx32 = _ref_stage(rfn32, [f[li]], go)
        y, dx, grads = _run_stage(m, run, [f[li]], go)
        pre = f"encoder.layer{li}."
        results[f"layer{li}"] = _check(f"layer{li}", y, yr, dx, dxr, grads, _ref_grads(ref16, pre), dx32, _ref_grads(ref32, pre))

    # ---- decoder blocks -----------------------------------------------------------------------------------------
    for i, (blk, rblk) in enumerate(zip(m.decoder.blocks, ref16.decoder.blocks)):
        skip = f[4 - i] if i < 4 else None
        ins = [dec_in[i]] + ([skip] if skip is not None else [])

        def run(ctx, *vs, blk=blk):
            return blk.run(ctx, vs[0], vs[1] if len(vs) > 1 else None)

        def rfn(*ts, rblk=rblk):
            return rblk(ts[0], ts[1] if len(ts) > 1 else None)

        def rfn32(*ts, i=i):
            return ref32.decoder.blocks[i](ts[0], ts[1] if len(ts) > 1 else None)

        ref16.zero_grad(); ref32.zero_grad()
        go = grad_like(dec_in[i + 1], 200 + i)
        yr, dxr = _ref_stage(rfn, ins, go)
        _, dx32 = _ref_stage(rfn32, ins, go)
        y, dx, grads = _run_stage(m, run, ins, go)
        pre = f"decoder.blocks.{i}."
        results[f"dec{i}"] = _check(f"dec{i}", y, yr, dx, dxr, grads, _ref_grads(ref16, pre), dx32, _ref_grads(ref32, pre))

    # ---- gates ---------------------------------------------------------------------------------------------------
    for tag, (e_y, dxe, we, bne) in results.items():
        assert e_y < 2e-2, (tag, "forward", e_y)          # north-star: 2e-2 in bf16, every stage
        for what, (err, nat) in (("dX", dxe), ("dW", we), ("dgamma/dbeta", bne)):
            # within 1.5x the distance of the bf16 reference from fp32 (module docstring); 2e-2 floor
            assert err < max(1.5 * nat, 2e-2), (tag, what, err, nat)


def test_head_teacher_forced():
    """Segmentation head (fp32 NCHW logits with bias) on the reference's decoder output: logits 2e-2 (measured ~1e-3),
    gradients 1e-3-class (one conv, no BatchNorm)."""
    m, ref16 = _models(seed=1)
    g = torch.Generator().manual_seed(7)
    d = torch.randn(2, 16, 256, 256, generator=g).bfloat16().float()
    go = torch.randn(2, 24, 256, 256, generator=g)
    dr = d.clone().requires_grad_()
    yr = ref16.segmentation_head(dr)
    yr.backward(go)
    dg = d.to(DEV).requires_grad_()
    y = m.segmentation_head(dg)
    y.backward(go.to(DEV))
    assert rel_err(y.detach().cpu(), yr.detach()) < 5e-3
    assert rel_err(dg.grad.cpu(), dr.grad) < 1e-2
    pg = dict(m.named_parameters()); rg = dict(ref16.named_parameters())
    assert rel_err(pg["segmentation_head.0.weight"].grad.cpu(), rg["segmentation_head.0.weight"].grad) < 5e-3
    # the bias gradient is a sum over 131072 bf16-rounded logit gradients of zero mean: 2^-9-class relative noise
    assert rel_err(pg["segmentation_head.0.bias"].grad.cpu(), rg["segmentation_head.0.bias"].grad) < 5e-3
