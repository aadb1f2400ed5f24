"""conv + BatchNorm + activation (+ residual) as ONE launch (``uda_conv2d_tc_fwd_bn_act``, north-star "BatchNorm+ReLU
fused in the epilogue"): the accumulators stay in tensor memory across a grid-wide barrier between the statistics pass
and the normalise pass.  Checked at the benchmarked shapes (B = 16, 512 x 512 input) of every layer family it serves

  * against the two-launch form it replaces (``uda_conv2d_tc_fwd`` with epilogue statistics + ``uda_bn_apply_fused``):
    z bit-identical, statistics equal up to the order of the fp64 atomics, a within one bf16 ulp;
  * against the fp32 oracle ops executed on the device (``oracle/ref_ops``: F.conv2d + batch-norm restatement of
    ``aten::batch_norm`` under the model of ``/root/reference/src/models/train.py:572-577``), bf16 tolerance 1e-2;
  * the running statistics are updated exactly once; shapes whose tiles do not fit one wave's TMEM decline (None).
"""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B = 16
CASES = [
    # name, H(=W) of the input, Cin, Cout, k, stride, residual, addend, slope
    ("layer2", 64, 128, 128, 3, 1, True, False, 0.0),
    ("l2.0_s2", 128, 64, 128, 3, 2, False, False, 0.0),
    ("l2_ds_1x1_s2", 128, 64, 128, 1, 2, False, False, 1.0),
    ("layer3_phalo", 32, 256, 256, 3, 1, True, False, 0.0),
    ("l3.0_s2", 64, 128, 256, 3, 2, False, False, 0.0),
    ("layer4", 16, 512, 512, 3, 1, True, False, 0.0),
    ("l4.0_s2", 32, 256, 512, 3, 2, False, False, 0.0),
    ("dec0.c1_phalo", 32, 768, 256, 3, 1, False, False, 0.0),
    ("dec1.c1_skip_half", 64, 128, 128, 3, 1, False, True, 0.0),
    ("leaky", 32, 128, 256, 3, 1, False, False, 0.2),
]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).bfloat16()


def _bn_params(C):
    g = torch.Generator(device=DEV).manual_seed(7)
    gamma = 1.0 + 0.2 * torch.randn(C, generator=g, device=DEV)
    beta = 0.1 * torch.randn(C, generator=g, device=DEV)
    return gamma, beta, 0.05 * torch.randn(C, generator=g, device=DEV), 1.0 + 0.1 * torch.rand(C, generator=g, device=DEV)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_conv_bn_act_matches_two_launch_form_and_oracle(case):
    from uda_aerial_semantic_segmentation_research_b200 import ops
    name, H, Cin, Cout, k, s, has_res, has_add, slope = case
    p = (k - 1) // 2
    Ho = H // s
    x = _rand((B, H, H, Cin), 1)
    w = _rand((Cout, k, k, Cin), 2, (k * k * Cin) ** -0.5)
    res = _rand((B, Ho, Ho, Cout), 3) if has_res else None
    add = _rand((B, Ho, Ho, Cout), 4) if has_add else None
    gamma, beta, rm0, rv0 = _bn_params(Cout)
    # ---- fused: one launch
    rm, rv = rm0.clone(), rv0.clone()
    slot = torch.zeros(2 * Cout + 1, dtype=torch.float64, device=DEV)
    out = ops.conv_bn_act_fused(x, w, slot, gamma, beta, rm, rv, 1e-5, 0.1, slope, s, p, addend=add, residual=res)
    assert out is not None, f"{name}: expected the fused launch to accept this shape"
    z, a, mean, rstd, scale, shift = out
    torch.cuda.synchronize()
    # ---- the two-launch form it replaces
    rm2, rv2 = rm0.clone(), rv0.clone()
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    z2 = ops.conv_fwd_add(x, w, add, bn_sums=sums, stride=s, pad=p) if has_add else ops.conv_fwd(x, w, None, s, p, bn_sums=sums)
    a2, mean2, rstd2, scale2, shift2 = ops.bn_apply_fused(z2, sums, gamma, beta, rm2, rv2, 1e-5, 0.1, res, slope)
    assert torch.equal(z, z2), name
    assert rel_err(slot[:2 * Cout], sums) < 1e-6
    for u, v in ((mean, mean2), (rstd, rstd2), (scale, scale2), (shift, shift2), (rm, rm2), (rv, rv2)):
        assert float((u - v).abs().max()) <= 1e-5 * float(v.abs().max()) + 1e-7, name
    # one bf16 ulp (2^-8 relative) where the statistics differ in the last bits
    d = (a.float() - a2.float()).abs()
    assert float((d - 2 ** -7 * a2.float().abs()).max()) <= 1e-6, name
    assert float((d > 0).float().mean()) < 1e-2, name
    # ---- fp32 oracle on the device
    zr = R.conv_fwd(x.float(), w.float(), None, s, p)
    if has_add:
        zr = zr + add.float()
    assert rel_err(z.float(), zr) < 1e-2
    zf = z.float()                                    # BatchNorm of the stored (bf16) pre-normalisation tensor
    m = zf.mean((0, 1, 2))
    var = zf.var((0, 1, 2), unbiased=False)
    y = (zf - m) * torch.rsqrt(var + 1e-5) * gamma + beta
    if has_res:
        y = y + res.float()
    ar = torch.where(y > 0, y, y * slope)
    assert rel_err(a.float(), ar) < 1e-2, name
    assert rel_err(mean, m) < 1e-4 and rel_err(rstd, torch.rsqrt(var + 1e-5)) < 1e-4
    n = B * Ho * Ho
    assert rel_err(rm, 0.9 * rm0 + 0.1 * m) < 1e-4
    assert rel_err(rv, 0.9 * rv0 + 0.1 * var * n / (n - 1)) < 1e-4


def test_fused_conv_bn_act_declines_non_resident_shapes():
    from uda_aerial_semantic_segmentation_research_b200 import ops
    for H, C in ((128, 64), (256, 32)):     # layer1 / decoder block 3 at B = 16: far more tiles than one wave holds
        x = _rand((B, H, H, C), 1)
        w = _rand((C, 3, 3, C), 2, 0.05)
        gamma, beta, rm, rv = _bn_params(C)
        slot = torch.zeros(2 * C + 1, dtype=torch.float64, device=DEV)
        assert ops.conv_bn_act_fused(x, w, slot, gamma, beta, rm, rv, 1e-5, 0.1, 0.0, 1, 1) is None
        assert float(slot.abs().sum()) == 0.0      # nothing was launched


def test_network_step_with_and_without_fused_bn_apply(monkeypatch):
    """Whole U-Net r34 training forward + backward at B=4, 256x256 (layer3/4 and decoder block 0 are TMEM-resident
    there): logits, loss and every parameter gradient of the fused engine path equal the two-launch path up to bf16
    rounding noise of the statistics."""
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200 import ops
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
    torch.manual_seed(0)
    model = U.Unet("resnet34", encoder_weights=None, in_channels=3, classes=24).to(DEV).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 256, 256, generator=g).to(DEV)
    t = torch.randint(0, 24, (4, 256, 256), generator=g).to(DEV)
    crit = CrossEntropyLoss()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    res = []
    for fuse in (True, False):
        monkeypatch.setattr(ops, "FUSE_BN_APPLY", fuse)
        model.load_state_dict(state)
        model.zero_grad(set_to_none=True)
        l0 = ops.LAUNCHES
        logits = model(x)
        loss = crit(logits, t)
        loss.backward()
        torch.cuda.synchronize()
        res.append((logits.detach().clone(), float(loss), [p.grad.clone() for p in model.parameters()],
                    ops.LAUNCHES - l0, {k: v.clone() for k, v in model.state_dict().items() if "running" in k}))
    (lf, lossf, gf, nf, rsf), (lu, lossu, gu, nu, rsu) = res
    assert nf < nu, "the fused path must issue fewer launches"
    # (train-mode BatchNorm over 47 layers amplifies one-ulp differences: DESIGN.md 4, 'chaotic in bf16')
    assert rel_err(lf, lu) < 1.5e-1
    assert abs(lossf - lossu) < 2e-3 * abs(lossu)
    num = sum(float(((a - b).double() ** 2).sum()) for a, b in zip(gf, gu))
    den = sum(float((b.double() ** 2).sum()) for b in gu)
    assert (num / den) ** 0.5 < 1.5e-1
    for k in rsf:
        assert rel_err(rsf[k], rsu[k]) < 2e-2, k


# ---- BatchNorm backward as ONE launch (reduce -> grid barrier -> apply, tiles resident in shared memory) -------------
BWD_CASES = [
    # name, B, H(=W), C, residual (mask from a, dres written), pre-existing residual gradient, slope
    ("layer3_conv1", 16, 32, 256, False, False, 0.0),
    ("layer3_conv2_res", 16, 32, 256, True, False, 0.0),
    ("layer3_res_accumulate", 16, 32, 256, True, True, 0.0),
    ("layer4_conv2_res", 16, 16, 512, True, False, 0.0),
    ("dec0", 16, 32, 256, False, False, 0.0),
    ("leaky_small", 4, 32, 128, False, False, 0.2),
    ("two_launch_path_layer2", 16, 64, 128, True, False, 0.0),     # 7 tiles per CTA: always reduce + apply
]


@pytest.mark.parametrize("case", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_bn_backward_single_launch_matches_oracle(case, monkeypatch):
    """``uda_bn_bwd`` with UDA_B200_BN_BWD_MERGED=1 (opt-in, read on every call) runs ``bn_bwd_merged_stream_kernel`` on
    L2-sized tensors; reference = ``oracle/ref_ops.bn_bwd``
    (the restatement of aten::native_batch_norm_backward + threshold_backward the reference's autograd executes under
    ``loss.backward()``, ``/root/reference/src/models/train.py:343``) in fp32 ON THE DEVICE.  Run twice: the kernel must
    leave its workspace (sums, barrier counters) zeroed."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    monkeypatch.setenv("UDA_B200_BN_BWD_MERGED", "1")
    name, Bn, H, C, residual, accumulate, slope = case
    x = _rand((Bn, H, H, C), 11, 1.5) + 0.25
    dy = _rand((Bn, H, H, C), 12)
    res = _rand((Bn, H, H, C), 13) if residual else None
    gamma, beta, rm, rv = _bn_params(C)
    mean, rstd, scale, shift = ops.bn_stats(x, gamma, beta, rm, rv, 1e-5, 0.1)
    a = ops.bn_apply(x, scale, shift, res, slope)
    use_a = residual and slope != 1.0
    zm = (not residual) and slope != 1.0
    for rep in range(2):
        dres0 = _rand((Bn, H, H, C), 14 + rep) if accumulate else None
        dg, db = torch.full((C,), 0.5, device=DEV), torch.full((C,), -0.25, device=DEV)
        dres = dres0.clone() if accumulate else (torch.empty_like(x) if residual else None)
        dx = ops.bn_bwd(dy, x, a if use_a else None, gamma, mean, rstd, slope, dg, db, dres=dres,
                        dres_accumulate=accumulate, scale=scale if zm else None, shift=shift if zm else None)
        dgr, dbr = torch.full((C,), 0.5, device=DEV), torch.full((C,), -0.25, device=DEV)
        dresr = dres0.clone() if accumulate else (torch.empty_like(x) if residual else None)
        dxr = R.bn_bwd(dy, x, a if use_a else None, gamma, mean, rstd, slope, dgr, dbr, dres=dresr,
                       dres_accumulate=accumulate, scale=scale if zm else None, shift=shift if zm else None)
        assert rel_err(dx.float(), dxr.float()) < 1e-2, (name, rep)
        assert rel_err(dg, dgr) < 1e-3 and rel_err(db, dbr) < 1e-3, (name, rep)
        if residual:
            assert rel_err(dres.float(), dresr.float()) < 1e-2, (name, rep)


POOL_CASES = [
    # B, H, W, C, slope   (the stem at the benchmarked shape first; strips crossing image boundaries, fewer pooled rows
    # than SMs, one-row images, leaky slope, NaN propagation)
    (16, 256, 256, 64, 0.0),
    (3, 38, 52, 64, 0.0),
    (5, 2, 4, 8, 0.0),
    (2, 64, 96, 128, 0.2),
    (1, 300, 16, 16, 0.0),
    (7, 6, 64, 256, 0.0),
]


@pytest.mark.parametrize("case", POOL_CASES, ids=["x".join(map(str, c[:4])) for c in POOL_CASES])
def test_bn_apply_maxpool_fused_matches_two_launch_form_and_torch(case):
    """Stem tail bn1 -> relu -> maxpool as ONE pass (``uda_bn_apply_maxpool_fused``): a / mean / rstd / scale / shift /
    running statistics bit-identical with ``uda_bn_apply_fused``, pooled values and winning taps bit-identical with
    ``uda_maxpool3x3s2_fwd`` of that a, pooled values equal to ``F.max_pool2d`` (``torchvision`` ResNet stem under the
    model of ``/root/reference/src/models/train.py:572-577``)."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    Bn, H, W, C, slope = case
    z = _rand((Bn, H, W, C), 11, 1.5)
    if (Bn, H) == (3, 38):
        z[1, 5, 7, 3] = float("nan")
        z[2, 37, 51, 9] = float("nan")
    zf = z.double()
    sums = torch.cat([torch.nan_to_num(zf).sum((0, 1, 2)), (torch.nan_to_num(zf) ** 2).sum((0, 1, 2))]).contiguous()
    gamma, beta, rm0, rv0 = _bn_params(C)
    rm1, rv1 = rm0.clone(), rv0.clone()
    a1, mean1, rstd1, sc1, sf1 = ops.bn_apply_fused(z, sums, gamma, beta, rm1, rv1, 1e-5, 0.1, None, slope)
    y1, i1 = ops.maxpool_fwd(a1)
    rm2, rv2 = rm0.clone(), rv0.clone()
    r = ops.bn_apply_maxpool_fused(z, sums, gamma, beta, rm2, rv2, 1e-5, 0.1, slope)
    assert r is not None, "fused stem tail declined a supported shape"
    a2, mean2, rstd2, sc2, sf2, y2, i2 = r
    torch.cuda.synchronize()
    for u, v in ((sc1, sc2), (sf1, sf2)):   # what the normalisation uses: bit-identical
        assert torch.equal(u, v)
    for u, v in ((mean1, mean2), (rstd1, rstd2), (rm1, rm2), (rv1, rv2)):   # published statistics: the small shapes run
        torch.testing.assert_close(u, v, rtol=1e-6, atol=1e-7)              # another kernel's copy of the same fp32 lines
    assert torch.equal(a1.view(torch.int16), a2.view(torch.int16))
    assert torch.equal(y1.view(torch.int16), y2.view(torch.int16))
    assert torch.equal(i1, i2)
    ref = torch.nn.functional.max_pool2d(a2.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(torch.nan_to_num(ref, nan=-7.0), torch.nan_to_num(y2.float(), nan=-7.0))


def test_bn_apply_maxpool_fused_declines_odd_sizes_and_wide_rows():
    from uda_aerial_semantic_segmentation_research_b200 import ops
    for shape in ((2, 9, 8, 16), (2, 8, 10, 4096), (1, 4, 1024, 64)):
        z = _rand(shape, 5)
        C = shape[-1]
        sums = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
        gamma, beta, rm, rv = _bn_params(C)
        rm0 = rm.clone()
        assert ops.bn_apply_maxpool_fused(z, sums, gamma, beta, rm, rv) is None
        assert torch.equal(rm, rm0)
