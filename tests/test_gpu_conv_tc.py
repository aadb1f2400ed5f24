"""GPU parity of the tcgen05/TMEM/TMA convolution against (a) the oracle's torch conv on the same bf16
inputs and (b) the FP32-pipe direct kernel.  bf16 output: 2^-8 relative per element; the fp32 NCHW edge
output is held to 2e-3 of the tensor's max (bf16 products, fp32 accumulation order)."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [
    # B, H, W, Cin, Cout, k, stride, pad            every U-Net r34 family at reduced spatial size
    (2, 32, 32, 64, 64, 3, 1, 1),       # layer1
    (2, 32, 32, 64, 128, 3, 2, 1),      # layer2.0 conv1 (stride 2)
    (2, 32, 32, 64, 128, 1, 2, 0),      # downsample 1x1 stride 2
    (2, 16, 16, 128, 128, 3, 1, 1),     # layer2
    (2, 8, 8, 256, 256, 3, 1, 1),       # layer3 (two images per tile)
    (4, 8, 8, 512, 512, 3, 1, 1),       # layer4: 4 N tiles
    (2, 8, 8, 768, 256, 3, 1, 1),       # decoder block 0 conv1 (concat input)
    (2, 16, 16, 192, 64, 3, 1, 1),      # decoder block 2 conv1
    (1, 64, 64, 128, 32, 3, 1, 1),      # decoder block 3 conv1 (N=32)
    (1, 64, 64, 32, 32, 3, 1, 1),       # KC=32 (64-byte swizzle)
    (1, 128, 128, 32, 16, 3, 1, 1),     # decoder block 4 conv1: N=16 padded to 32, TW=128
    (1, 128, 128, 16, 16, 3, 1, 1),     # KC=16 (32-byte swizzle)
    (1, 128, 128, 16, 24, 3, 1, 1),     # head: N=24, bias, NCHW fp32 edge
    (2, 64, 64, 64, 128, 4, 2, 1),      # discriminator layer 2 (4x4 stride 2)
    (1, 256, 256, 64, 64, 3, 1, 1),     # wide rows: TW=128, 2 tiles per row
    (2, 128, 128, 64, 64, 3, 1, 1),     # halo kernels: layer1 / dec2.c2
    (1, 128, 256, 128, 32, 3, 1, 1),    # halo: two channel chunks (dec3.c1), dgrad 32->128
    (1, 64, 128, 32, 32, 3, 1, 1),      # halo: H=64 rows
    (2, 16, 128, 16, 24, 3, 1, 1),      # halo: head-like, short image
]


def _ops():
    from uda_aerial_semantic_segmentation_research_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).bfloat16()


@pytest.mark.parametrize("cfg", SHAPES, ids=[f"B{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}_k{c[5]}s{c[6]}" for c in SHAPES])
def test_tc_forward_and_dgrad(cfg):
    ops = _ops()
    B, H, W, Cin, Cout, k, s, p = cfg
    assert ops.tc_supported(0, B, H, W, Cin, Cout, k, k, s, p), "shape expected on the tensor-core path"
    x = _rand((B, H, W, Cin), 1)
    w = _rand((Cout, k, k, Cin), 2, (k * k * Cin) ** -0.5)
    bias = torch.randn(Cout)
    xg, wg, bg = x.to(DEV), w.to(DEV), bias.to(DEV)
    yr = R.conv_fwd(x.float(), w.float(), bias, s, p)                    # fp32 reference on bf16-valued inputs
    y = ops.conv_fwd(xg, wg, bg, s, p)
    yd = ops.conv_fwd(xg, wg, bg, s, p, force_direct=True)
    e_tc, e_direct = rel_err(y.float().cpu(), yr), rel_err(yd.float().cpu(), yr)
    assert e_tc < 1e-2, (e_tc, e_direct)
    yn = ops.conv_fwd(xg, wg, bg, s, p, nchw_out=True)
    assert rel_err(yn.cpu(), yr.permute(0, 3, 1, 2)) < 2e-3
    y0 = ops.conv_fwd(xg, wg, None, s, p)
    assert rel_err(y0.float().cpu(), R.conv_fwd(x.float(), w.float(), None, s, p)) < 1e-2
    # BatchNorm batch statistics emitted by the epilogue == sums over the bf16 output it wrote
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    y1 = ops.conv_fwd(xg, wg, None, s, p, bn_sums=sums)
    assert torch.equal(y1, y0)
    yf = y0.double().reshape(-1, Cout)
    assert rel_err(sums[:Cout].cpu(), yf.sum(0).cpu()) < 1e-5 + 1e-4 * float(yf.abs().sum(0).max() / (yf.sum(0).abs().max() + 1e-9)) * 1e-2
    assert rel_err(sums[Cout:].cpu(), (yf * yf).sum(0).cpu()) < 1e-4
    g, b = torch.rand(Cout, device=DEV) + 0.5, torch.randn(Cout, device=DEV)
    rm1, rv1 = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    rm2, rv2 = rm1.clone(), rv1.clone()
    a1, mean1, rstd1, _sc1, _sh1 = ops.bn_apply_fused(y0, sums, g, b, rm1, rv1, 1e-5, 0.1, None, 0.0)
    mean2, rstd2, sc2, sh2 = ops.bn_stats(y0, g, b, rm2, rv2, 1e-5, 0.1)
    a2 = ops.bn_apply(y0, sc2, sh2, None, 0.0)
    assert rel_err(mean1, mean2) < 1e-4 and rel_err(rstd1, rstd2) < 1e-4
    assert rel_err(rm1, rm2) < 1e-4 and rel_err(rv1, rv2) < 1e-4
    assert rel_err(a1.float(), a2.float()) < 1e-2
    # dgrad (stride 1: one launch on flipped weights; stride 2: one launch per output parity)
    assert ops.tc_supported(1, B, H, W, Cin, Cout, k, k, s, p)
    dy = _rand(tuple(yr.shape), 3)
    dxr = R.conv_dgrad(dy.float(), w.float(), x.shape, s, p)
    dx = ops.conv_dgrad(dy.to(DEV), wg, x.shape, s, p)
    assert rel_err(dx.float().cpu(), dxr) < 1e-2
    add = _rand(tuple(x.shape), 4)
    acc = add.clone().to(DEV)
    out = ops.conv_dgrad(dy.to(DEV), wg, x.shape, s, p, addend=acc)
    assert out.data_ptr() == acc.data_ptr()
    assert rel_err(out.float().cpu(), dxr + add.float()) < 1e-2
    # wgrad (MN-major operands straight from the NHWC tensors), accumulating into fp32
    assert ops.tc_supported(2, B, H, W, Cin, Cout, k, k, s, p)
    dwr = R.conv_wgrad(dy.float(), x.float(), torch.zeros(Cout, k, k, Cin), s, p)
    dw = torch.zeros((Cout, k, k, Cin), device=DEV)
    ops.conv_wgrad(dy.to(DEV), xg, dw, s, p)
    e_w = rel_err(dw.cpu(), dwr)
    dwd = torch.zeros((Cout, k, k, Cin), device=DEV)
    ops.conv_wgrad(dy.to(DEV), xg, dwd, s, p, force_direct=True)
    assert e_w < 2e-3, (e_w, rel_err(dwd.cpu(), dwr))        # fp32 accumulation of exact bf16 products
    ops.conv_wgrad(dy.to(DEV), xg, dw, s, p)                 # accumulates
    assert rel_err(dw.cpu(), 2 * dwr) < 2e-3


@pytest.mark.parametrize("tile", ["1,128", "2,128", "1,256", "2,256"])
@pytest.mark.parametrize("cfg", [(4, 16, 16, 256, 256, 3, 1, 1), (2, 32, 32, 128, 256, 3, 2, 1), (4, 16, 16, 384, 128, 3, 1, 1)],
                         ids=["256to256", "128to256_s2", "384to128"])
def test_tile_shapes_of_the_wide_layers(cfg, tile, monkeypatch):
    """Every (pixel tile, channel tile) shape the persistent kernel's cost model can pick — 128/256 pixels x
    128/256 channels, the last one with single-buffered TMEM — against the fp32 reference (forward with the fused
    BatchNorm statistics, and dgrad with an addend)."""
    ops = _ops()
    B, H, W, Cin, Cout, k, s, p = cfg
    if tile.endswith("256") and Cout % 256:
        pytest.skip("channel tile wider than the layer")
    monkeypatch.setenv("UDA_B200_TC_TILE", tile)
    x = _rand((B, H, W, Cin), 11)
    w = _rand((Cout, k, k, Cin), 12, (k * k * Cin) ** -0.5)
    xg, wg = x.to(DEV), w.to(DEV)
    yr = R.conv_fwd(x.float(), w.float(), None, s, p)
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    y = ops.conv_fwd(xg, wg, None, s, p, bn_sums=sums)
    assert rel_err(y.float().cpu(), yr) < 1e-2
    yf = y.double().reshape(-1, Cout)
    assert rel_err(sums[Cout:].cpu(), (yf * yf).sum(0).cpu()) < 1e-4
    assert (sums[:Cout].cpu() - yf.sum(0).cpu()).abs().max() < 1e-3 * float(yf.abs().sum(0).max())
    dy = _rand(tuple(yr.shape), 13)
    dxr = R.conv_dgrad(dy.float(), w.float(), x.shape, s, p)
    add = _rand(tuple(x.shape), 14)
    acc = add.clone().to(DEV)
    out = ops.conv_dgrad(dy.to(DEV), wg, x.shape, s, p, addend=acc)
    assert rel_err(out.float().cpu(), dxr + add.float()) < 1e-2


@pytest.mark.parametrize("cfg", [(2, 32, 32, 128, 128), (2, 32, 32, 256, 256), (1, 64, 64, 128, 128), (3, 32, 32, 64, 64),
                                 (2, 16, 32, 64, 128), (1, 64, 64, 192, 64), (5, 24, 64, 64, 128)],
                         ids=lambda c: "B%d_%dx%d_%dto%d" % c)
def test_pitched_halo_kernel(cfg, monkeypatch):
    """conv_tc_phalo.cu (3x3 stride 1 on W = 32 / 64 images: pitched positions, one halo box for nine taps): forward
    with the fused BatchNorm statistics and dgrad with an addend against the fp32 reference and the persistent
    kernel; ragged cases: an odd number of blocks (last tile partly empty), a last block that is mostly junk
    positions, tiles that straddle images."""
    ops = _ops()
    B, H, W, Cin, Cout = cfg
    x = _rand((B, H, W, Cin), 31)
    w = _rand((Cout, 3, 3, Cin), 32, (9 * Cin) ** -0.5)
    xg, wg = x.to(DEV), w.to(DEV)
    yr = R.conv_fwd(x.float(), w.float(), None, 1, 1)
    monkeypatch.setenv("UDA_B200_TC_PHALO", "0")
    y_persist = ops.conv_fwd(xg, wg, None, 1, 1)
    monkeypatch.setenv("UDA_B200_TC_PHALO", "2")
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    y = ops.conv_fwd(xg, wg, None, 1, 1, bn_sums=sums)
    assert rel_err(y.float().cpu(), yr) < 1e-2
    assert rel_err(y.float().cpu(), y_persist.float().cpu()) < 4e-3      # same products, different summation order
    yf = y.double().reshape(-1, Cout)
    assert rel_err(sums[Cout:].cpu(), (yf * yf).sum(0).cpu()) < 1e-4
    assert (sums[:Cout].cpu() - yf.sum(0).cpu()).abs().max() < 1e-3 * float(yf.abs().sum(0).max())
    dy = _rand(tuple(yr.shape), 33)
    dxr = R.conv_dgrad(dy.float(), w.float(), x.shape, 1, 1)
    add = _rand(tuple(x.shape), 34)
    acc = add.clone().to(DEV)
    out = ops.conv_dgrad(dy.to(DEV), wg, x.shape, 1, 1, addend=acc)
    assert out.data_ptr() == acc.data_ptr()
    assert rel_err(out.float().cpu(), dxr + add.float()) < 1e-2


def test_weight_flip_transpose():
    ops = _ops()
    w = _rand((6, 3, 3, 4), 5)
    wt = ops.weight_flip_transpose(w.to(DEV)).cpu()
    ref = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()
    assert torch.equal(wt, ref)


def test_weight_flip_transpose_batch():
    """All conv weights of a network in one launch (64 x 64 shared-memory transposes): ragged channel counts
    (3, 24, 100), 1x1 / 3x3 / 4x4 / 7x7 taps, several tiles per weight."""
    ops = _ops()
    shapes = [(64, 7, 7, 3), (24, 3, 3, 16), (128, 1, 1, 64), (256, 3, 3, 192), (100, 4, 4, 72), (8, 3, 3, 8)]
    offs, rows, off = [], [], 0
    for O, KH, KW, I in shapes:
        offs.append(off)
        rows.append([off, O, I, KH, KW])
        off += (O * KH * KW * I + 7) // 8 * 8
    g = torch.Generator().manual_seed(3)
    base = torch.randn(off, generator=g).bfloat16()
    out = torch.zeros(off, dtype=torch.bfloat16, device=DEV)
    table = torch.tensor(rows, dtype=torch.int32, device=DEV)
    ops.weight_flip_transpose_batch(base.to(DEV), out, table)
    out = out.cpu()
    for (O, KH, KW, I), o in zip(shapes, offs):
        n = O * KH * KW * I
        w = base[o:o + n].view(O, KH, KW, I)
        ref = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()
        assert torch.equal(out[o:o + n].view(I, KH, KW, O), ref), (O, KH, KW, I)


def test_unsupported_shapes_fall_back_loudly():
    ops = _ops()
    from uda_aerial_semantic_segmentation_research_b200 import _lib
    assert not ops.tc_supported(0, 2, 16, 16, 3, 64, 7, 7, 2, 3)      # stem: Cin=3
    with pytest.raises(_lib.UdaError, match="not covered"):
        _lib.call("conv2d_tc_fwd", _lib.ptr(torch.zeros(8, device=DEV)), _lib.ptr(torch.zeros(8, device=DEV)), None,
                  _lib.ptr(torch.zeros(8, device=DEV)), None, None, _lib.ci(2), _lib.ci(16), _lib.ci(16), _lib.ci(3),
                  _lib.ci(64), _lib.ci(7), _lib.ci(7), _lib.ci(2), _lib.ci(3), None)


@pytest.mark.parametrize("cfg", [(2, 64, 64, 64, 7, 3), (2, 128, 256, 64, 7, 3), (2, 64, 64, 64, 4, 1), (16, 32, 32, 64, 7, 3)],
                         ids=["unet_stem_64", "unet_stem_128x256", "disc_stem_64", "unet_stem_b16"])
def test_stem_on_tensor_cores(cfg):
    """Cin=3 stems (U-Net 7x7 s2 p3, discriminator 4x4 s2 p1) through the padded 4-channel row view."""
    ops = _ops()
    B, H, W, Cout, K, pad = cfg
    if not ops.stem_supported(B, H, W, 3, Cout, K, 2, pad):
        pytest.skip("stem path disabled (UDA_B200_TC_PERSIST=0 debugging toggle)")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, H, W, generator=g).bfloat16().float()          # bf16-valued image
    w = _rand((Cout, K, K, 3), 12, 0.1)
    bias = torch.randn(Cout, generator=g)
    xs = ops.stem_pack_input(x.to(DEV), pad)
    assert xs.shape == (B, H + 8, W + 8, 4)
    assert torch.equal(xs[:, pad:pad + H, pad:pad + W, :3].float().cpu(), x.permute(0, 2, 3, 1))
    assert float(xs[..., 3].abs().max()) == 0 and float(xs[:, :pad].abs().max()) == 0
    ws = ops.stem_pack_weight(w.to(DEV))
    y = ops.stem_fwd(xs, ws, bias.to(DEV), H, W, K, pad)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    yr = R.conv_fwd(x_nhwc, w.float(), bias, 2, pad)
    assert rel_err(y.float().cpu(), yr) < 1e-2
    dy = _rand(tuple(yr.shape), 13)
    dw = torch.zeros((Cout, K, K, 3), device=DEV)
    ops.stem_wgrad(dy.to(DEV), xs, dw, H, W, K, pad)
    dwr = R.conv_wgrad(dy.float(), x_nhwc, torch.zeros(Cout, K, K, 3), 2, pad)
    assert rel_err(dw.cpu(), dwr) < 2e-3
    ops.stem_wgrad(dy.to(DEV), xs, dw, H, W, K, pad)
    assert rel_err(dw.cpu(), 2 * dwr) < 2e-3
