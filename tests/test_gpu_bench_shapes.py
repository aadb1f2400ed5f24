"""GPU parity AT THE BENCHMARKED SHAPE ("no parity test at the benchmarked shape"): every convolution
family of U-Net r34 at B = 16 and its true H x W for a 512 x 512 input — the shapes for which the tile / kernel
selection in the launchers (persistent 256x128 / 128x256 / 256x256 tiles, weight-stationary mode, halo rows,
pitched halo, pair kernels, multi-accumulator wgrad groups) is the one ``bench.py`` times.  Reference = the oracle op
(``oracle/ref_ops``: F.conv2d / torch.nn.grad in fp32, TF32 off) executed ON THE DEVICE on the same bf16 inputs.
Tolerances as in test_gpu_conv_tc.py: bf16 outputs 1e-2 of the tensor's max, fp32 weight gradients 2e-3."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, S = 16, 512
s2, s4, s8, s16, s32 = S // 2, S // 4, S // 8, S // 16, S // 32
FAMILIES = [
    # name, H(=W), Cin, Cout, k, stride
    ("layer1", s4, 64, 64, 3, 1), ("l2.0_s2", s4, 64, 128, 3, 2), ("l2_ds_1x1_s2", s4, 64, 128, 1, 2),
    ("layer2", s8, 128, 128, 3, 1), ("l3.0_s2", s8, 128, 256, 3, 2), ("layer3", s16, 256, 256, 3, 1),
    ("l4.0_s2", s16, 256, 512, 3, 2), ("layer4", s32, 512, 512, 3, 1), ("dec0.c1", s16, 768, 256, 3, 1),
    ("dec0.c2", s16, 256, 256, 3, 1), ("dec1.c1", s8, 384, 128, 3, 1), ("dec1.c2", s8, 128, 128, 3, 1),
    ("dec2.c1", s4, 192, 64, 3, 1), ("dec2.c2", s4, 64, 64, 3, 1), ("dec3.c1", s2, 128, 32, 3, 1),
    ("dec3.c2", s2, 32, 32, 3, 1), ("dec4.c1", S, 32, 16, 3, 1), ("dec4.c2", S, 16, 16, 3, 1),
    ("head", S, 16, 24, 3, 1),
]


@pytest.fixture(autouse=True)
def _true_fp32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b
    torch.cuda.empty_cache()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).bfloat16()


@pytest.mark.parametrize("fam", FAMILIES, ids=[f[0] for f in FAMILIES])
def test_conv_family_at_benchmark_shape(fam):
    from uda_aerial_semantic_segmentation_research_b200 import ops
    name, H, Cin, Cout, k, s = fam
    p = (k - 1) // 2
    assert ops.tc_supported(0, B, H, H, Cin, Cout, k, k, s, p)
    x = _rand((B, H, H, Cin), 1)
    w = _rand((Cout, k, k, Cin), 2, (k * k * Cin) ** -0.5)
    # forward with the fused BatchNorm statistics (what the training step launches)
    yr = R.conv_fwd(x.float(), w.float(), None, s, p)
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device=DEV)
    y = ops.conv_fwd(x, w, None, s, p, bn_sums=sums)
    assert rel_err(y.float(), yr) < 1e-2, name
    yf = y.double().reshape(-1, Cout)
    assert rel_err(sums[Cout:], (yf * yf).sum(0)) < 1e-4
    assert float((sums[:Cout] - yf.sum(0)).abs().max()) < 1e-3 * float(yf.abs().sum(0).max())
    if name == "head":      # fp32 NCHW edge output with bias
        bias = torch.randn(Cout, device=DEV)
        yn = ops.conv_fwd(x, w, bias, s, p, nchw_out=True)
        assert rel_err(yn, R.conv_fwd(x.float(), w.float(), bias, s, p, nchw_out=True)) < 2e-3
    del yf
    # dgrad accumulating into an existing gradient (residual / skip consumers)
    dy = _rand(tuple(yr.shape), 3)
    del yr, y
    dxr = R.conv_dgrad(dy.float(), w.float(), x.shape, s, p)
    add = _rand(tuple(x.shape), 4)
    acc = add.clone()
    out = ops.conv_dgrad(dy, w, x.shape, s, p, addend=acc)
    assert out.data_ptr() == acc.data_ptr()
    assert rel_err(out.float(), dxr + add.float()) < 1e-2, name
    dx = ops.conv_dgrad(dy, w, x.shape, s, p)
    assert rel_err(dx.float(), dxr) < 1e-2, name
    del dxr, add, acc, out, dx
    # wgrad (fp32 accumulation over B*Ho*Wo pixels, atomics across pixel splits)
    dwr = R.conv_wgrad(dy.float(), x.float(), torch.zeros(Cout, k, k, Cin, device=DEV), s, p)
    dw = torch.zeros((Cout, k, k, Cin), device=DEV)
    ops.conv_wgrad(dy, x, dw, s, p)
    assert rel_err(dw, dwr) < 2e-3, name


def test_stem_at_benchmark_shape():
    """U-Net stem 7x7 stride 2 (Cin = 3) on the tensor cores, B = 16 @ 512 x 512."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    if not ops.stem_supported(B, S, S, 3, 64, 7, 2, 3):
        pytest.skip("stem path disabled")
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(B, 3, S, S, generator=g, device=DEV).bfloat16().float()
    w = _rand((64, 7, 7, 3), 12, 0.1)
    xs = ops.stem_pack_input(x, 3)
    sums = torch.zeros(128, dtype=torch.float64, device=DEV)
    y = ops.stem_fwd(xs, ops.stem_pack_weight(w), None, S, S, 7, 3, bn_sums=sums)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    yr = R.conv_fwd(x_nhwc, w.float(), None, 2, 3)
    assert rel_err(y.float(), yr) < 1e-2
    yf = y.double().reshape(-1, 64)
    assert rel_err(sums[64:], (yf * yf).sum(0)) < 1e-4
    dy = _rand(tuple(yr.shape), 13)
    dw = torch.zeros((64, 7, 7, 3), device=DEV)
    ops.stem_wgrad(dy, xs, dw, S, S, 7, 3)
    dwr = R.conv_wgrad(dy.float(), x_nhwc, torch.zeros(64, 7, 7, 3, device=DEV), 2, 3)
    assert rel_err(dw, dwr) < 2e-3


DEC_BLOCKS = [  # name, low-res H(=W) of x, C1 (upsampled channels), C2 (skip channels), Cout
    ("dec0.c1", s32, 512, 256, 256), ("dec1.c1", s16, 256, 128, 128), ("dec2.c1", s8, 128, 64, 64),
    ("dec3.c1", s4, 64, 64, 32), ("dec4.c1", s2, 32, 0, 16),
]


@pytest.mark.parametrize("blk", DEC_BLOCKS, ids=[b[0] for b in DEC_BLOCKS])
def test_decoder_conv1_without_materialised_concat(blk):
    """conv3x3(cat(upsample2x(x), skip), W) as conv_transpose4x4_s2(x, W4) + conv3x3(skip, Ws) at the benchmarked shape:
    forward (with the BatchNorm statistics of the sum), dx at low resolution, dskip and the merged weight gradient against
    the fp32 oracle ops (nearest upsample + concat + conv / torch.nn.grad on the device) on the same bf16 inputs."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    name, h, C1, C2, O = blk
    H = 2 * h
    x = _rand((B, h, h, C1), 21)
    skip = _rand((B, H, H, C2), 22) if C2 else None
    w = _rand((O, 3, 3, C1 + C2), 23, (9 * (C1 + C2)) ** -0.5)
    cat = R.upcat_fwd(x.float(), skip.float() if C2 else None)
    zr = R.conv_fwd(cat, w.float(), None, 1, 1)
    wx, ws, w4, ws_ft = ops.upconv_split_weights(w, C1, backward=True)
    assert torch.equal(w4, ops.weight_flip_transpose(wx)) and (not C2 or torch.equal(ws_ft, ops.weight_flip_transpose(ws)))
    sums = torch.zeros(2 * O, dtype=torch.float64, device=DEV)
    if C2:
        z = ops.conv_fwd_add(skip, ws, ops.upconv_fwd(x, wx), bn_sums=sums)
    else:
        z = ops.upconv_fwd(x, wx, bn_sums=sums)
    assert rel_err(z.float(), zr) < 1e-2, name
    zf = z.double().reshape(-1, O)
    assert rel_err(sums[O:], (zf * zf).sum(0)) < 1e-4
    assert float((sums[:O] - zf.sum(0)).abs().max()) < 1e-3 * float(zf.abs().sum(0).max())
    del zf, z
    dz = _rand(tuple(zr.shape), 24)
    del zr
    dcat = R.conv_dgrad(dz.float(), w.float(), cat.shape, 1, 1)
    dxr, dskipr = R.upcat_bwd(dcat, C1, C2)
    del dcat
    dx = ops.conv_fwd(dz, w4, None, 2, 1)
    assert rel_err(dx.float(), dxr) < 1e-2, name
    if C2:
        dskip = ops.conv_dgrad(dz, ws, skip.shape, 1, 1, w_ft=ws_ft)
        assert rel_err(dskip.float(), dskipr) < 1e-2, name
        del dskip
    del dx, dxr, dskipr
    dwr = R.conv_wgrad(dz.float(), cat, torch.zeros(O, 3, 3, C1 + C2, device=DEV), 1, 1)
    del cat
    dw4 = torch.zeros((C1, 4, 4, O), device=DEV)
    ops.conv_wgrad(x, dz, dw4, 2, 1)
    dws = None
    if C2:
        dws = torch.zeros((O, 3, 3, C2), device=DEV)
        ops.conv_wgrad(dz, skip, dws, 1, 1)
    dw = torch.zeros((O, 3, 3, C1 + C2), device=DEV)
    ops.upconv_merge_wgrad(dw4, dws, dw, C1)
    assert rel_err(dw, dwr) < 2e-3, name


@pytest.mark.parametrize("shape", [(16, 256, 32, 16, True), (16, 128, 64, 32, False), (2, 4, 32, 8, True),
                                   (3, 6, 96, 24, True)],
                         ids=["dec4", "dec3", "tiny", "c96_o24"])
def test_transposed_conv_halo_kernel_matches_persistent_path_and_oracle(shape, monkeypatch):
    """``uda_upconv_tc_fwd`` on wide images with <= 32 output channels runs ``conv_tc_uphalo_kernel`` (x halo loaded once,
    all four output parities in one accumulator set, N-packed instructions): same result as the four-tap-class launch of
    the persistent kernel (UDA_B200_UPHALO=0) up to accumulation order, and both against the fp32 oracle
    conv3x3(upsample2x(x)) on the device; BatchNorm statistics of the bf16-rounded output."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    Bn, h, C1, O, with_sums = shape
    w_lo = 128 if h < 64 else h          # the kernel needs w % 128 == 0
    x = _rand((Bn, h, w_lo, C1), 31)
    w = _rand((O, 3, 3, C1), 32, (9 * C1) ** -0.5)
    wx, _ = ops.upconv_split_weights(w, C1)
    yr = R.conv_fwd(R.upcat_fwd(x.float(), None), w.float(), None, 1, 1)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("UDA_B200_UPHALO", mode)
        sums = torch.zeros(2 * O, dtype=torch.float64, device=DEV) if with_sums else None
        y = ops.upconv_fwd(x, wx, bn_sums=sums)
        torch.cuda.synchronize()
        assert rel_err(y.float(), yr) < 1e-2, (shape, mode)
        if with_sums:
            yf = y.double().reshape(-1, O)
            assert rel_err(sums[O:], (yf * yf).sum(0)) < 1e-4, (shape, mode)
            assert float((sums[:O] - yf.sum(0)).abs().max()) < 1e-3 * float(yf.abs().sum(0).max()), (shape, mode)
        res[mode] = y
    assert rel_err(res["1"].float(), res["0"].float()) < 1e-2
    assert float((res["1"].float() - res["0"].float()).abs().mean()) < 2e-3 * float(res["0"].float().abs().mean())


@pytest.mark.parametrize("shape", [(16, 512, 32), (2, 256, 16), (3, 8, 24)], ids=["dec4_dx", "b2_c16", "r2_c24"])
def test_stride2_4x4_halo_kernel_matches_persistent_path_and_oracle(shape, monkeypatch):
    """``uda_conv2d_tc_fwd`` for a 4x4 stride-2 pad-1 convolution of a wide 16-channel tensor (dx of the transposed half of
    decoder conv1, block 4) runs ``conv_tc_downhalo_kernel`` (space-to-depth halo loaded once, sixteen shifted
    descriptors): same result as the persistent kernel's sixteen TMA boxes (UDA_B200_DOWNHALO=0) up to accumulation
    order, and both against F.conv2d in fp32 on the device."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    Bn, H, Co = shape
    W = max(H, 256)                      # the kernel needs (W/2) % 128 == 0
    x = _rand((Bn, H, W, 16), 41)
    w = _rand((Co, 4, 4, 16), 42, (16 * 16) ** -0.5)
    yr = R.conv_fwd(x.float(), w.float(), None, 2, 1)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("UDA_B200_DOWNHALO", mode)
        y = ops.conv_fwd(x, w, None, 2, 1)
        torch.cuda.synchronize()
        assert rel_err(y.float(), yr) < 1e-2, (shape, mode)
        res[mode] = y
    assert float((res["1"].float() - res["0"].float()).abs().mean()) < 2e-3 * float(res["0"].float().abs().mean())


@pytest.mark.parametrize("shape", [(16, 512, 32), (2, 256, 16), (3, 8, 24)], ids=["dec4_dw4", "b2_c16", "r2_c24"])
def test_stride2_4x4_halo_wgrad_matches_generic_path_and_oracle(shape, monkeypatch):
    """Weight gradient of the 4x4 stride-2 pad-1 convolution of a wide 16-channel tensor (dW4 of the transposed half of
    decoder conv1, block 4): ``conv_tc_wgrad_downhalo_kernel`` vs the generic tap-by-tap kernel (UDA_B200_DOWNHALO=0) and
    vs torch.nn.grad.conv2d_weight in fp32 on the device (2e-3, as every wgrad test)."""
    from uda_aerial_semantic_segmentation_research_b200 import ops
    Bn, H, Co = shape
    W = max(H, 256)
    x = _rand((Bn, H, W, 16), 51)
    dy = _rand((Bn, H // 2, W // 2, Co), 52)
    dwr = R.conv_wgrad(dy.float(), x.float(), torch.zeros(Co, 4, 4, 16, device=DEV), 2, 1)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("UDA_B200_DOWNHALO", mode)
        dw = torch.full((Co, 4, 4, 16), 0.25, device=DEV)        # the kernels ACCUMULATE
        ops.conv_wgrad(dy, x, dw, 2, 1)
        torch.cuda.synchronize()
        assert rel_err(dw - 0.25, dwr) < 2e-3, (shape, mode)
        res[mode] = dw
    assert rel_err(res["1"], res["0"]) < 1e-3
