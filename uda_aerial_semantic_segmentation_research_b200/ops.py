"""Thin tensor-level wrappers over the C ABI (``include/uda_b200.h``).

Every function validates device / dtype / contiguity, allocates outputs through the torch caching
allocator, and launches on the current torch CUDA stream.  No function here computes anything in
PyTorch: if the CUDA library is missing the call raises (``_lib.UdaError``).

Activations are NHWC tensors ``[B,H,W,C]`` (fp32 or bf16); conv weights are OHWI tensors
``[Cout,KH,KW,Cin]`` (the physical layout of a channels_last parameter).
"""
import os
import torch

from . import _lib
from ._lib import call, ptr, ll, ci, F32, BF16, I64, U8

I64_T = torch.int64
_WS = {}
#: set to "0" to force the FP32-pipe direct convolution everywhere (debug / A-B comparison)
USE_TC = os.environ.get("UDA_B200_USE_TC", "1") != "0"
#: BatchNorm batch statistics from the tensor-core conv epilogue (set "0" to use the separate bn_stats pass)
FUSE_BN_STATS = os.environ.get("UDA_B200_FUSE_BN_STATS", "1") != "0"
#: BatchNorm-backward reduction in the epilogue of the dgrad that produces dL/da.  OFF by default: measured on B200
#: (B=16, 512x512) the extra epilogue work (reading `a`/`z`, 62 shuffles per 32x32 chunk, not overlapped when a CTA
#: owns one or two tiles) costs +1.2 ms of dgrad time per step and saves only 0.8 ms of bn_bwd_reduce launches.
FUSE_BN_BWD = os.environ.get("UDA_B200_FUSE_BN_BWD", "0") == "1"
#: the persistent / halo tensor-core kernels are the ones with the fused epilogues (UDA_B200_TC_PERSIST=0 disables)
TC_PERSIST = os.environ.get("UDA_B200_TC_PERSIST", "1") != "0"
#: eval mode: fold BatchNorm into the convolution weights and run conv + BN (+ residual) + activation as one launch
#: (set "0" to run the separate normalise pass, e.g. to A/B the two inference paths)
FOLD_BN_EVAL = os.environ.get("UDA_B200_FOLD_BN_EVAL", "1") != "0"
#: backward: launch the weight-gradient kernels on a second stream (they are off the dgrad -> BatchNorm-backward critical
#: path, so their CTAs fill the SMs that the tails / prologues of the chain leave idle); joined at the end of backward
#: (measured at B=16, 512x512: 8.99 -> 8.73 ms per supervised step; UDA_B200_WGRAD_STREAM=0 keeps everything on one stream)
WGRAD_STREAM = os.environ.get("UDA_B200_WGRAD_STREAM", "1") != "0"
#: refresh the flipped / transposed dgrad weight copies on the side stream during the forward
PREFETCH_WFT = os.environ.get("UDA_B200_PREFETCH_WFT", "1") != "0"
#: training: conv + BatchNorm + activation (+ residual) as ONE launch for the layers whose output tiles fit the tensor
#: memory of one wave of CTAs (grid barrier between the statistics and the normalise pass; uda_conv2d_tc_fwd_bn_act):
#: layer2-4 and decoder blocks 0-1 at B=16, 512x512 (33 of the 46 BatchNorm layers).  Measured 8.37-8.43 vs 8.53 ms per
#: step; UDA_B200_FUSE_BN_APPLY=0 runs the separate normalise pass everywhere
FUSE_BN_APPLY = os.environ.get("UDA_B200_FUSE_BN_APPLY", "1") != "0"
#: training: the stem's normalise + activation pass also produces the 3x3 stride-2 max-pool of its output (one pass over
#: the largest tensor of the encoder instead of two; uda_bn_apply_maxpool_fused).  UDA_B200_FUSE_BN_POOL=0 = two launches
FUSE_BN_POOL = os.environ.get("UDA_B200_FUSE_BN_POOL", "1") != "0"
#: decoder conv1 as conv_transpose4x4(x) + conv3x3(skip): the upsampled / concatenated tensor is never materialised
#: (UDA_B200_FUSE_UPCAT=0 runs the upsample+concat copy kernel and one 3x3 convolution over the concatenation)
FUSE_UPCAT = os.environ.get("UDA_B200_FUSE_UPCAT", "1") != "0"
#: counts launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = 0
#: algorithmic FLOPs (2*M*N*K) and call count of the convolutions routed to the tcgen05 kernels
TC_FLOPS = 0
TC_CALLS = 0


def _tc_account(B, Ho, Wo, Cout, Cin, KH, KW):
    global TC_FLOPS, TC_CALLS
    TC_FLOPS += 2 * B * Ho * Wo * Cout * Cin * KH * KW
    TC_CALLS += 1


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {t.dtype} (float32 or bfloat16)")


def _chk(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.UdaError(f"{name}: expected a CUDA tensor (the uda_b200 hot path has no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.UdaError(f"{name}: tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


_WS_RETIRED = []


def workspace(nbytes, device):
    """Stream-ordered scratch (re-used by consecutive launches on the same stream).  One block per (device, stream):
    launches on different streams never share scratch.  A block that has to grow is RETIRED, not freed — a captured
    CUDA graph may have baked its address in, and handing the memory back to the caching allocator would let graph
    replays scribble over somebody else's tensor."""
    dev = device.index if device.index is not None else torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _WS_RETIRED.append(ws)
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


_BN_WS = {}


def bn_workspace(device, C):
    """Once-zeroed scratch of the BatchNorm reductions (the kernels leave it zeroed: no per-call memset)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _BN_WS.get(key)
    need = 16 + 2 * 4096 * 8 + 3 * 4096 * 4   # fixed layout, see uda_bn_stats
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 17), dtype=torch.uint8, device=device)
        _BN_WS[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc(x, dtype, cpad=None):
    _chk(x, "nchw_to_nhwc.x", torch.float32)
    B, C, H, W = x.shape
    cpad = cpad or C
    out = torch.empty((B, H, W, cpad), dtype=dtype, device=x.device)
    call("nchw_f32_to_nhwc", ptr(x), ptr(out), ci(dt(out)), ci(B), ci(C), ci(cpad), ll(H * W), _stream())
    _count()
    return out


def nhwc_to_nchw(x, C=None):
    _chk(x, "nhwc_to_nchw.x")
    B, H, W, cpad = x.shape
    C = C or cpad
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    call("nhwc_to_nchw_f32", ptr(x), ci(dt(x)), ptr(out), ci(B), ci(C), ci(cpad), ll(H * W), _stream())
    _count()
    return out


def cast_f32(src, dst):
    _chk(src, "cast.src", torch.float32)
    _chk(dst, "cast.dst")
    assert src.numel() == dst.numel()
    call("cast_f32", ptr(src), ptr(dst), ci(dt(dst)), ll(src.numel()), _stream())
    _count()
    return dst


# ------------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------------
def _geom(x_shape, w_shape, stride, pad):
    B, H, W, Cin = x_shape
    Cout, KH, KW, Cin_w = w_shape
    if Cin != Cin_w:
        raise _lib.UdaError(f"conv: input has {Cin} channels, weight expects {Cin_w}")
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    return B, H, W, Cin, Cout, KH, KW, Ho, Wo


def tc_supported(op, B, H, W, Cin, Cout, KH, KW, stride, pad):
    if not USE_TC:
        return False
    return bool(_lib.lib().uda_conv2d_tc_supported(ci(op), ci(B), ci(H), ci(W), ci(Cin), ci(Cout), ci(KH), ci(KW),
                                                    ci(stride), ci(pad)))


def conv_fwd(x, w, bias=None, stride=1, pad=1, nchw_out=False, bn_sums=None, force_direct=False):
    """y = conv2d(x, w) (+bias).  Returns NHWC ``y`` (dtype of x) or, with ``nchw_out``, fp32 NCHW."""
    _chk(x, "conv_fwd.x"); _chk(w, "conv_fwd.w")
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x.shape, w.shape, stride, pad)
    y_nhwc = None if nchw_out else torch.empty((B, Ho, Wo, Cout), dtype=x.dtype, device=x.device)
    y_nchw = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device) if nchw_out else None
    if bias is not None:
        _chk(bias, "conv_fwd.bias", torch.float32)
    use_tc = (not force_direct and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
              and tc_supported(0, B, H, W, Cin, Cout, KH, KW, stride, pad))
    if use_tc:
        call("conv2d_tc_fwd", ptr(x), ptr(w), ptr(bias), ptr(y_nhwc), ptr(y_nchw), ptr(bn_sums), ci(B), ci(H), ci(W),
             ci(Cin), ci(Cout), ci(KH), ci(KW), ci(stride), ci(pad), _stream())
        _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    else:
        if bn_sums is not None:
            raise _lib.UdaError("conv_fwd: fused BN statistics need the tensor-core path")
        call("conv2d_direct_fwd", ptr(x), ci(dt(x)), ptr(w), ci(dt(w)), ptr(bias), ptr(y_nhwc), ptr(y_nchw), ci(B),
             ci(H), ci(W), ci(Cin), ci(Cout), ci(KH), ci(KW), ci(stride), ci(pad), _stream())
    _count()
    return y_nchw if nchw_out else y_nhwc


def conv_fwd_fused(x, w, bias, act_slope, addend=None, nchw_out=False, stride=1, pad=1):
    """Inference form: y = act(conv(x, w) + bias (+ addend)) in one launch (``act_slope`` 0 = ReLU, 0.2 = LeakyReLU,
    1 = none).  With BatchNorm folded into ``w`` / ``bias`` (``bn_fold_conv``) this is an eval-mode conv + BN (+ residual)
    + activation.  Tensor-core path only (bf16); the caller checks ``tc_supported``."""
    _chk(x, "conv_fwd_fused.x", torch.bfloat16); _chk(w, "conv_fwd_fused.w", torch.bfloat16)
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x.shape, w.shape, stride, pad)
    if bias is not None:
        _chk(bias, "conv_fwd_fused.bias", torch.float32)
    if addend is not None:
        _chk(addend, "conv_fwd_fused.addend", torch.bfloat16)
        if tuple(addend.shape) != (B, Ho, Wo, Cout):
            raise _lib.UdaError("conv_fwd_fused: addend must have the output's shape")
    y_nhwc = None if nchw_out else torch.empty((B, Ho, Wo, Cout), dtype=x.dtype, device=x.device)
    y_nchw = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device) if nchw_out else None
    call("conv2d_tc_fwd_fused", ptr(x), ptr(w), ptr(bias), ptr(addend), float(act_slope), ptr(y_nhwc), ptr(y_nchw), ci(B),
         ci(H), ci(W), ci(Cin), ci(Cout), ci(KH), ci(KW), ci(stride), ci(pad), _stream())
    _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    _count()
    return y_nchw if nchw_out else y_nhwc


def bn_fold_conv(w_f32, conv_bias, gamma, beta, running_mean, running_var, eps=1e-5):
    """Eval-mode BatchNorm folded into a convolution: returns (bf16 OHWI weights scaled per output channel, fp32 bias)."""
    _chk(w_f32, "bn_fold_conv.w", torch.float32)
    O = w_f32.shape[0]
    w_out = torch.empty(w_f32.shape, dtype=torch.bfloat16, device=w_f32.device)
    b_out = torch.empty(O, dtype=torch.float32, device=w_f32.device)
    call("bn_fold_conv", ptr(w_f32), ptr(conv_bias), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
         float(eps), ptr(w_out), ptr(b_out), ci(O), ci(w_f32.numel() // O), _stream())
    _count()
    return w_out, b_out


# ---- decoder conv1 without the upsampled / concatenated tensor (see uda_upconv_* in include/uda_b200.h) ----------
def upconv_split_weights(w, C1, backward=False):
    """OHWI bf16 [O,3,3,C1+C2] -> (wx_ft bf16 [O,4,4,C1] tap groups summed, ws bf16 [O,3,3,C2] or None); with
    ``backward`` also (w4 bf16 [C1,4,4,O] = flip-transpose(wx_ft), ws_ft bf16 [C2,3,3,O] or None) — one launch."""
    _chk(w, "upconv_split_weights.w", torch.bfloat16)
    O, KH, KW, C = w.shape
    if KH != 3 or KW != 3 or C1 <= 0 or C1 > C:
        raise _lib.UdaError("upconv_split_weights: expects a 3x3 weight and 0 < C1 <= Cin")
    C2 = C - C1
    wx = torch.empty((O, 4, 4, C1), dtype=torch.bfloat16, device=w.device)
    ws = torch.empty((O, 3, 3, C2), dtype=torch.bfloat16, device=w.device) if C2 else None
    w4 = torch.empty((C1, 4, 4, O), dtype=torch.bfloat16, device=w.device) if backward else None
    wsf = torch.empty((C2, 3, 3, O), dtype=torch.bfloat16, device=w.device) if (backward and C2) else None
    call("upconv_split_weights", ptr(w), ptr(wx), ptr(ws), ptr(w4), ptr(wsf), ci(O), ci(C1), ci(C2), _stream())
    _count()
    return (wx, ws, w4, wsf) if backward else (wx, ws)


def upconv_fwd(x, wx_ft, bias=None, addend=None, act_slope=1.0, bn_sums=None):
    """y[B,2H,2W,O] = act(conv_transpose4x4_s2_p1(x, wx_ft) + bias (+ addend)) == conv3x3(upsample2x(x), Wx) (...)."""
    _chk(x, "upconv_fwd.x", torch.bfloat16); _chk(wx_ft, "upconv_fwd.wx_ft", torch.bfloat16)
    B, h, w_, C1 = x.shape
    O = wx_ft.shape[0]
    y = torch.empty((B, 2 * h, 2 * w_, O), dtype=torch.bfloat16, device=x.device)
    if addend is not None:
        _chk(addend, "upconv_fwd.addend", torch.bfloat16)
    call("upconv_tc_fwd", ptr(x), ptr(wx_ft), ptr(bias), ptr(addend), float(act_slope), ptr(y), ptr(bn_sums), ci(B),
         ci(2 * h), ci(2 * w_), ci(C1), ci(O), _stream())
    _tc_account(B, h, w_, O, C1, 4, 4)
    _count()
    return y


def conv_fwd_add(x, w, addend, bn_sums=None, stride=1, pad=1):
    """y = conv(x, w) + addend, BatchNorm statistics of the sum in the epilogue (tensor-core path)."""
    _chk(x, "conv_fwd_add.x", torch.bfloat16); _chk(w, "conv_fwd_add.w", torch.bfloat16)
    _chk(addend, "conv_fwd_add.addend", torch.bfloat16)
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x.shape, w.shape, stride, pad)
    if tuple(addend.shape) != (B, Ho, Wo, Cout):
        raise _lib.UdaError("conv_fwd_add: addend must have the output's shape")
    y = torch.empty((B, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    call("conv2d_tc_fwd_add", ptr(x), ptr(w), ptr(addend), ptr(y), ptr(bn_sums), ci(B), ci(H), ci(W), ci(Cin), ci(Cout),
         ci(KH), ci(KW), ci(stride), ci(pad), _stream())
    _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    _count()
    return y


_FUSED_BN_SHAPES = {}   # (B,H,W,Cin,Cout,k,stride,pad) -> did the fused conv + BatchNorm launch accept the shape?


def conv_bn_act_fused(x, w, slot, gamma, beta, running_mean, running_var, eps, momentum, slope, stride=1, pad=1,
                      addend=None, residual=None):
    """Training-mode a = act(BN(conv(x, w) (+ addend)) (+ residual)) as ONE launch (``uda_conv2d_tc_fwd_bn_act``).
    ``slot``: zeroed float64[2*Cout + 1] (statistics + the grid barrier's arrival counter).  Returns
    (z, a, mean, rstd, scale, shift), or None when the layer's output tiles do not fit one wave's tensor memory
    (nothing was launched: the caller runs the two-launch form)."""
    _chk(x, "conv_bn_act_fused.x", torch.bfloat16); _chk(w, "conv_bn_act_fused.w", torch.bfloat16)
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x.shape, w.shape, stride, pad)
    key = (B, H, W, Cin, Cout, KH, stride, pad)
    if _FUSED_BN_SHAPES.get(key) is False:
        return None
    for t, nm in ((addend, "addend"), (residual, "residual")):
        if t is not None:
            _chk(t, "conv_bn_act_fused." + nm, torch.bfloat16)
            if tuple(t.shape) != (B, Ho, Wo, Cout):
                raise _lib.UdaError(f"conv_bn_act_fused: {nm} must have the output's shape")
    if slot.dtype != torch.float64 or slot.numel() < 2 * Cout + 1 or not slot.is_contiguous():
        raise _lib.UdaError("conv_bn_act_fused: slot must be a contiguous float64[2*Cout+1]")
    z = torch.empty((B, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    a = torch.empty_like(z)
    st = torch.empty((4, Cout), dtype=torch.float32, device=x.device)
    ok = call("conv2d_tc_fwd_bn_act", ptr(x), ptr(w), ptr(addend), ptr(residual), ptr(z), ptr(a), ptr(slot),
              ptr(slot[2 * Cout:]), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(st[0]), ptr(st[1]),
              ptr(st[2]), ptr(st[3]), ci(B), ci(H), ci(W), ci(Cin), ci(Cout), ci(KH), ci(KW), ci(stride), ci(pad),
              float(eps), float(momentum), float(slope), _stream(), unsupported_ok=True)
    _FUSED_BN_SHAPES[key] = ok
    if not ok:
        return None
    _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    _count()
    return z, a, st[0], st[1], st[2], st[3]


def upconv_merge_wgrad(dw4, dws, dw, C1):
    """dw (fp32 OHWI [O,3,3,C1+C2]) += un-grouped dw4 (fp32 [C1,4,4,O]) on the x channels, dws on the skip channels."""
    _chk(dw4, "upconv_merge_wgrad.dw4", torch.float32); _chk(dw, "upconv_merge_wgrad.dw", torch.float32)
    O, _, _, C = dw.shape
    call("upconv_merge_wgrad", ptr(dw4), ptr(dws), ptr(dw), ci(O), ci(C1), ci(C - C1), _stream())
    _count()
    return dw


def weight_flip_transpose(w, out=None):
    """OHWI bf16 weights -> [Cin][KH][KW][Cout] flipped copy (the forward-conv weights of dgrad)."""
    _chk(w, "weight_flip_transpose.w", torch.bfloat16)
    O, KH, KW, I = w.shape
    if out is None:
        out = torch.empty((I, KH, KW, O), dtype=torch.bfloat16, device=w.device)
    call("conv2d_weight_flip_transpose", ptr(w), ptr(out), ci(O), ci(I), ci(KH), ci(KW), _stream())
    _count()
    return out


# ---- Cin = 3 stems on the tensor cores ----------------------------------------------------------
def stem_supported(B, H, W, Cin, Cout, K, stride, pad):
    if not USE_TC:
        return False
    return bool(_lib.lib().uda_stem_tc_supported(ci(B), ci(H), ci(W), ci(Cin), ci(Cout), ci(K), ci(stride), ci(pad)))


def stem_pack_input(x, pad):
    """fp32 NCHW image -> zero-padded 4-channel bf16 buffer [B, H+8, W+8, 4]."""
    _chk(x, "stem_pack_input.x", torch.float32)
    B, C, H, W = x.shape
    if C != 3:
        raise _lib.UdaError("stem_pack_input: expected 3 input channels")
    xs = torch.empty((B, H + 8, W + 8, 4), dtype=torch.bfloat16, device=x.device)
    call("stem_pack_input", ptr(x), ptr(xs), ci(B), ci(H), ci(W), ci(pad), _stream())
    _count()
    return xs


def stem_pack_weight(w):
    """OHWI bf16 [Cout,K,K,3] -> [Cout, K, KS*4] with KS = 8 (K=7) or 4 (K=4)."""
    _chk(w, "stem_pack_weight.w", torch.bfloat16)
    O, K, _, I = w.shape
    KS = 8 if K == 7 else 4
    ws = torch.empty((O, K, KS * 4), dtype=torch.bfloat16, device=w.device)
    call("stem_pack_weight", ptr(w), ptr(ws), ci(O), ci(K), _stream())
    _count()
    return ws


def stem_fwd(xs, ws, bias, H, W, K, pad, bn_sums=None, act_slope=None):
    """``act_slope`` (eval mode, BatchNorm folded into ws / bias): activation in the epilogue."""
    B = xs.shape[0]
    O = ws.shape[0]
    y = torch.empty((B, H // 2, W // 2, O), dtype=torch.bfloat16, device=xs.device)
    if act_slope is not None:
        call("stem_tc_fwd_act", ptr(xs), ptr(ws), ptr(bias), ptr(y), float(act_slope), ci(B), ci(H), ci(W), ci(O), ci(K),
             ci(pad), _stream())
    else:
        call("stem_tc_fwd", ptr(xs), ptr(ws), ptr(bias), ptr(y), ptr(bn_sums), ci(B), ci(H), ci(W), ci(O), ci(K), ci(pad),
             _stream())
    _tc_account(B, H // 2, W // 2, O, 3, K, K)
    _count()
    return y


def stem_wgrad(dy, xs, dw, H, W, K, pad):
    """dw (fp32 OHWI [Cout,K,K,3]) += wgrad."""
    _chk(dy, "stem_wgrad.dy", torch.bfloat16); _chk(dw, "stem_wgrad.dw", torch.float32)
    B, O = xs.shape[0], dw.shape[0]
    KS = 8 if K == 7 else 4
    scratch = workspace(O * K * KS * 4 * 4, dy.device)
    call("stem_tc_wgrad", ptr(dy), ptr(xs), ptr(dw), ptr(scratch), ci(B), ci(H), ci(W), ci(O), ci(K), ci(pad), _stream())
    _tc_account(B, H // 2, W // 2, O, 3, K, K)
    _count(3)
    return dw


def weight_flip_transpose_batch(base, out, table):
    """Flip/transpose every conv weight of a flat bf16 buffer in one launch (table: int32 [n,5] on device)."""
    call("conv2d_weight_flip_transpose_batch", ptr(base), ptr(out), ptr(table), ci(table.shape[0]), _stream())
    _count()
    return out


def dgrad_bnstats_supported(x_shape, w_shape, stride, pad, dtype):
    """Can the tensor-core dgrad of this convolution also emit the BatchNorm-backward statistics of its output?"""
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x_shape, w_shape, stride, pad)
    if not (USE_TC and TC_PERSIST and FUSE_BN_BWD and dtype == torch.bfloat16) or (stride == 2 and KH < 2) or Cin <= 32:
        return False
    return bool(tc_supported(1, B, H, W, Cin, Cout, KH, KW, stride, pad)
                and _lib.lib().uda_bn_bwd_fused_supported(ci(BF16), ll(B * H * W), ci(Cin)))


def conv_dgrad(dy, w, x_shape, stride=1, pad=1, addend=None, force_direct=False, w_ft=None, bn_stats=None):
    """dx = conv_transpose(dy, w) (+ addend, accumulated in place into ``addend``'s buffer when given).

    ``w_ft`` (optional): cached ``weight_flip_transpose(w)`` for the tensor-core path.
    ``bn_stats`` = (a, z_or_None, slope, sums): also accumulate the BatchNorm-backward sums of dx (see
    ``uda_conv2d_tc_dgrad_bnstats``); tensor-core path only."""
    _chk(dy, "conv_dgrad.dy"); _chk(w, "conv_dgrad.w")
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x_shape, w.shape, stride, pad)
    if tuple(dy.shape) != (B, Ho, Wo, Cout):
        raise _lib.UdaError(f"conv_dgrad: dy shape {tuple(dy.shape)} != {(B, Ho, Wo, Cout)}")
    if addend is not None:
        _chk(addend, "conv_dgrad.addend", dy.dtype)
        if tuple(addend.shape) != (B, H, W, Cin):
            raise _lib.UdaError("conv_dgrad: addend shape mismatch")
        dx = addend
    else:
        dx = torch.empty((B, H, W, Cin), dtype=dy.dtype, device=dy.device)
    use_tc = (not force_direct and dy.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
              and tc_supported(1, B, H, W, Cin, Cout, KH, KW, stride, pad))
    if use_tc:
        if w_ft is None:
            w_ft = weight_flip_transpose(w)
        if bn_stats is not None:
            a, z, slope, sums = bn_stats
            _chk(a, "conv_dgrad.bn_stats.a", dy.dtype)
            if tuple(a.shape) != (B, H, W, Cin) or sums.dtype != torch.float64 or sums.numel() != 2 * Cin:
                raise _lib.UdaError("conv_dgrad: bn_stats shape mismatch")
            call("conv2d_tc_dgrad_bnstats", ptr(dy), ptr(w_ft), ptr(addend), ptr(dx), ci(B), ci(H), ci(W), ci(Cin), ci(Cout),
                 ci(KH), ci(KW), ci(stride), ci(pad), ptr(a), ptr(z), float(slope), ptr(sums), _stream())
        else:
            call("conv2d_tc_dgrad", ptr(dy), ptr(w_ft), ptr(addend), ptr(dx), ci(B), ci(H), ci(W), ci(Cin), ci(Cout), ci(KH),
                 ci(KW), ci(stride), ci(pad), _stream())
        _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    elif bn_stats is not None:
        raise _lib.UdaError("conv_dgrad: bn_stats needs the tensor-core path")
    else:
        call("conv2d_direct_dgrad", ptr(dy), ci(dt(dy)), ptr(w), ci(dt(w)), ptr(addend), ptr(dx), ci(B), ci(H), ci(W), ci(Cin),
             ci(Cout), ci(KH), ci(KW), ci(stride), ci(pad), _stream())
    _count()
    return dx


def conv_wgrad(dy, x, dw, stride=1, pad=1, force_direct=False):
    """dw (fp32 OHWI) += dy^T * im2col(x)."""
    _chk(dy, "conv_wgrad.dy"); _chk(x, "conv_wgrad.x"); _chk(dw, "conv_wgrad.dw", torch.float32)
    B, H, W, Cin, Cout, KH, KW, Ho, Wo = _geom(x.shape, dw.shape, stride, pad)
    if tuple(dy.shape) != (B, Ho, Wo, Cout) or dy.dtype != x.dtype:
        raise _lib.UdaError("conv_wgrad: dy/x mismatch")
    use_tc = (not force_direct and x.dtype == torch.bfloat16
              and tc_supported(2, B, H, W, Cin, Cout, KH, KW, stride, pad))
    if use_tc:
        call("conv2d_tc_wgrad", ptr(dy), ptr(x), ptr(dw), ci(B), ci(H), ci(W), ci(Cin), ci(Cout), ci(KH), ci(KW),
             ci(stride), ci(pad), _stream())
        _tc_account(B, Ho, Wo, Cout, Cin, KH, KW)
    else:
        call("conv2d_direct_wgrad", ptr(dy), ptr(x), ci(dt(x)), ptr(dw), ci(B), ci(H), ci(W), ci(Cin), ci(Cout),
             ci(KH), ci(KW), ci(stride), ci(pad), _stream())
    _count()
    return dw


# ------------------------------------------------------------------------------------------------
# output-space adversarial path (softmax -> discriminator behind a gradient-reversal layer)
# ------------------------------------------------------------------------------------------------
def softmax_nhwc(logits, cpad):
    """fp32 NCHW logits [B,C,H,W] -> channels-last bf16 probabilities [B,H,W,cpad] (channels >= C are zero)."""
    _chk(logits, "softmax_nhwc.logits", torch.float32)
    B, C, H, W = logits.shape
    out = torch.empty((B, H, W, cpad), dtype=torch.bfloat16, device=logits.device)
    call("softmax_nchw_to_nhwc", ptr(logits), ptr(out), ci(B), ci(C), ci(cpad), ll(H * W), _stream())
    _count()
    return out


def softmax_bwd_grl(probs, dprobs, scale, C, out=None):
    """dlogits (fp32 NCHW) = scale * softmax_backward(probs, dprobs); accumulated into ``out`` when given.
    ``scale`` = -alpha folds the gradient-reversal layer (reference src/models/uda.py:103-112) into this pass."""
    _chk(probs, "softmax_bwd_grl.probs", torch.bfloat16); _chk(dprobs, "softmax_bwd_grl.dprobs", torch.bfloat16)
    B, H, W, cpad = probs.shape
    if tuple(dprobs.shape) != tuple(probs.shape):
        raise _lib.UdaError("softmax_bwd_grl: probs / dprobs shape mismatch")
    acc = out is not None
    if acc:
        _chk(out, "softmax_bwd_grl.out", torch.float32)
        if tuple(out.shape) != (B, C, H, W):
            raise _lib.UdaError("softmax_bwd_grl: out must be fp32 [B,C,H,W]")
    else:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=probs.device)
    call("softmax_bwd_grl", ptr(probs), ptr(dprobs), ptr(out), float(scale), ci(1 if acc else 0), ci(B), ci(C), ci(cpad),
         ll(H * W), _stream())
    _count()
    return out


def scale(x, s):
    """y = s * x in one pass (fp32 / bf16)."""
    _chk(x, "scale.x")
    y = torch.empty_like(x)
    call("scale", ptr(x), ptr(y), ci(dt(x)), float(s), ll(x.numel()), _stream())
    _count()
    return y


def pad_channels(w, cpad):
    """bf16 [..., c] -> [..., cpad] zero-padded copy (weights whose input channel count is not a tensor-core atom)."""
    _chk(w, "pad_channels.w", torch.bfloat16)
    c = w.shape[-1]
    out = torch.empty(tuple(w.shape[:-1]) + (cpad,), dtype=torch.bfloat16, device=w.device)
    call("pad_channels", ptr(w), ptr(out), ll(w.numel() // c), ci(c), ci(cpad), _stream())
    _count()
    return out


def unpad_channels_add(src, dst):
    """dst[..., c] (fp32) += src[..., :c]."""
    _chk(src, "unpad_channels_add.src", torch.float32); _chk(dst, "unpad_channels_add.dst", torch.float32)
    c, cpad = dst.shape[-1], src.shape[-1]
    call("unpad_channels_add", ptr(src), ptr(dst), ll(dst.numel() // c), ci(c), ci(cpad), _stream())
    _count()
    return dst


# ------------------------------------------------------------------------------------------------
# batch norm / activations / pooling / concat
# ------------------------------------------------------------------------------------------------
def bn_stats(x, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1):
    """Batch statistics of NHWC x -> (mean, rstd, scale, shift) fp32 [C]; updates running stats in place."""
    _chk(x, "bn_stats.x")
    C = x.shape[-1]
    M = x.numel() // C
    st = torch.empty((4, C), dtype=torch.float32, device=x.device)
    ws = bn_workspace(x.device, C)
    call("bn_stats", ptr(x), ci(dt(x)), ll(M), ci(C), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
         ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), float(eps), float(momentum), ptr(ws), _stream())
    _count(1)
    return st[0], st[1], st[2], st[3]


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps=1e-5):
    C = running_mean.numel()
    st = torch.empty((2, C), dtype=torch.float32, device=running_mean.device)
    call("bn_eval_coeffs", ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(st[0]), ptr(st[1]), ci(C),
         float(eps), _stream())
    _count()
    return st[0], st[1]


def bn_apply(x, scale, shift, residual=None, slope=0.0, out=None):
    _chk(x, "bn_apply.x")
    C = x.shape[-1]
    y = torch.empty_like(x) if out is None else out
    if residual is not None:
        _chk(residual, "bn_apply.residual", x.dtype)
    call("bn_apply", ptr(x), ptr(residual), ptr(y), ci(dt(x)), ptr(scale), ptr(shift), ll(x.numel() // C), ci(C),
         float(slope), _stream())
    _count()
    return y


def bn_apply_fused(x, sums, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1, residual=None, slope=0.0):
    """Normalise + activation with batch statistics taken from the conv epilogue (``sums`` double[2C]).
    Returns (y, mean, rstd); running statistics are updated in place."""
    _chk(x, "bn_apply_fused.x")
    C = x.shape[-1]
    M = x.numel() // C
    st = torch.empty((4, C), dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    call("bn_apply_fused", ptr(x), ptr(residual), ptr(y), ci(dt(x)), ptr(sums), ptr(gamma), ptr(beta),
         ptr(running_mean), ptr(running_var), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ll(M), ci(C), float(eps),
         float(momentum), float(slope), _stream())
    _count()
    return y, st[0], st[1], st[2], st[3]


def bn_apply_maxpool_fused(x, sums, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1, slope=0.0):
    """``bn_apply_fused`` + ``maxpool_fwd`` of its output as ONE pass over x (the ResNet stem tail).  Returns
    (a, mean, rstd, scale, shift, y, idx), or None when the launch declines (nothing ran)."""
    if not FUSE_BN_POOL or x.dtype != torch.bfloat16:
        return None
    _chk(x, "bn_apply_maxpool_fused.x")
    B, H, W, C = x.shape
    if H % 2 or W % 2:
        return None
    st = torch.empty((4, C), dtype=torch.float32, device=x.device)
    a = torch.empty_like(x)
    y = torch.empty((B, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    idx = torch.empty((B, H // 2, W // 2, C), dtype=torch.uint8, device=x.device)
    ok = call("bn_apply_maxpool_fused", ptr(x), ptr(a), ptr(y), ptr(idx), ci(dt(x)), ptr(sums), ptr(gamma), ptr(beta),
              ptr(running_mean), ptr(running_var), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ci(B), ci(H), ci(W),
              ci(C), float(eps), float(momentum), float(slope), _stream(), unsupported_ok=True)
    if ok is False:
        return None
    _count()
    return a, st[0], st[1], st[2], st[3], y, idx


def bn_bwd(dy, x, a, gamma, mean, rstd, slope, dgamma, dbeta, dres=None, dres_accumulate=False,
           param_accumulate=True, scale=None, shift=None):
    """BatchNorm(+activation) backward.  Returns dx; accumulates dgamma/dbeta; fills/accumulates dres.
    ``a=None`` with ``scale``/``shift``: the activation mask is recomputed from x*scale+shift."""
    _chk(dy, "bn_bwd.dy"); _chk(x, "bn_bwd.x", dy.dtype)
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    ws = bn_workspace(x.device, C)
    call("bn_bwd", ptr(dy), ptr(x), ptr(a), ci(dt(x)), ptr(gamma), ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
         ptr(dx), ptr(dres),
         ci(1 if dres_accumulate else 0), ptr(dgamma), ptr(dbeta), ci(1 if param_accumulate else 0), ll(M), ci(C),
         float(slope), ptr(ws), _stream())
    _count(2)
    return dx


def bn_bwd_apply_fused(dy, x, a, sums, v_is_z, gamma, beta, mean, rstd, slope, dgamma, dbeta, dres=None,
                       dres_accumulate=False, param_accumulate=True, scale=None, shift=None):
    """BatchNorm(+activation) backward whose per-channel sums were produced by the dgrad epilogue
    (``conv_dgrad(bn_stats=...)``): finalize + apply in one launch."""
    _chk(dy, "bn_bwd_apply_fused.dy"); _chk(x, "bn_bwd_apply_fused.x", dy.dtype)
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    call("bn_bwd_apply_fused", ptr(dy), ptr(x), ptr(a), ci(dt(x)), ptr(sums), ci(1 if v_is_z else 0), ptr(gamma), ptr(beta),
         ptr(mean), ptr(rstd), ptr(scale), ptr(shift), ptr(dx), ptr(dres), ci(1 if dres_accumulate else 0), ptr(dgamma),
         ptr(dbeta), ci(1 if param_accumulate else 0), ll(M), ci(C), float(slope), _stream())
    _count()
    return dx


def act_bwd(dy, a, slope):
    dx = torch.empty_like(dy)
    call("act_bwd", ptr(dy), ptr(a), ptr(dx), ci(dt(dy)), ll(dy.numel()), float(slope), _stream())
    _count()
    return dx


def bias_act(x, bias, slope, out=None):
    C = x.shape[-1]
    y = torch.empty_like(x) if out is None else out
    call("bias_act", ptr(x), ptr(bias), ptr(y), ci(dt(x)), ll(x.numel() // C), ci(C), float(slope), _stream())
    _count()
    return y


def colsum(x, out, scale=1.0, accumulate=True):
    C = x.shape[-1]
    ws = workspace(C * 8, x.device)
    call("colsum", ptr(x), ci(dt(x)), ptr(out), ll(x.numel() // C), ci(C), float(scale), ci(1 if accumulate else 0),
         ptr(ws), _stream())
    _count(2)
    return out


def maxpool_fwd(x):
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), dtype=x.dtype, device=x.device)
    idx = torch.empty((B, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    call("maxpool3x3s2_fwd", ptr(x), ptr(y), ptr(idx), ci(dt(x)), ci(B), ci(H), ci(W), ci(C), _stream())
    _count()
    return y, idx


def maxpool_bwd(dy, idx, x_shape, addend=None):
    B, H, W, C = x_shape
    dx = addend if addend is not None else torch.empty((B, H, W, C), dtype=dy.dtype, device=dy.device)
    call("maxpool3x3s2_bwd", ptr(dy), ptr(idx), ptr(addend), ptr(dx), ci(dt(dy)), ci(B), ci(H), ci(W), ci(C), _stream())
    _count()
    return dx


def upcat_fwd(x, skip=None):
    B, H2, W2, C1 = x.shape
    C2 = skip.shape[-1] if skip is not None else 0
    out = torch.empty((B, 2 * H2, 2 * W2, C1 + C2), dtype=x.dtype, device=x.device)
    call("upsample2x_concat_fwd", ptr(x), ptr(skip), ptr(out), ci(dt(x)), ci(B), ci(2 * H2), ci(2 * W2), ci(C1),
         ci(C2), _stream())
    _count()
    return out


def upcat_bwd(dout, C1, C2):
    B, H, W, Ct = dout.shape
    assert Ct == C1 + C2
    dx = torch.empty((B, H // 2, W // 2, C1), dtype=dout.dtype, device=dout.device)
    dskip = torch.empty((B, H, W, C2), dtype=dout.dtype, device=dout.device) if C2 else None
    call("upsample2x_concat_bwd", ptr(dout), ptr(dx), ptr(dskip), ci(dt(dout)), ci(B), ci(H), ci(W), ci(C1), ci(C2),
         _stream())
    _count()
    return dx, dskip


def gap_linear_sigmoid_fwd(x, w, b):
    B, H, W, C = x.shape
    pooled = torch.empty((B, C), dtype=torch.float32, device=x.device)
    y = torch.empty((B, 1), dtype=torch.float32, device=x.device)
    call("gap_linear_sigmoid_fwd", ptr(x), ci(dt(x)), ptr(w), ptr(b), ptr(pooled), ptr(y), ci(B), ll(H * W), ci(C),
         _stream())
    _count()
    return y, pooled


def gap_linear_sigmoid_bwd(dout, y, pooled, w, dw, db, x_shape, dtype, accumulate=True):
    B, H, W, C = x_shape
    dx = torch.empty((B, H, W, C), dtype=dtype, device=dout.device)
    call("gap_linear_sigmoid_bwd", ptr(dout), ptr(y), ptr(pooled), ptr(w), ptr(dw), ptr(db), ptr(dx), ci(dt(dx)),
         ci(B), ll(H * W), ci(C), ci(1 if accumulate else 0), _stream())
    _count()
    return dx


# ------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, clip_coef=None,
              dev_step=None):
    """``dev_step`` (int32[1] device tensor): graph-capturable mode, incremented by the call and used instead of ``step``."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _chk(t, "adam." + n, torch.float32)
    if dev_step is not None:
        _chk(dev_step, "adam.dev_step", torch.int32)
    call("adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), ll(p.numel()), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), ci(step), float(grad_scale), ptr(clip_coef), ptr(dev_step),
         _stream())
    _count(2 if dev_step is not None else 1)


def grad_clip_coef(g, max_norm, pre_scale=1.0):
    """Returns (coef[1], norm[1]) device tensors: coef = min(1, max_norm/(pre_scale*||g||+1e-6))."""
    out = torch.empty(2, dtype=torch.float32, device=g.device)
    ws = workspace(8, g.device)
    call("grad_clip_coef", ptr(g), ll(g.numel()), float(max_norm), float(pre_scale), ptr(out[0:1]), ptr(out[1:2]),
         ptr(ws), _stream())
    _count(2)
    return out[0:1], out[1:2]


# ------------------------------------------------------------------------------------------------
# losses / evaluation  (NCHW edge layout)
# ------------------------------------------------------------------------------------------------
CE_NONE, CE_PLAIN, CE_FOCAL = 0, 1, 2


def seg_loss(logits, target=None, soft_target=None, class_weights=None, ce_mode=CE_PLAIN, use_dice=False,
             alpha=0.25, gamma=2.0, mean=True, ignore_index=-100, smooth=1.0, w_ce=1.0, w_dice=1.0, out_scale=1.0):
    """Fused segmentation loss + gradient.  Returns (out4, grad): out4 = [ce, dice, total, n_bad]."""
    _chk(logits, "seg_loss.logits")
    B, C = logits.shape[:2]
    HW = logits.numel() // (B * C)
    if target is not None:
        _chk(target, "seg_loss.target", torch.int64)
        if target.numel() != B * HW:
            raise _lib.UdaError(f"seg_loss: target has {target.numel()} elements, expected {B * HW}")
    if soft_target is not None:
        _chk(soft_target, "seg_loss.soft_target", torch.float32)
        if soft_target.numel() != logits.numel():
            raise _lib.UdaError("seg_loss: soft target shape mismatch")
    if class_weights is not None:
        _chk(class_weights, "seg_loss.class_weights", torch.float32)
    grad = torch.empty_like(logits)
    out4 = torch.empty(4, dtype=torch.float32, device=logits.device)
    nbytes = _lib.lib().uda_seg_loss_workspace_bytes(ci(B), ci(C))
    ws = workspace(nbytes, logits.device)
    call("seg_loss_fwd_bwd", ptr(logits), ci(dt(logits)), ptr(target), ptr(soft_target), ptr(class_weights),
         ptr(grad), ptr(out4), ptr(ws), ci(B), ci(C), ll(HW), ci(ce_mode), ci(1 if use_dice else 0), float(alpha),
         float(gamma), ci(1 if mean else 0), ll(ignore_index), float(smooth), float(w_ce), float(w_dice),
         float(out_scale), _stream())
    _count(4)
    return out4, grad


def scale_by_device_scalar(x, scalar):
    call("scale_by_device_scalar", ptr(x), ci(dt(x)), ll(x.numel()), ptr(scalar), _stream())
    _count()
    return x


def consistency(z1, z2, temperature=0.5, out_scale=1.0):
    _chk(z1, "consistency.z1"); _chk(z2, "consistency.z2", z1.dtype)
    if z1.shape != z2.shape:
        raise _lib.UdaError("consistency: shape mismatch")
    B, C = z1.shape[:2]
    HW = z1.numel() // (B * C)
    g1, g2 = torch.empty_like(z1), torch.empty_like(z2)
    out = torch.empty(1, dtype=torch.float32, device=z1.device)
    ws = workspace(8, z1.device)
    call("consistency_fwd_bwd", ptr(z1), ptr(z2), ci(dt(z1)), ptr(g1), ptr(g2), ptr(out), ptr(ws), ci(B), ci(C),
         ll(HW), float(temperature), float(out_scale), _stream())
    _count(3)
    return out, g1, g2


def entropy(z, out_scale=1.0):
    _chk(z, "entropy.z")
    B, C = z.shape[:2]
    HW = z.numel() // (B * C)
    g = torch.empty_like(z)
    out = torch.empty(1, dtype=torch.float32, device=z.device)
    ws = workspace(8, z.device)
    call("entropy_fwd_bwd", ptr(z), ci(dt(z)), ptr(g), ptr(out), ptr(ws), ci(B), ci(C), ll(HW), float(out_scale),
         _stream())
    _count(3)
    return out, g


def bce_logits(x, label, scale=1.0, out=None, accumulate=False, want_grad=True):
    _chk(x, "bce.x", torch.float32)
    grad = torch.empty_like(x) if want_grad else None
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=x.device)
    call("bce_logits_fwd_bwd", ptr(x), ptr(grad), ptr(out), ll(x.numel()), float(label), float(scale),
         ci(1 if accumulate else 0), _stream())
    _count()
    return out, grad


def _chk_hist(hist, C, who):
    """A caller-supplied accumulator is indexed as C*C int64 by the kernels: anything else is out of bounds."""
    _chk(hist, f"{who}.hist", torch.int64)
    if tuple(hist.shape) != (C, C):
        raise _lib.UdaError(f"{who}: hist must be an int64 [{C},{C}] tensor (got {tuple(hist.shape)})")


def argmax_confmat(logits, target=None, num_classes=None, ignore_index=None, want_mask=True, mask_dtype=torch.int64,
                   hist=None):
    """argmax over dim 1 (+ confusion matrix against ``target``).  Returns (mask | None, hist | None)."""
    _chk(logits, "argmax_confmat.logits")
    B, C = logits.shape[:2]
    HW = logits.numel() // (B * C)
    spatial = tuple(logits.shape[2:])
    mask64 = mask8 = None
    if want_mask:
        if mask_dtype == torch.int64:
            mask64 = torch.empty((B,) + spatial, dtype=torch.int64, device=logits.device)
        elif mask_dtype == torch.uint8:
            mask8 = torch.empty((B,) + spatial, dtype=torch.uint8, device=logits.device)
        else:
            raise TypeError("mask dtype must be int64 or uint8")
    zero = 0
    if target is not None:
        _chk(target, "argmax_confmat.target", torch.int64)
        if num_classes is not None and num_classes != C:
            raise _lib.UdaError("argmax_confmat: num_classes must equal the logits' channel count")
        if hist is None:
            hist = torch.empty((C, C), dtype=torch.int64, device=logits.device)
            zero = 1
        else:
            _chk_hist(hist, C, "argmax_confmat")
    call("argmax_confmat", ptr(logits), ci(dt(logits)), ptr(target), ptr(mask64), ptr(mask8),
         ptr(hist) if target is not None else ptr(None), ci(B), ci(C), ll(HW),
         ll(ignore_index if ignore_index is not None else 0), ci(0 if ignore_index is None else 1), ci(zero), _stream())
    _count()
    return (mask64 if mask64 is not None else mask8), (hist if target is not None else None)


def confmat(pred, target, num_classes, ignore_index=None, hist=None):
    _chk(pred, "confmat.pred"); _chk(target, "confmat.target", torch.int64)
    if pred.numel() != target.numel():
        raise _lib.UdaError("confmat: pred/target size mismatch")
    if pred.dtype == torch.int64:
        pd = I64
    elif pred.dtype == torch.uint8:
        pd = U8
    else:
        raise TypeError("confmat: pred must be int64 or uint8")
    zero = 0
    if hist is None:
        hist = torch.empty((num_classes, num_classes), dtype=torch.int64, device=pred.device)
        zero = 1
    else:
        _chk_hist(hist, num_classes, "confmat")
    bad = torch.empty(1, dtype=torch.int64, device=pred.device)
    call("confmat", ptr(pred), ci(pd), ptr(target), ptr(hist), ptr(bad), ll(pred.numel()), ci(num_classes),
         ll(ignore_index if ignore_index is not None else 0), ci(0 if ignore_index is None else 1), ci(zero), _stream())
    _count()
    return hist, bad


def metrics_from_hist(hist):
    """Device-side metrics of an int64 [C,C] confusion matrix -> float64 [4 + 2C] device tensor (no host sync):
    [0] in-tree mean IoU, [1] pixel accuracy, [2] torchmetrics-style macro Jaccard, [3] pixels, [4:4+C] class IoU,
    [4+C:4+2C] per-class binary Jaccard (see ``uda_metrics_from_hist``)."""
    _chk(hist, "metrics_from_hist.hist", I64_T)
    C = hist.shape[0]
    if tuple(hist.shape) != (C, C):
        raise _lib.UdaError("metrics_from_hist: hist must be [C,C]")
    out = torch.empty(4 + 2 * C, dtype=torch.float64, device=hist.device)
    call("metrics_from_hist", ptr(hist), ci(C), ptr(out), _stream())
    _count()
    return out


def _f3(v):
    import ctypes
    return (ctypes.c_float * 3)(*[float(x) for x in v])


def gather_windows_u8(tile_hwc, win, first, count, mean, std, out=None):
    """Windows [first, first+count) of a uint8 [H,W,3] tile as normalised fp32 NCHW [count,3,win,win]."""
    _chk(tile_hwc, "gather_windows_u8.tile", torch.uint8)
    H, W, Cc = tile_hwc.shape
    if Cc != 3:
        raise _lib.UdaError("gather_windows_u8: tile must be [H,W,3] uint8")
    if out is None:
        out = torch.empty((count, 3, win, win), dtype=torch.float32, device=tile_hwc.device)
    call("gather_windows_u8", ptr(tile_hwc), ptr(out), ci(H), ci(W), ci(win), ci(first), ci(count), _f3(mean), _f3(std),
         _stream())
    _count()
    return out


def gather_label_windows(tile, win, first, count, out=None):
    """The same windows of an int64 / uint8 [H,W] label tile as int64 [count,win,win]."""
    _chk(tile, "gather_label_windows.tile")
    if tile.dtype == torch.int64:
        d = I64
    elif tile.dtype == torch.uint8:
        d = U8
    else:
        raise TypeError("gather_label_windows: tile must be int64 or uint8")
    H, W = tile.shape
    if out is None:
        out = torch.empty((count, win, win), dtype=torch.int64, device=tile.device)
    call("gather_label_windows", ptr(tile), ci(d), ptr(out), ci(H), ci(W), ci(win), ci(first), ci(count), _stream())
    _count()
    return out


def scatter_window_masks(masks, tile_mask, win, first):
    """uint8 [count,win,win] window masks -> their place in the uint8 [H,W] tile mask."""
    _chk(masks, "scatter_window_masks.masks", torch.uint8); _chk(tile_mask, "scatter_window_masks.tile_mask", torch.uint8)
    H, W = tile_mask.shape
    call("scatter_window_masks", ptr(masks), ptr(tile_mask), ci(H), ci(W), ci(win), ci(first), ci(masks.shape[0]), _stream())
    _count()
    return tile_mask


def strong_augment(images, table):
    """One augmented view of an fp32 NCHW batch from a [B,12] device parameter table (see ``augment.build_table``)."""
    _chk(images, "strong_augment.images", torch.float32); _chk(table, "strong_augment.table", torch.float32)
    B, C, H, W = images.shape
    if tuple(table.shape) != (B, 12):
        raise _lib.UdaError("strong_augment: table must be [B,12] float32")
    out = torch.empty_like(images)
    call("strong_augment", ptr(images), ptr(out), ptr(table), ci(B), ci(C), ci(H), ci(W), _stream())
    _count()
    return out
