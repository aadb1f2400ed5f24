"""Per-kernel PyTorch restatements of the C-ABI ops — TEST INFRASTRUCTURE (see oracle/__init__.py).

Same function names and tensor conventions as ``uda_aerial_semantic_segmentation_research_b200/ops.py``
(NHWC activations, OHWI weights), implemented with plain fp32 torch ops so that

  * each CUDA kernel has a reference to be compared against on the GPU box (``tests/test_gpu_*``), and
  * the engine's hand-written backward (tape, residual / skip gradient folding) can be validated on
    CPU against autograd of the fp32 oracle U-Net by monkeypatching ``engine.ops`` in a test.

The product never imports this module.
"""
import torch
import torch.nn.functional as F

LAUNCHES = 0
USE_TC = False
TC_PERSIST = False
FOLD_BN_EVAL = False
WGRAD_STREAM = False
FUSE_UPCAT = False


def dgrad_bnstats_supported(*a):
    """The oracle has no fused epilogues: the engine then takes the separate BatchNorm-backward reduce path."""
    return False


def stem_supported(*a):
    return False
#: arithmetic dtype of the restatements; tests switch it to float64 to check the engine's backward
#: *logic* free of the ReLU-mask-flip noise that fp32 round-off causes in deep randomly-initialised nets
_F = torch.float32


def set_precision(dtype):
    global _F
    _F = dtype

CE_NONE, CE_PLAIN, CE_FOCAL = 0, 1, 2


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _w_oihw(w):
    return w.permute(0, 3, 1, 2).to(_F)


def nchw_to_nhwc(x, dtype, cpad=None):
    y = _nhwc(x.to(_F))
    if cpad and cpad > y.shape[-1]:
        y = F.pad(y, (0, cpad - y.shape[-1]))
    return y.to(dtype).contiguous()


def nhwc_to_nchw(x, C=None):
    C = C or x.shape[-1]
    return _nchw(x[..., :C].to(_F)).contiguous()


def cast_f32(src, dst):
    dst.copy_(src.to(dst.dtype))
    return dst


def conv_fwd(x, w, bias=None, stride=1, pad=1, nchw_out=False, bn_sums=None, force_direct=False):
    y = F.conv2d(_nchw(x.to(_F)), _w_oihw(w), bias.to(_F) if bias is not None else None, stride, pad)
    if nchw_out:
        return y.contiguous()
    return _nhwc(y).to(x.dtype)


def conv_dgrad(dy, w, x_shape, stride=1, pad=1, addend=None, force_direct=False, w_ft=None, bn_stats=None):
    B, H, W, Cin = x_shape
    dx = torch.nn.grad.conv2d_input((B, Cin, H, W), _w_oihw(w), _nchw(dy.to(_F)), stride, pad)
    dx = _nhwc(dx)
    if addend is not None:
        addend.copy_((addend.to(_F) + dx).to(addend.dtype))
        return addend
    return dx.to(dy.dtype)


def conv_wgrad(dy, x, dw, stride=1, pad=1, force_direct=False):
    O, KH, KW, I = dw.shape
    g = torch.nn.grad.conv2d_weight(_nchw(x.to(_F)), (O, I, KH, KW), _nchw(dy.to(_F)), stride, pad)
    dw += g.permute(0, 2, 3, 1)
    return dw


def bn_stats(x, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1):
    C = x.shape[-1]
    xf = x.to(_F).reshape(-1, C)
    M = xf.shape[0]
    mean = xf.mean(0)
    var = xf.var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + eps)
    g = gamma if gamma is not None else torch.ones_like(mean)
    b = beta if beta is not None else torch.zeros_like(mean)
    scale = g * rstd
    shift = b - mean * scale
    if running_mean is not None:
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(momentum * mean)
            running_var.mul_(1 - momentum).add_(momentum * var * (M / max(M - 1, 1)))
    return mean, rstd, scale.detach(), shift.detach()


def bn_eval_coeffs(gamma, beta, running_mean, running_var, eps=1e-5):
    rstd = 1.0 / torch.sqrt(running_var + eps)
    scale = gamma * rstd
    return scale.detach(), (beta - running_mean * scale).detach()


def _act(v, slope):
    return torch.where(v > 0, v, v * slope)


def bn_apply(x, scale, shift, residual=None, slope=0.0, out=None):
    v = x.to(_F) * scale + shift
    if residual is not None:
        v = v + residual.to(_F)
    y = _act(v, slope).to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def bn_bwd(dy, x, a, gamma, mean, rstd, slope, dgamma, dbeta, dres=None, dres_accumulate=False,
           param_accumulate=True, scale=None, shift=None):
    C = x.shape[-1]
    g = dy.to(_F)
    if a is not None:
        g = g * torch.where(a.to(_F) > 0, torch.ones_like(g), torch.full_like(g, slope))
    elif scale is not None:
        pre = x.to(_F) * scale + shift
        g = g * torch.where(pre > 0, torch.ones_like(g), torch.full_like(g, slope))
    xhat = (x.to(_F) - mean) * rstd
    M = x.numel() // C
    s1 = g.reshape(-1, C).sum(0)
    s2 = (g * xhat).reshape(-1, C).sum(0)
    gam = gamma.detach() if gamma is not None else torch.ones_like(s1)
    k0 = gam * rstd
    dx = k0 * (g - s1 / M - xhat * s2 / M)
    if dgamma is not None:
        dgamma.copy_((dgamma if param_accumulate else 0) + s2)
    if dbeta is not None:
        dbeta.copy_((dbeta if param_accumulate else 0) + s1)
    if dres is not None:
        dres.copy_(((dres.to(_F) if dres_accumulate else 0) + g).to(dres.dtype))
    return dx.to(x.dtype)


def act_bwd(dy, a, slope):
    return (dy.to(_F) * torch.where(a.to(_F) > 0, 1.0, slope)).to(dy.dtype)


def bias_act(x, bias, slope, out=None):
    v = x.to(_F) + (bias if bias is not None else 0)
    y = _act(v, slope).to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def colsum(x, out, scale=1.0, accumulate=True):
    s = x.to(_F).reshape(-1, x.shape[-1]).sum(0) * scale
    out.copy_((out if accumulate else 0) + s)
    return out


def maxpool_fwd(x):
    y, idx = F.max_pool2d(_nchw(x.to(_F)), 3, 2, 1, return_indices=True)
    return _nhwc(y).to(x.dtype), idx  # idx: flat h*W+w indices in NCHW (oracle-private format)


def maxpool_bwd(dy, idx, x_shape, addend=None):
    B, H, W, C = x_shape
    g = _nchw(dy.to(_F)).reshape(B, C, -1)
    dx = torch.zeros(B, C, H * W, dtype=_F).scatter_add_(2, idx.reshape(B, C, -1), g).reshape(B, C, H, W)
    dx = _nhwc(dx)
    if addend is not None:
        addend.copy_((addend.to(_F) + dx).to(addend.dtype))
        return addend
    return dx.to(dy.dtype)


def upcat_fwd(x, skip=None):
    up = x.repeat_interleave(2, 1).repeat_interleave(2, 2)
    return (torch.cat([up, skip], -1) if skip is not None else up).contiguous()


def upcat_bwd(dout, C1, C2):
    B, H, W, _ = dout.shape
    d1 = dout[..., :C1].to(_F).reshape(B, H // 2, 2, W // 2, 2, C1).sum((2, 4)).to(dout.dtype)
    return d1.contiguous(), (dout[..., C1:].contiguous() if C2 else None)


def gap_linear_sigmoid_fwd(x, w, b):
    pooled = x.to(_F).mean((1, 2))
    y = torch.sigmoid(pooled @ w.detach().to(_F).t() + b.detach().to(_F))
    return y, pooled


def gap_linear_sigmoid_bwd(dout, y, pooled, w, dw, db, x_shape, dtype, accumulate=True):
    B, H, W, C = x_shape
    dz = dout.to(_F) * y * (1 - y)
    dw.copy_((dw if accumulate else 0) + (dz.t() @ pooled).to(dw.dtype))
    db.copy_((db if accumulate else 0) + dz.sum(0).to(db.dtype))
    dx = (dz @ w.detach().to(_F)).reshape(B, 1, 1, C).expand(B, H, W, C) / (H * W)
    return dx.to(dtype).contiguous()


def adam_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, clip_coef=None,
              dev_step=None):
    if dev_step is not None:
        dev_step += 1
        step = int(dev_step)
    gi = g * grad_scale * (clip_coef if clip_coef is not None else 1.0)
    if weight_decay:
        gi = gi + weight_decay * p
    m.mul_(beta1).add_((1 - beta1) * gi)
    v.mul_(beta2).add_((1 - beta2) * gi * gi)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    p.sub_((lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + eps))
    if shadow is not None:
        shadow.copy_(p.to(shadow.dtype))


def grad_clip_coef(g, max_norm, pre_scale=1.0):
    norm = g.to(_F).norm() * pre_scale
    return torch.clamp(max_norm / (norm + 1e-6), max=1.0).reshape(1), norm.reshape(1)


# ---- loss / eval ops (NCHW edge layout): built on the restated reference losses ------------------
def seg_loss(logits, target=None, soft_target=None, class_weights=None, ce_mode=CE_PLAIN, use_dice=False,
             alpha=0.25, gamma=2.0, mean=True, ignore_index=-100, smooth=1.0, w_ce=1.0, w_dice=1.0, out_scale=1.0):
    from . import ref_losses as R
    z = logits.detach().to(_F).requires_grad_()
    ce = torch.zeros(())
    dice = torch.zeros(())
    if ce_mode == CE_PLAIN:
        ce = F.cross_entropy(z, target, weight=class_weights, ignore_index=ignore_index,
                             reduction="mean" if mean else "sum")
    elif ce_mode == CE_FOCAL:
        w = class_weights if class_weights is not None else torch.ones(z.shape[1])
        c = F.cross_entropy(z, target, reduction="none", weight=w)
        f = alpha * (1 - torch.exp(-c)) ** gamma * c
        ce = f.mean() if mean else f.sum()
    if use_dice:
        dice = R.dice_loss(z, soft_target if soft_target is not None else target, smooth)
    total = out_scale * (w_ce * ce + w_dice * dice)
    (grad,) = torch.autograd.grad(total, z)
    out4 = torch.stack([ce.detach(), dice.detach(), total.detach(), torch.zeros(())]).to(_F)
    return out4, grad.to(logits.dtype)


def scale_by_device_scalar(x, scalar):
    x.mul_(scalar.reshape(()).to(x.dtype))
    return x


def consistency(z1, z2, temperature=0.5, out_scale=1.0):
    from . import ref_losses as R
    a = z1.detach().to(_F).requires_grad_()
    b = z2.detach().to(_F).requires_grad_()
    loss = out_scale * R.consistency_loss(a, b, temperature)
    g1, g2 = torch.autograd.grad(loss, [a, b])
    return loss.detach().reshape(1), g1.to(z1.dtype), g2.to(z2.dtype)


def entropy(z, out_scale=1.0):
    from . import ref_losses as R
    a = z.detach().to(_F).requires_grad_()
    loss = out_scale * R.entropy_loss(a)
    (g,) = torch.autograd.grad(loss, a)
    return loss.detach().reshape(1), g.to(z.dtype)


def bce_logits(x, label, scale=1.0, out=None, accumulate=False, want_grad=True):
    a = x.detach().to(_F).requires_grad_()
    loss = scale * F.binary_cross_entropy_with_logits(a, torch.full_like(a, label))
    (g,) = torch.autograd.grad(loss, a)
    val = loss.detach().reshape(1)
    if out is not None:
        out.copy_((out if accumulate else 0) + val)
    else:
        out = val
    return out, (g if want_grad else None)


def argmax_confmat(logits, target=None, num_classes=None, ignore_index=None, want_mask=True,
                   mask_dtype=torch.int64, hist=None):
    from . import ref_metrics as M
    C = logits.shape[1]
    mask = torch.from_numpy(M.argmax_mask(logits.detach().to(_F).cpu().numpy()))
    h = None
    if target is not None:
        h = torch.from_numpy(M.fast_hist(mask.numpy(), target.cpu().numpy(), C, ignore_index))
        if hist is not None:
            hist += h
            h = hist
    return (mask.to(mask_dtype) if want_mask else None), h


def confmat(pred, target, num_classes, ignore_index=None, hist=None):
    from . import ref_metrics as M
    h = torch.from_numpy(M.fast_hist(pred.cpu().numpy(), target.cpu().numpy(), num_classes, ignore_index))
    if hist is not None:
        hist += h
        h = hist
    return h, torch.zeros(1, dtype=torch.int64)


def weight_flip_transpose_batch(base, out, table):
    for off, O, I, KH, KW in table.tolist():
        n = O * I * KH * KW
        w = base[off:off + n].view(O, KH, KW, I)
        out[off:off + n].copy_(w.flip(1, 2).permute(3, 1, 2, 0).reshape(-1))
    return out
