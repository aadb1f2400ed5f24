"""Loss entry points of the reference (``src/models/losses.py``, ``src/models/uda.py:99-112``) on the
fused CUDA kernels: every loss returns a 0-dim autograd tensor on the input's device, and the
logit-gradient is produced by the same kernel pass that computes the loss value.

Same class names, constructor arguments and call signatures as the reference so that the trainers'
call sites (``criterion(outputs, masks)``, ``adversarial_loss.discriminator_loss(s, t)``,
``FineTuningLoss(...)(pred1, pred2, domain_pred, epoch, ...)``) work unchanged.  CUDA tensors only:
there is no CPU fallback (a CPU tensor raises).
"""
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: uda_b200 losses run on CUDA tensors only (no CPU fallback)")


class _GradHolder:
    """Gradient computed in the forward pass, scaled by the incoming grad_output in backward."""

    @staticmethod
    def finish(grad, gout):
        # grad_output is a 0-dim device scalar; the kernel exits immediately when it equals 1
        return ops.scale_by_device_scalar(grad, gout.reshape(1).float().contiguous())


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, soft_target, class_weights, kw):
        out4, grad = ops.seg_loss(logits, target, soft_target, class_weights, **kw)
        ctx.grad = grad
        ctx.out4 = out4
        return out4[2].clone()

    @staticmethod
    def backward(ctx, gout):
        g = _GradHolder.finish(ctx.grad, gout)
        ctx.grad = None
        return g, None, None, None, None


def _seg_loss(logits, target=None, soft_target=None, class_weights=None, **kw):
    _require_cuda(logits, "segmentation loss")
    if logits.dim() < 3:
        raise ValueError("expected logits of shape [B,C,...]")
    z = logits.contiguous()
    if z.dtype not in (torch.float32, torch.bfloat16):
        z = z.float()
    if target is not None:
        target = target.to(device=z.device, dtype=torch.long).contiguous()
    if soft_target is not None:
        soft_target = soft_target.to(device=z.device, dtype=torch.float32).contiguous()
    if class_weights is not None:
        class_weights = class_weights.to(device=z.device, dtype=torch.float32).contiguous()
    return _SegLossFn.apply(z, target, soft_target, class_weights, kw)


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss()`` as the reference uses it (``src/models/train.py:208``): mean over
    non-ignored pixels, optional class weights; fused single-pass loss + gradient."""

    def __init__(self, weight: Optional[torch.Tensor] = None, ignore_index: int = -100, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise NotImplementedError("reduction must be 'mean' or 'sum'")
        self.register_buffer("weight", weight)
        self.ignore_index, self.reduction = ignore_index, reduction

    def forward(self, inputs, targets):
        return _seg_loss(inputs, targets, class_weights=self.weight, ce_mode=ops.CE_PLAIN, use_dice=False,
                         mean=self.reduction == "mean", ignore_index=self.ignore_index)


class DiceLoss(nn.Module):
    """Reference ``DiceLoss`` (``src/models/losses.py:110-152``): softmax, per-(b,c) soft Dice with
    ``smooth``, ``1 - mean``.  Targets: class indices ``[B,H,W]`` or one-hot/soft ``[B,C,H,W]``."""

    def __init__(self, smooth=1.0):
        super().__init__()
        self.smooth = smooth

    def forward(self, predictions, targets):
        if targets.dim() == predictions.dim() - 1:
            return _seg_loss(predictions, targets, ce_mode=ops.CE_NONE, use_dice=True, smooth=self.smooth)
        if targets.shape != predictions.shape:
            raise ValueError("DiceLoss: targets must be [B,H,W] indices or match the predictions' shape")
        return _seg_loss(predictions, None, soft_target=targets, ce_mode=ops.CE_NONE, use_dice=True,
                         smooth=self.smooth)


class CombinedCEDiceLoss(nn.Module):
    """``CrossEntropyLoss() + DiceLoss()`` (BASELINE config 1, ``test_system`` loss suites) in one fused
    two-pass kernel pair: ``ce_weight * CE + dice_weight * Dice``."""

    def __init__(self, ce_weight=1.0, dice_weight=1.0, smooth=1.0, ignore_index=-100):
        super().__init__()
        self.ce_weight, self.dice_weight, self.smooth, self.ignore_index = ce_weight, dice_weight, smooth, ignore_index

    def forward(self, inputs, targets):
        return _seg_loss(inputs, targets, ce_mode=ops.CE_PLAIN, use_dice=True, smooth=self.smooth,
                         w_ce=self.ce_weight, w_dice=self.dice_weight, ignore_index=self.ignore_index)


class WeightedSegmentationLoss(nn.Module):
    """Reference ``WeightedSegmentationLoss`` (``src/models/losses.py:154-215``):
    ``domain_weight * (focal(weighted CE) + Dice(one-hot))``."""

    def __init__(self, num_classes: int, class_weights: Optional[torch.Tensor] = None, alpha: float = 0.25,
                 gamma: float = 2.0, reduction: str = "mean"):
        super().__init__()
        self.num_classes = num_classes
        self.register_buffer("class_weights", class_weights if class_weights is not None
                             else torch.ones(num_classes))
        self.alpha, self.gamma, self.reduction = alpha, gamma, reduction
        self.dice_loss = DiceLoss()

    def forward(self, inputs: torch.Tensor, targets: torch.Tensor, domain_weight: float = 1.0) -> torch.Tensor:
        if inputs.shape[1] != self.num_classes:
            raise ValueError("WeightedSegmentationLoss: channel count != num_classes")
        return _seg_loss(inputs, targets, class_weights=self.class_weights, ce_mode=ops.CE_FOCAL, use_dice=True,
                         alpha=self.alpha, gamma=self.gamma, mean=self.reduction == "mean",
                         smooth=self.dice_loss.smooth, out_scale=float(domain_weight))


class _BCEPairFn(torch.autograd.Function):
    """sum_i scale_i * mean BCEWithLogits(x_i, label_i) over up to two inputs."""

    @staticmethod
    def forward(ctx, spec, *xs):
        out = None
        grads = []
        for x, (label, scale) in zip(xs, spec):
            out, g = ops.bce_logits(x, label, scale, out=out, accumulate=out is not None)
            grads.append(g)
        ctx.grads = grads
        return out.reshape(())

    @staticmethod
    def backward(ctx, gout):
        gs = [_GradHolder.finish(g, gout) for g in ctx.grads]
        ctx.grads = None
        return (None,) + tuple(gs)


def _as_bce_input(t, who):
    _require_cuda(t, who)
    return t.contiguous().float()


class AdversarialLoss:
    """Reference ``AdversarialLoss`` (``src/models/losses.py:7-51``).  Inputs are the discriminator's
    outputs (already sigmoided in the reference — the BCE-with-logits on top of them is reproduced)."""

    def __init__(self, lambda_adv=0.001):
        self.lambda_adv = lambda_adv

    def discriminator_loss(self, source_pred, target_pred):
        s = _as_bce_input(source_pred, "discriminator_loss")
        t = _as_bce_input(target_pred, "discriminator_loss")
        return _BCEPairFn.apply(((1.0, 0.5), (0.0, 0.5)), s, t)

    def generator_loss(self, target_pred):
        t = _as_bce_input(target_pred, "generator_loss")
        return _BCEPairFn.apply(((1.0, float(self.lambda_adv)),), t)


class _ConsistencyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, temperature, out_scale):
        out, g1, g2 = ops.consistency(z1, z2, temperature, out_scale)
        ctx.g = (g1, g2)
        return out.reshape(())

    @staticmethod
    def backward(ctx, gout):
        g1, g2 = ctx.g
        ctx.g = None
        return _GradHolder.finish(g1, gout), _GradHolder.finish(g2, gout), None, None


def _logits(t, who):
    _require_cuda(t, who)
    z = t.contiguous()
    return z if z.dtype in (torch.float32, torch.bfloat16) else z.float()


class ConsistencyLoss(nn.Module):
    """Reference ``ConsistencyLoss`` (``src/models/losses.py:53-108``): symmetric KL between
    ``softmax(pred/T)`` of two views, ``batchmean``; one fused pass emits the loss and both gradients."""

    def __init__(self, temperature=0.5):
        super().__init__()
        self.temperature = temperature

    def forward(self, pred1, pred2, _scale: float = 1.0):
        return _ConsistencyFn.apply(_logits(pred1, "ConsistencyLoss"), _logits(pred2, "ConsistencyLoss"),
                                    float(self.temperature), float(_scale))

    def get_similarity_matrix(self, pred1, pred2):
        """Visualisation helper (``losses.py:92-108``), not on the training path: cosine similarity of the
        two softmax maps, computed with plain torch ops on the inputs' device."""
        p1 = torch.softmax(pred1.float(), dim=1)
        p2 = torch.softmax(pred2.float(), dim=1)
        return torch.nn.functional.cosine_similarity(p1, p2, dim=1)


class _EntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, out_scale):
        out, g = ops.entropy(z, out_scale)
        ctx.g = g
        return out.reshape(())

    @staticmethod
    def backward(ctx, gout):
        g = _GradHolder.finish(ctx.g, gout)
        ctx.g = None
        return g, None


class EntropyMinimizationLoss(nn.Module):
    """Target-domain entropy minimisation ``mean_px(-sum_c p log p)`` — north-star extension; the
    reference has no such loss, parity is pinned against the plain torch expression."""

    def __init__(self, weight=1.0):
        super().__init__()
        self.weight = weight

    def forward(self, logits):
        return _EntropyFn.apply(_logits(logits, "EntropyMinimizationLoss"), float(self.weight))


def calculate_class_weights(dataset, num_classes: int, method: str = "effective_samples") -> torch.Tensor:
    """Reference ``calculate_class_weights`` (``src/models/losses.py:217-254``); host-side dataset
    statistics (not on the GPU hot path).  Class counting uses one bincount per mask."""
    class_counts = torch.zeros(num_classes, dtype=torch.float64)
    for _, mask in dataset:
        m = torch.as_tensor(mask).reshape(-1).long()
        m = m[(m >= 0) & (m < num_classes)]
        class_counts += torch.bincount(m, minlength=num_classes).double()
    class_counts = torch.clamp(class_counts.float(), min=1.0)
    if method == "effective_samples":
        beta = 0.9999
        weights = (1.0 - beta) / (1.0 - torch.pow(beta, class_counts))
    else:
        weights = 1.0 / class_counts
    return weights / weights.sum() * num_classes


class FineTuningLoss(nn.Module):
    """Reference ``FineTuningLoss`` (``src/models/losses.py:256-342``): ramped consistency + ramped
    domain confusion (note the reference's double ``domain_weight``) + optional supervised Dice."""

    def __init__(self, consistency_weight: float = 1.0, domain_weight: float = 0.1, supervised_weight: float = 0.1,
                 rampup_length: int = 40, temperature: float = 0.5):
        super().__init__()
        self.consistency_loss = ConsistencyLoss(temperature=temperature)
        self.domain_loss = AdversarialLoss(lambda_adv=domain_weight)
        self.supervised_loss = DiceLoss()
        self.consistency_weight = consistency_weight
        self.domain_weight = domain_weight
        self.supervised_weight = supervised_weight
        self.rampup_length = rampup_length

    def rampup(self, epoch: int) -> float:
        if epoch >= self.rampup_length:
            return 1.0
        return float(epoch) / self.rampup_length

    def forward(self, pred1, pred2, domain_pred, epoch, supervised_pred=None,
                supervised_target=None) -> Dict[str, torch.Tensor]:
        ramp = self.rampup(epoch)
        # the static weight is folded into the kernel (no extra pass over the gradients); at ramp == 0
        # the unweighted value is still reported and the gradient is scaled by 0 in backward
        wc = float(self.consistency_weight * ramp)
        if wc != 0.0:
            weighted_consistency = self.consistency_loss(pred1, pred2, _scale=wc)
            consistency = weighted_consistency.detach() / wc
        else:
            consistency_t = self.consistency_loss(pred1, pred2)
            weighted_consistency = consistency_t * 0.0
            consistency = consistency_t.detach()
        domain_confusion = self.domain_loss.generator_loss(domain_pred)
        total = weighted_consistency + domain_confusion * float(self.domain_weight * ramp)
        supervised = torch.zeros((), device=pred1.device)   # (a fill kernel: capturable in a CUDA graph)
        if supervised_pred is not None and supervised_target is not None:
            if supervised_target.dtype != torch.long:
                supervised_target = supervised_target.long()
            supervised = self.supervised_loss(supervised_pred, supervised_target)
            total = total + supervised * self.supervised_weight
        return {
            "total": total,
            "consistency": consistency,
            "domain_confusion": domain_confusion.detach(),
            "supervised": supervised.detach(),
            "rampup_weight": torch.tensor(ramp),
        }


class GradientReverseFunction(torch.autograd.Function):
    """Gradient-reversal layer with the reference's interface (``src/models/uda.py:103-112``; defined but never called
    by the reference's trainers): the forward is the identity, the backward multiplies the incoming
    gradient by ``-alpha`` in ONE kernel pass (``uda_scale``).  When the layer sits between softmax(logits) and the
    discriminator, ``discriminator.OutputSpaceAdversary`` folds the factor into the softmax-backward pass instead."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = float(alpha)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output
        if g.is_cuda and g.dtype in (torch.float32, torch.bfloat16):
            return ops.scale(g.contiguous(), -ctx.alpha), None
        raise RuntimeError("gradient_reverse_layer: CUDA float32 / bfloat16 gradients only (no CPU fallback)")


def gradient_reverse_layer(x, alpha):
    """``src/models/uda.py:99-101``."""
    return GradientReverseFunction.apply(x, alpha)
