"""Oracle restatement of the reference loss path — TEST INFRASTRUCTURE (see oracle/__init__.py).

Plain PyTorch (CPU, fp32 or fp64), written as explicit formulas so that autograd
provides the reference gradients.  Each function cites the reference lines it
follows (paths relative to the reference tree).  Pinned against the reference's
own modules by ``oracle/gen_golden.py`` -> ``tests/golden/losses_*.npz`` and live
in ``tests/test_oracle.py`` when /root/reference is present.
"""
import torch


def _log_softmax(z, dim=1):
    m = z.max(dim=dim, keepdim=True).values
    return z - m - (z - m).exp().sum(dim=dim, keepdim=True).log()


def cross_entropy(logits, target, ignore_index=-100):
    """nn.CrossEntropyLoss() defaults — src/models/train.py:208,342 (mean over non-ignored pixels)."""
    ls = _log_softmax(logits, 1)
    valid = target != ignore_index
    t = target.clamp(min=0)
    picked = ls.gather(1, t.unsqueeze(1)).squeeze(1)
    return -(picked * valid).sum() / valid.sum()


def dice_loss(logits, target, smooth=1.0):
    """DiceLoss.forward — src/models/losses.py:118-152.

    softmax over C (:131); index targets are one-hot encoded (:134-142); per-(b,c)
    I = sum p*t, U = sum p + sum t (:145-146); dice=(2I+s)/(U+s) (:149); 1-mean (:152).
    """
    B, C = logits.shape[:2]
    p = _log_softmax(logits, 1).exp()
    if target.dim() == 3:
        t = torch.zeros_like(p).scatter_(1, target.long().unsqueeze(1), 1.0)
    else:
        t = target.to(p.dtype)
    inter = (p * t).flatten(2).sum(-1)
    union = p.flatten(2).sum(-1) + t.flatten(2).sum(-1)
    return 1.0 - ((2.0 * inter + smooth) / (union + smooth)).mean()


def weighted_segmentation_loss(logits, target, class_weights=None, alpha=0.25, gamma=2.0,
                               reduction="mean", domain_weight=1.0):
    """WeightedSegmentationLoss — src/models/losses.py:154-215.

    ce_i = w[y_i] * (-log p_{y_i}) (:180-181, reduction='none' so no weight normalisation);
    pt = exp(-ce) (:182); focal = alpha*(1-pt)^gamma*ce (:183); mean|sum (:185-187);
    + DiceLoss on the one-hot targets (:207-211); * domain_weight (:215).
    """
    C = logits.shape[1]
    w = torch.ones(C, dtype=logits.dtype) if class_weights is None else class_weights.to(logits.dtype)
    ls = _log_softmax(logits, 1)
    ce = -ls.gather(1, target.unsqueeze(1)).squeeze(1) * w[target]
    pt = (-ce).exp()
    focal = alpha * (1 - pt) ** gamma * ce
    focal = focal.mean() if reduction == "mean" else focal.sum()
    return domain_weight * (focal + dice_loss(logits, target))


def bce_with_logits(x, y):
    """nn.BCEWithLogitsLoss() mean — src/models/losses.py:16."""
    return (x.clamp(min=0) - x * y + (1 + (-x.abs()).exp()).log()).mean()


def discriminator_loss(src_pred, tgt_pred):
    """AdversarialLoss.discriminator_loss — src/models/losses.py:18-36 (labels src=1, tgt=0, averaged)."""
    return (bce_with_logits(src_pred, torch.ones_like(src_pred))
            + bce_with_logits(tgt_pred, torch.zeros_like(tgt_pred))) / 2


def generator_loss(tgt_pred, lambda_adv=0.001):
    """AdversarialLoss.generator_loss — src/models/losses.py:38-51."""
    return lambda_adv * bce_with_logits(tgt_pred, torch.ones_like(tgt_pred))


def consistency_loss(z1, z2, temperature=0.5):
    """ConsistencyLoss.forward — src/models/losses.py:62-90.

    p_i = softmax(z_i/T); kl_div(log p1, p2, 'batchmean') = sum p2*(log p2 - log p1) / B (:78-82);
    symmetric mean of both directions (:90).  Gradients flow through log p AND the
    probability "targets" (the reference does not detach them).
    """
    B = z1.shape[0]
    l1 = _log_softmax(z1 / temperature, 1)
    l2 = _log_softmax(z2 / temperature, 1)
    p1, p2 = l1.exp(), l2.exp()
    kl1 = (p2 * (l2 - l1)).sum() / B
    kl2 = (p1 * (l1 - l2)).sum() / B
    return (kl1 + kl2) / 2


def entropy_loss(logits):
    """Target-domain entropy minimisation: mean over pixels of -sum_c p log p.

    NOT in the reference — north-star extension; **parity unpinned**
    by the reference, pinned only against this expression.
    """
    ls = _log_softmax(logits, 1)
    return -(ls.exp() * ls).sum(1).mean()


def rampup(epoch, rampup_length=40):
    """FineTuningLoss.rampup — src/models/losses.py:279-285."""
    return 1.0 if epoch >= rampup_length else float(epoch) / rampup_length


def fine_tuning_loss(pred1, pred2, domain_pred, epoch, supervised_pred=None, supervised_target=None,
                     consistency_weight=1.0, domain_weight=0.1, supervised_weight=0.1,
                     rampup_length=40, temperature=0.5):
    """FineTuningLoss.forward — src/models/losses.py:287-342 (note the double domain weight :271,:318-319)."""
    r = rampup(epoch, rampup_length)
    cons = consistency_loss(pred1, pred2, temperature)
    dom = generator_loss(domain_pred, lambda_adv=domain_weight)
    total = cons * consistency_weight * r + dom * domain_weight * r
    sup = torch.tensor(0.0)
    if supervised_pred is not None and supervised_target is not None:
        sup = dice_loss(supervised_pred, supervised_target.long())
        total = total + sup * supervised_weight
    return {"total": total, "consistency": cons.detach(), "domain_confusion": dom.detach(),
            "supervised": sup.detach(), "rampup_weight": torch.tensor(r)}


class _GRL(torch.autograd.Function):
    """GradientReverseFunction — src/models/uda.py:103-112 (fwd identity, bwd -alpha*g)."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.alpha * g, None


def gradient_reverse_layer(x, alpha):
    """src/models/uda.py:99-101."""
    return _GRL.apply(x, alpha)
