"""Generate golden fixtures by EXECUTING the reference's own modules — TEST INFRASTRUCTURE.

Run in the build container (the reference tree is not on the GPU box):

    python oracle/gen_golden.py [/root/reference]

Imports ``src.models.losses``, ``src.models.discriminator`` and
``src.analysis.metrics`` from the reference tree (they import cleanly; the U-Net
cannot be generated this way because ``segmentation_models_pytorch`` is absent —
see oracle/ref_unet.py) and writes small seeded input/output vectors, including
autograd gradients, to ``tests/golden/*.npz``.
"""
import os
import sys
import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, REF)
from src.models import losses as L                      # noqa: E402
from src.models.discriminator import DomainDiscriminator  # noqa: E402
from src.analysis.metrics import SegmentationMetrics      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def g(seed):
    return torch.Generator().manual_seed(seed)


def losses_case(name, B, C, H, W, seed, scale=3.0, blocky=False):
    z1 = (torch.randn(B, C, H, W, generator=g(seed)) * scale).requires_grad_()
    z2 = (torch.randn(B, C, H, W, generator=g(seed + 1)) * scale).requires_grad_()
    if blocky:
        t = torch.randint(0, C, (B, max(H // 4, 1), max(W // 4, 1)), generator=g(seed + 2))
        t = t.repeat_interleave(4, 1).repeat_interleave(4, 2)[:, :H, :W].contiguous()
    else:
        t = torch.randint(0, C, (B, H, W), generator=g(seed + 2))
    w = torch.rand(C, generator=g(seed + 3)) + 0.5
    out = {"z1": z1.detach().numpy(), "z2": z2.detach().numpy(), "target": t.numpy(), "class_weights": w.numpy()}

    def rec(key, loss, wrt):
        grads = torch.autograd.grad(loss, wrt)
        out[key] = np.float64(loss.item())
        for i, gr in enumerate(grads):
            out[f"{key}_grad{i}"] = gr.numpy()

    rec("ce", torch.nn.CrossEntropyLoss()(z1, t), [z1])
    rec("dice", L.DiceLoss()(z1, t), [z1])
    rec("ce_plus_dice", torch.nn.CrossEntropyLoss()(z1, t) + L.DiceLoss()(z1, t), [z1])
    rec("weighted", L.WeightedSegmentationLoss(C, w)(z1, t, domain_weight=0.7), [z1])
    rec("weighted_noweights_sum", L.WeightedSegmentationLoss(C, reduction="sum")(z1, t), [z1])
    rec("consistency", L.ConsistencyLoss(temperature=0.5)(z1, z2), [z1, z2])
    rec("consistency_T1", L.ConsistencyLoss(temperature=1.0)(z1, z2), [z1, z2])
    d = torch.rand(B, 1, generator=g(seed + 4)).requires_grad_()
    s = torch.rand(B, 1, generator=g(seed + 5)).requires_grad_()
    out["d_src"], out["d_tgt"] = s.detach().numpy(), d.detach().numpy()
    adv = L.AdversarialLoss(lambda_adv=0.001)
    rec("disc_loss", adv.discriminator_loss(s, d), [s, d])
    rec("gen_loss", adv.generator_loss(d), [d])
    ft = L.FineTuningLoss()
    for ep in (0, 20, 60):
        r = ft(z1, z2, d, ep, supervised_pred=z1, supervised_target=t)
        out[f"ft_total_ep{ep}"] = np.float64(r["total"].item())
        out[f"ft_ramp_ep{ep}"] = np.float64(r["rampup_weight"].item())
        if ep == 20:
            gz1, gz2, gd = torch.autograd.grad(r["total"], [z1, z2, d])
            out["ft_ep20_grad_z1"], out["ft_ep20_grad_z2"], out["ft_ep20_grad_d"] = gz1.numpy(), gz2.numpy(), gd.numpy()
            out["ft_ep20_consistency"] = np.float64(r["consistency"].item())
            out["ft_ep20_domain"] = np.float64(r["domain_confusion"].item())
            out["ft_ep20_supervised"] = np.float64(r["supervised"].item())
    # evaluation path on the same logits: argmax (predict.py:129) + confusion matrix (analysis/metrics.py:17-42)
    pred = z1.detach().argmax(dim=1)
    sm = SegmentationMetrics(C)
    out["argmax"] = pred.numpy()
    out["hist"] = sm._fast_hist(pred.flatten(), t.flatten())
    iou = sm.batch_iou(pred, t)
    out["mean_iou"] = np.float64(iou["mean_iou"])
    out["class_iou"] = np.array([iou["class_iou"][i] for i in range(C)])
    out["pixel_acc"] = np.float64(sm.pixel_accuracy(pred, t))
    out["f1"] = np.array(sm.f1_score(pred, t))
    sm_ign = SegmentationMetrics(C, ignore_index=0)
    out["hist_ignore0"] = sm_ign._fast_hist(pred.flatten(), t.flatten())
    np.savez_compressed(os.path.join(OUT, f"losses_{name}.npz"), **out)
    print("wrote", name, {k: float(out[k]) for k in ("ce", "dice", "weighted", "consistency", "disc_loss", "gen_loss")})


def discriminator_case():
    torch.manual_seed(11)
    D = DomainDiscriminator(3)
    x = torch.randn(2, 3, 64, 64, generator=g(12), requires_grad=True)
    D.train()
    y = D(x)
    loss = L.AdversarialLoss().discriminator_loss(y[:1], y[1:])
    grads = torch.autograd.grad(loss, [x] + list(D.parameters()))
    out = {"x": x.detach().numpy(), "y_train": y.detach().numpy(), "loss": np.float64(loss.item()),
           "grad_x": grads[0].numpy()}
    # weights are re-created from the seed by the test (torch.manual_seed(11); default init) and verified
    # against these checksums; only the small tensors are stored to keep the fixture small
    for (k, v) in D.state_dict().items():
        out["sd_abs_sum/" + k] = np.float64(v.double().abs().sum().item())
        if v.numel() <= 4096:
            out["sd_after/" + k] = v.numpy()
    for (k, _), gr in zip(D.named_parameters(), grads[1:]):
        out["grad_abs_sum/" + k] = np.float64(gr.double().abs().sum().item())
        if gr.numel() <= 4096:
            out["grad/" + k] = gr.numpy()
    # weights before the forward == after (no optimiser step); running stats were updated by the train fwd
    D.eval()
    out["y_eval"] = D(x).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "discriminator_small.npz"), **out)
    print("wrote discriminator", out["y_train"].ravel(), out["loss"])


if __name__ == "__main__":
    losses_case("c24_small", B=2, C=24, H=16, W=16, seed=7)
    losses_case("c24_blocky", B=2, C=24, H=24, W=40, seed=21, blocky=True)
    losses_case("c5_ragged", B=1, C=5, H=7, W=9, seed=33, scale=6.0)
    discriminator_case()
