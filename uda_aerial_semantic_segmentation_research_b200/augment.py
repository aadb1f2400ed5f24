"""Device-side strong augmentation for the unsupervised fine-tuning step.

The reference builds two strongly augmented views of every target batch on the host, image by image
(``src/models/unsupervised_trainer.py:99-114``: tensor -> numpy -> ``get_strong_augmentation()`` -> tensor -> device; the
pipeline is ``src/models/augmentation.py:40-80``).  ``StrongAugmentation`` keeps the pipeline's random DECISIONS on the
host — they are a dozen numbers per image, drawn with the pipeline's probabilities and limits — and does the pixel work
on the device in one gather kernel per view (``uda_strong_augment``):

    RandomRotate90(p=0.7), Flip(p=0.7), Transpose(p=0.7)   -> one element of the dihedral group D4
    ShiftScaleRotate(shift 0.1, scale 0.3, rotate 60, p=0.5) -> a similarity transform about the image centre
    GaussNoise(var 20..80 on the 0..255 scale, OneOf p=0.4)  -> additive N(0, sigma^2), counter-based generator
    RandomBrightnessContrast(0.3, 0.3; OneOf p=0.5 x 0.4/1.6) -> value * alpha + beta

composed into ONE inverse 2x3 map per image (bilinear, reflect-101 border = albumentations' default).  The blur family,
optical / grid / elastic distortion, CLAHE / Sharpen / Emboss and HueSaturationValue are not covered.
albumentations is a third-party dependency that is absent here and draws from its own global RNG: the augmentation is
random by construction, so parity is checked for the KERNEL (against the numpy restatement ``oracle/ref_augment.py``
on the same parameter table), not for a random stream.
"""
import math

import numpy as np
import torch

from . import ops


def _d4_matrix(k, hflip, vflip, transpose):
    """2x2 integer matrix (acting on centred coordinates) of: rot90 k times, then flips, then transpose."""
    m = np.eye(2)
    rot = np.array([[0.0, 1.0], [-1.0, 0.0]])      # image coordinates (y down): k = 1 equals np.rot90(img, 1)
    for _ in range(k % 4):
        m = rot @ m
    if hflip:
        m = np.array([[-1.0, 0.0], [0.0, 1.0]]) @ m
    if vflip:
        m = np.array([[1.0, 0.0], [0.0, -1.0]]) @ m
    if transpose:
        m = np.array([[0.0, 1.0], [1.0, 0.0]]) @ m
    return m


def build_table(params, H, W):
    """Parameter rows for ``uda_strong_augment`` from per-image decisions.

    ``params``: list of dicts with keys k (rot90 count), hflip, vflip, transpose, angle (degrees), scale, dx, dy
    (shift as a fraction of width / height), alpha, beta, sigma, seed.  The forward map is
    ``p' = c + S(angle, scale) * D4 * (p - c) + (dx W, dy H)`` with ``c`` the image centre; the table holds its inverse
    (output pixel -> source coordinates).  Square images only when the D4 element swaps the axes."""
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    rows = np.zeros((len(params), 12), dtype=np.float32)
    for i, p in enumerate(params):
        d4 = _d4_matrix(p.get("k", 0), p.get("hflip", False), p.get("vflip", False), p.get("transpose", False))
        if H != W and abs(d4[0, 0]) < 0.5:
            raise ValueError("axis-swapping symmetries need square images")
        a = math.radians(p.get("angle", 0.0))
        s = p.get("scale", 1.0)
        sim = s * np.array([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
        fwd = sim @ d4
        inv = np.linalg.inv(fwd)
        t = np.array([p.get("dx", 0.0) * W, p.get("dy", 0.0) * H])
        c = np.array([cx, cy])
        off = c - inv @ (c + t)                     # src = inv (dst - c - t) + c
        rows[i, 0:3] = (inv[0, 0], inv[0, 1], off[0])
        rows[i, 3:6] = (inv[1, 0], inv[1, 1], off[1])
        rows[i, 6:10] = (p.get("alpha", 1.0), p.get("beta", 0.0), p.get("sigma", 0.0), float(p.get("seed", 0)))
    return rows


class StrongAugmentation:
    """``view = StrongAugmentation(seed)(images)`` — one strongly augmented view of a CUDA fp32 NCHW batch.

    ``value_scale``: size of one 8-bit grey level in the images' units (1/255 for images in [0,1], 1/(255*std) after
    Normalize); the noise variances and brightness shifts of the pipeline are given on the 0..255 scale."""

    def __init__(self, seed=0, value_scale=1.0 / 255.0):
        self.rng = np.random.default_rng(seed)
        self.value_scale = value_scale
        self._n = 0

    def sample(self, B):
        """Draw the pipeline's decisions for ``B`` images (probabilities / limits of augmentation.py:40-80)."""
        r = self.rng
        out = []
        for _ in range(B):
            p = {"k": int(r.integers(0, 4)) if r.random() < 0.7 else 0,
                 "hflip": False, "vflip": False, "transpose": bool(r.random() < 0.7)}
            if r.random() < 0.7:                       # A.Flip: d in {-1, 0, 1} = both / vertical / horizontal
                d = int(r.integers(-1, 2))
                p["hflip"], p["vflip"] = d in (-1, 1), d in (-1, 0)
            if r.random() < 0.5:                       # ShiftScaleRotate
                p.update(angle=float(r.uniform(-60, 60)), scale=float(1.0 + r.uniform(-0.3, 0.3)),
                         dx=float(r.uniform(-0.1, 0.1)), dy=float(r.uniform(-0.1, 0.1)))
            if r.random() < 0.4:                       # OneOf(GaussNoise(30..80), GaussNoise(20..60))
                lo, hi = ((30.0, 80.0), (20.0, 60.0))[int(r.integers(0, 2))]
                p["sigma"] = math.sqrt(float(r.uniform(lo, hi))) * 255.0 * self.value_scale
            if r.random() < 0.5 and r.random() < 0.25:  # OneOf(...) p=0.5, RandomBrightnessContrast weight 0.4 of 1.6
                p["alpha"] = float(1.0 + r.uniform(-0.3, 0.3))
                p["beta"] = float(r.uniform(-0.3, 0.3)) * 255.0 * self.value_scale
            self._n += 1
            p["seed"] = self._n * 7919 % (1 << 23)
            out.append(p)
        return out

    def __call__(self, images, params=None):
        if not (images.is_cuda and images.dtype == torch.float32 and images.dim() == 4):
            raise RuntimeError("StrongAugmentation: CUDA float32 NCHW batches only (no CPU fallback)")
        B, C, H, W = images.shape
        params = self.sample(B) if params is None else params
        table = torch.from_numpy(build_table(params, H, W)).pin_memory().to(images.device, non_blocking=True)
        return ops.strong_augment(images.contiguous(), table)
