"""Oracle restatement of the device-side strong augmentation — TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy restatement of ``uda_strong_augment``'s pixel math for a given parameter table: inverse 2x3 map, bilinear sampling
with OpenCV's BORDER_REFLECT_101 (albumentations' default border mode for its geometric transforms), value * alpha +
beta.  The reference's own pipeline (``src/models/augmentation.py:40-80``) is albumentations — third-party, absent, and
random — so the pin is on the transform definitions: RandomRotate90 / Flip / Transpose must reproduce numpy's
rot90 / flip / transpose exactly (tests/test_gpu_layers.py::test_strong_augmentation_kernel), and a similarity
transform must map the known control points.  The additive noise is checked statistically."""
import numpy as np


def _reflect101(i, n):
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i < n, i, period - i)


def strong_augment(images, table):
    """images [B,C,H,W] float, table [B,12] -> augmented batch without the noise term (float64 arithmetic)."""
    x = np.asarray(images, dtype=np.float64)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    for b in range(B):
        m = np.asarray(table[b], dtype=np.float64)
        sx = np.float32(m[0]) * xs + np.float32(m[1]) * ys + np.float32(m[2])
        sy = np.float32(m[3]) * xs + np.float32(m[4]) * ys + np.float32(m[5])
        fx, fy = np.floor(sx), np.floor(sy)
        ax, ay = sx - fx, sy - fy
        x0, x1 = _reflect101(fx.astype(np.int64), W), _reflect101(fx.astype(np.int64) + 1, W)
        y0, y1 = _reflect101(fy.astype(np.int64), H), _reflect101(fy.astype(np.int64) + 1, H)
        for c in range(C):
            p = x[b, c]
            v = (p[y0, x0] * (1 - ax) + p[y0, x1] * ax) * (1 - ay) + (p[y1, x0] * (1 - ax) + p[y1, x1] * ax) * ay
            out[b, c] = v * m[6] + m[7]
    return out
