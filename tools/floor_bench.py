import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops
dev = "cuda"
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for C, H, B in ((512, 16, 16), (256, 32, 16), (128, 64, 16), (64, 128, 16)):
    x = torch.randn(B, H, H, C, device=dev).bfloat16(); dy = torch.randn_like(x)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    mean, rstd, sc, sh = ops.bn_stats(x, g, b, rm, rv)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    xf = x.double().reshape(-1, C); sums[:C] = xf.sum(0); sums[C:] = (xf * xf).sum(0)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    a = ops.bn_apply(x, sc, sh, None, 0.0)
    print(f"C={C} {H}x{H} ({x.numel()*2/1e6:.1f} MB): apply {timeit(lambda: ops.bn_apply(x, sc, sh, None, 0.0)):.1f} us | "
          f"apply_fused {timeit(lambda: ops.bn_apply_fused(x, sums, g, b, rm, rv)):.1f} us | "
          f"bwd(z) {timeit(lambda: ops.bn_bwd(dy, x, None, g, mean, rstd, 0.0, dg, db, scale=sc, shift=sh)):.1f} us | "
          f"bwd(a) {timeit(lambda: ops.bn_bwd(dy, x, a, g, mean, rstd, 0.0, dg, db)):.1f} us | "
          f"empty_like {timeit(lambda: torch.empty_like(x)):.1f} us | zeros64 {timeit(lambda: torch.zeros(2*C, dtype=torch.float64, device=dev)):.1f} us")
os.environ["X"] = "1"
