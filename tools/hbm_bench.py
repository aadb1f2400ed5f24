#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound kernels at the cfg2 shapes (CUDA events, L2 flushed): time and
effective GB/s of the algorithmic (compulsory) bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops

B = int(os.environ.get("B", 16)); S = int(os.environ.get("S", 512))
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.max()   # read-only L2 flush: leaves no dirty lines behind
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]

def line(name, ms, nbytes):
    print(f"{name:34s} {ms*1e3:9.1f} us {nbytes/ms/1e6:8.0f} GB/s  ({nbytes/1e6:8.1f} MB)")

print(f"B={B} S={S}")
for C, H in ((16, S), (64, S // 4), (64, S // 2), (128, S // 8), (512, S // 32)):
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    res = torch.randn_like(x); dy = torch.randn_like(x)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    n = x.numel() * 2
    mean, rstd, sc, sh = ops.bn_stats(x, g, b, rm, rv)
    line(f"bn_stats   C={C:3d} {H}x{H}", timeit(lambda: ops.bn_stats(x, g, b, rm, rv)), n)
    a = ops.bn_apply(x, sc, sh, None, 0.0)
    line(f"bn_apply   C={C:3d} {H}x{H}", timeit(lambda: ops.bn_apply(x, sc, sh, None, 0.0)), 2 * n)
    line(f"bn_apply+r C={C:3d} {H}x{H}", timeit(lambda: ops.bn_apply(x, sc, sh, res, 0.0)), 3 * n)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    line(f"bn_bwd(a)  C={C:3d} {H}x{H}", timeit(lambda: ops.bn_bwd(dy, x, a, g, mean, rstd, 0.0, dg, db)), 7 * n)
    line(f"bn_bwd(z)  C={C:3d} {H}x{H}", timeit(lambda: ops.bn_bwd(dy, x, None, g, mean, rstd, 0.0, dg, db, scale=sc, shift=sh)), 5 * n)
    del x, res, dy, a
lo = torch.randn(B, S // 2, S // 2, 32, device=dev).bfloat16()
up = ops.upcat_fwd(lo, None)
line("upcat_fwd 32ch 256->512", timeit(lambda: ops.upcat_fwd(lo, None)), lo.numel() * 2 + up.numel() * 2)
line("upcat_bwd 32ch 512->256", timeit(lambda: ops.upcat_bwd(up, 32, 0)), lo.numel() * 2 + up.numel() * 2)
del lo, up
lo, sk = torch.randn(B, S // 4, S // 4, 64, device=dev).bfloat16(), torch.randn(B, S // 2, S // 2, 64, device=dev).bfloat16()
up = ops.upcat_fwd(lo, sk)
line("upcat_fwd 64+64 128->256", timeit(lambda: ops.upcat_fwd(lo, sk)), (lo.numel() + sk.numel() + up.numel()) * 2)
del lo, sk, up
x = torch.relu(torch.randn(B, S // 2, S // 2, 64, device=dev)).bfloat16()
y, idx = ops.maxpool_fwd(x)
line("maxpool_fwd 64ch 256->128", timeit(lambda: ops.maxpool_fwd(x)), x.numel() * 2 + y.numel() * 3)
line("maxpool_bwd 64ch", timeit(lambda: ops.maxpool_bwd(y, idx, x.shape)), x.numel() * 2 + y.numel() * 3)
del x, y, idx
z = torch.randn(B, 24, S, S, device=dev); t = torch.randint(0, 24, (B, S, S), device=dev)
P = B * S * S
line("CE fwd+bwd fp32", timeit(lambda: ops.seg_loss(z, t, ce_mode=1, use_dice=False)), 2 * P * 24 * 4 + 8 * P)
line("CE+Dice fwd+bwd fp32", timeit(lambda: ops.seg_loss(z, t, ce_mode=1, use_dice=True)), 3 * P * 24 * 4 + 16 * P)
z2 = torch.randn_like(z)
line("consistency fp32", timeit(lambda: ops.consistency(z, z2)), 4 * P * 24 * 4)
del z2
line("entropy fp32", timeit(lambda: ops.entropy(z)), 2 * P * 24 * 4)
line("argmax+confmat fp32 (i64 mask)", timeit(lambda: ops.argmax_confmat(z, t)), P * 24 * 4 + 16 * P)
line("argmax+confmat fp32 (no mask)", timeit(lambda: ops.argmax_confmat(z, t, want_mask=False)), P * 24 * 4 + 8 * P)
line("nchw->nhwc dlogits", timeit(lambda: ops.nchw_to_nhwc(z, torch.bfloat16)), P * 24 * 6)
zb = z.bfloat16()
line("CE fwd+bwd bf16", timeit(lambda: ops.seg_loss(zb, t, ce_mode=1, use_dice=False)), 2 * P * 24 * 2 + 8 * P)
line("CE+Dice fwd+bwd bf16", timeit(lambda: ops.seg_loss(zb, t, ce_mode=1, use_dice=True)), 3 * P * 24 * 2 + 16 * P)
zb2 = torch.randn_like(zb)
line("consistency bf16", timeit(lambda: ops.consistency(zb, zb2)), 4 * P * 24 * 2)
del zb2
line("argmax+confmat bf16 (u8 mask)", timeit(lambda: ops.argmax_confmat(zb, t, mask_dtype=torch.uint8)), P * 24 * 2 + 9 * P)
