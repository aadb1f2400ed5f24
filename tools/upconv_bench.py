#!/usr/bin/env python
"""Decoder conv1 per block at the benchmarked shape: materialised (upsample+concat copy, one 3x3 conv over the
concatenation) vs fused (conv_transpose4x4_s2(x) + conv3x3(skip)) — forward, input gradients, weight gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops

B, S = int(os.environ.get("B", 16)), int(os.environ.get("S", 512))
BLOCKS = [("dec0.c1", S // 32, 512, 256, 256), ("dec1.c1", S // 16, 256, 128, 128), ("dec2.c1", S // 8, 128, 64, 64),
          ("dec3.c1", S // 4, 64, 64, 32), ("dec4.c1", S // 2, 32, 0, 16)]


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, h, C1, C2, O in BLOCKS:
    H = 2 * h
    x = torch.randn(B, h, h, C1, device="cuda").bfloat16()
    skip = torch.randn(B, H, H, C2, device="cuda").bfloat16() if C2 else None
    w = (torch.randn(O, 3, 3, C1 + C2, device="cuda") * 0.05).bfloat16()
    wx, ws = ops.upconv_split_weights(w, C1)
    w4 = ops.weight_flip_transpose(wx)
    wft = ops.weight_flip_transpose(w)
    cat = ops.upcat_fwd(x, skip)
    z = ops.conv_fwd(cat, w, None, 1, 1)
    dz = torch.randn_like(z)
    dw = torch.zeros(O, 3, 3, C1 + C2, device="cuda")
    dw4 = torch.zeros(C1, 4, 4, O, device="cuda")
    dws = torch.zeros(O, 3, 3, C2, device="cuda") if C2 else None
    t = {}
    t["mat fwd"] = timeit(lambda: ops.conv_fwd(ops.upcat_fwd(x, skip), w, None, 1, 1))
    t["fus fwd"] = timeit(lambda: ops.conv_fwd_add(skip, ws, ops.upconv_fwd(x, wx)) if C2 else ops.upconv_fwd(x, wx))
    t["  upconv only"] = timeit(lambda: ops.upconv_fwd(x, wx))
    t["mat dgrad"] = timeit(lambda: ops.upcat_bwd(ops.conv_dgrad(dz, w, cat.shape, 1, 1, w_ft=wft), C1, C2))
    t["fus dgrad"] = timeit(lambda: (ops.conv_fwd(dz, w4, None, 2, 1), ops.conv_dgrad(dz, ws, skip.shape, 1, 1) if C2 else None))
    t["  dx only"] = timeit(lambda: ops.conv_fwd(dz, w4, None, 2, 1))
    t["mat wgrad"] = timeit(lambda: ops.conv_wgrad(dz, cat, dw, 1, 1))
    t["fus wgrad"] = timeit(lambda: (ops.conv_wgrad(x, dz, dw4, 2, 1), ops.conv_wgrad(dz, skip, dws, 1, 1) if C2 else None,
                                     ops.upconv_merge_wgrad(dw4, dws, dw, C1)))
    t["  dw4 only"] = timeit(lambda: ops.conv_wgrad(x, dz, dw4, 2, 1))
    print(f"{name} C1={C1} C2={C2} O={O} lowres {h}: " + "  ".join(f"{k} {v:6.1f}us" for k, v in t.items()))
