#!/usr/bin/env python
"""ncu target: ONE forward / dgrad / wgrad launch of selected U-Net r34 convolution shapes at the benchmarked size
(B=16, 512x512 input), so that `ncu --set full --import-source on -k regex:conv_tc` captures exactly the kernels the
training step runs.  Usage: python tools/ncu_conv.py [name-substring ...]   (default: the dominant shapes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops

B = int(os.environ.get("B", 16)); S = int(os.environ.get("S", 512))
s2, s4, s8, s16, s32 = S // 2, S // 4, S // 8, S // 16, S // 32
SHAPES = [("layer1", s4, 64, 64, 3, 1), ("l2.0", s4, 64, 128, 3, 2), ("layer2", s8, 128, 128, 3, 1),
          ("l3.0", s8, 128, 256, 3, 2), ("layer3", s16, 256, 256, 3, 1), ("layer4", s32, 512, 512, 3, 1),
          ("dec0.c1", s16, 768, 256, 3, 1), ("dec1.c1", s8, 384, 128, 3, 1), ("dec2.c1", s4, 192, 64, 3, 1),
          ("dec3.c1", s2, 128, 32, 3, 1), ("dec4.c2", S, 16, 16, 3, 1), ("head", S, 16, 24, 3, 1)]
want = sys.argv[1:] or ["layer1", "layer2", "layer3", "layer4", "dec0.c1", "l3.0", "dec3.c1"]
for name, H, Cin, Cout, k, s in SHAPES:
    if not any(w == name for w in want):
        continue
    p = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda") * 0.05).bfloat16()
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv_fwd(x, w, None, s, p, bn_sums=sums)
    dy = torch.randn_like(y)
    wft = ops.weight_flip_transpose(w)
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push(name)
    ops.conv_fwd(x, w, None, s, p, bn_sums=sums)
    ops.conv_dgrad(dy, w, x.shape, s, p, w_ft=wft)
    ops.conv_wgrad(dy, x, dw, s, p)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("ran", name, flush=True)
