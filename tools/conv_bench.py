#!/usr/bin/env python
"""Per-shape micro-benchmark of the convolution kernels (CUDA events, L2 flushed between repetitions).
Prints one line per (shape, op): time, algorithmic TFLOP/s, effective HBM GB/s of the compulsory traffic."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops

B = int(os.environ.get("B", 16)); S = int(os.environ.get("S", 512))
# (name, H, Cin, Cout, k, stride, count) of U-Net r34 at input S
def shapes(S):
    s2, s4, s8, s16, s32 = S // 2, S // 4, S // 8, S // 16, S // 32
    return [("layer1 64->64", s4, 64, 64, 3, 1, 6), ("l2.0 64->128 s2", s4, 64, 128, 3, 2, 1),
            ("l2 ds 1x1 s2", s4, 64, 128, 1, 2, 1), ("layer2 128->128", s8, 128, 128, 3, 1, 7),
            ("l3.0 128->256 s2", s8, 128, 256, 3, 2, 1), ("layer3 256->256", s16, 256, 256, 3, 1, 11),
            ("l4.0 256->512 s2", s16, 256, 512, 3, 2, 1), ("layer4 512->512", s32, 512, 512, 3, 1, 5),
            ("dec0.c1 768->256", s16, 768, 256, 3, 1, 1), ("dec0.c2 256->256", s16, 256, 256, 3, 1, 1),
            ("dec1.c1 384->128", s8, 384, 128, 3, 1, 1), ("dec1.c2 128->128", s8, 128, 128, 3, 1, 1),
            ("dec2.c1 192->64", s4, 192, 64, 3, 1, 1), ("dec2.c2 64->64", s4, 64, 64, 3, 1, 1),
            ("dec3.c1 128->32", s2, 128, 32, 3, 1, 1), ("dec3.c2 32->32", s2, 32, 32, 3, 1, 1),
            ("dec4.c1 32->16", S, 32, 16, 3, 1, 1), ("dec4.c2 16->16", S, 16, 16, 3, 1, 1),
            ("head 16->24", S, 16, 24, 3, 1, 1)]

def timeit(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.max()   # read-only L2 flush
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print(f"B={B} S={S}")
for name, H, Cin, Cout, k, s, cnt in shapes(S):
    p = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda") * 0.05).bfloat16()
    y = ops.conv_fwd(x, w, None, s, p)
    dy = torch.randn_like(y)
    wft = ops.weight_flip_transpose(w)
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    gf = 2.0 * y.numel() * Cin * k * k / 1e9
    bytes_fwd = (x.numel() + y.numel() + w.numel()) * 2
    t_f = timeit(lambda: ops.conv_fwd(x, w, None, s, p))
    t_d = timeit(lambda: ops.conv_dgrad(dy, w, x.shape, s, p, w_ft=wft))
    t_w = timeit(lambda: ops.conv_wgrad(dy, x, dw, s, p))
    tot["fwd"] += cnt * t_f; tot["dgrad"] += cnt * t_d; tot["wgrad"] += cnt * t_w
    print(f"{name:18s} x{cnt:2d} {gf:7.1f} GF | fwd {t_f*1e3:7.1f} us {gf/t_f:7.1f} TF/s {bytes_fwd/t_f/1e6:7.0f} GB/s"
          f" | dgrad {t_d*1e3:7.1f} us {gf/t_d:7.1f} TF/s | wgrad {t_w*1e3:7.1f} us {gf/t_w:7.1f} TF/s")
print("weighted totals per step (ms):", {k: round(v, 3) for k, v in tot.items()})
