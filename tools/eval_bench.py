#!/usr/bin/env python
"""BASELINE config 5: sliding-window prediction of one 4096x4096 synthetic tile (64 windows of 512x512, 16 per
iteration) with the fused argmax + confusion-matrix kernel.  Prints windows/s, the argmax+histogram kernel's HBM
GB/s, and checks the histogram / mask bit-exactly against torch.argmax + bincount on the same logits."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import uda_aerial_semantic_segmentation_research_b200 as U
from uda_aerial_semantic_segmentation_research_b200 import ops
from uda_aerial_semantic_segmentation_research_b200.predict import sliding_window_evaluate, sliding_window_evaluate_u8, GraphedForward

C, S = 24, int(os.environ.get("TILE", 4096))
dev = torch.device("cuda")
torch.manual_seed(0)
model = U.Unet("resnet34", encoder_weights=None, in_channels=3, classes=C).to(dev).eval()
g = torch.Generator().manual_seed(11)
tile = torch.randn(3, S, S, generator=g).to(dev)
target = torch.randint(0, C, (S // 16, S // 16), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1).to(dev)

def run(return_mask=False):
    return sliding_window_evaluate(model, tile, target, C, window=512, batch=16, return_mask=return_mask)

out = run(True)
torch.cuda.synchronize()
# bit-exact check of argmax + histogram against torch on the same logits (window batch 0)
from uda_aerial_semantic_segmentation_research_b200.predict import tile_windows
with torch.no_grad():
    wins = tile_windows(tile, 512)[:16].contiguous()
    tw = tile_windows(target.reshape(1, S, S), 512).reshape(-1, 512, 512)[:16].contiguous()
    logits = model(wins)
    m_ref = logits.argmax(1)
    h_ref = torch.bincount((C * tw.reshape(-1) + m_ref.reshape(-1)), minlength=C * C).reshape(C, C)
    m, h = ops.argmax_confmat(logits.contiguous(), tw.long(), want_mask=True)
assert torch.equal(m, m_ref) and torch.equal(h, h_ref), "argmax / confusion matrix mismatch"
assert int(out["hist"].sum()) == S * S
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# the argmax+hist kernel alone on one batch of fp32 logits
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(10):
    ops.argmax_confmat(logits, tw.long(), want_mask=False)
f1.record(); torch.cuda.synchronize()
us = f0.elapsed_time(f1) / 10 * 1e3
P = 16 * 512 * 512
nbytes = P * C * 4 + 8 * P
# the raw-tile pipeline: uint8 HWC tile -> gather + normalise kernel -> graph-captured conv-only forward -> argmax + histogram
tile_u8 = torch.randint(0, 256, (S, S, 3), generator=g, dtype=torch.uint8).to(dev)
target_u8 = target.to(torch.uint8)
gf = GraphedForward(model, torch.zeros(16, 3, 512, 512, device=dev))
def run_u8(graphed):
    return sliding_window_evaluate_u8(model, tile_u8, target_u8, C, window=512, batch=16, graphed=graphed)
res = {}
for name, gr in (("eager", None), ("graphed", gf)):
    out_u8 = run_u8(gr); torch.cuda.synchronize()
    assert int(out_u8["hist"].sum()) == S * S
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(reps):
        run_u8(gr)
    g1.record(); torch.cuda.synchronize()
    res[name] = g0.elapsed_time(g1) / reps
h_e = run_u8(None)["hist"]; h_g = run_u8(gf)["hist"]
assert torch.equal(h_e, h_g), "graphed and eager evaluation disagree"
print(json.dumps({"workload": f"sliding-window evaluation of one {S}x{S} tile, {(S // 512) ** 2} windows of 512x512, 16 per iteration, "
                              "U-Net r34 eval-mode forward + fused argmax + confusion matrix (BASELINE configs[4])",
                  "ms_per_tile": ms, "windows_per_s": (S // 512) ** 2 / (ms * 1e-3),
                  "raw_u8_tile_ms_eager": res["eager"], "raw_u8_tile_ms_graphed": res["graphed"],
                  "raw_u8_windows_per_s_graphed": (S // 512) ** 2 / (res["graphed"] * 1e-3), "mean_iou": out["mean_iou"],
                  "argmax_confmat_us_per_16_windows": us, "argmax_confmat_gbs": nbytes / (us * 1e-6) / 1e9,
                  "bit_exact_vs_torch_argmax_bincount": True}))
