"""Times uda_bn_apply_maxpool_fused alone at the stem's benchmarked shape (B=16, 256x256x64), L2 flushed between calls."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uda_aerial_semantic_segmentation_research_b200 import ops
dev = "cuda:0"
z = torch.randn(16, 256, 256, 64, device=dev).bfloat16()
zf = z.double()
sums = torch.cat([zf.sum((0, 1, 2)), (zf ** 2).sum((0, 1, 2))]).contiguous()
g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.bn_apply_maxpool_fused(z, sums, g, b, rm, rv); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
med = sorted(ts)[len(ts) // 2]
# algorithmic bytes: z read, a written, pooled y (1/4) written, winning taps (1 byte per pooled element)
print(f"bn_apply_maxpool_fused: {med:.1f} us (min {min(ts):.1f}) = {(2.25 * z.numel() * 2 + z.numel() // 4) / med / 1e6:.2f} TB/s")
