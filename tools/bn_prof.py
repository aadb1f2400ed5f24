import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops
dev = "cuda"
B, H, C = 8, 256, 64
x = torch.randn(B, H, H, C, device=dev).bfloat16(); dy = torch.randn_like(x); res = torch.randn_like(x)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
for _ in range(2):
    mean, rstd, sc, sh = ops.bn_stats(x, g, b, rm, rv)
    a = ops.bn_apply(x, sc, sh, None, 0.0)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    ops.bn_bwd(dy, x, None, g, mean, rstd, 0.0, dg, db, scale=sc, shift=sh)
    ops.bn_bwd(dy, x, a, g, mean, rstd, 0.0, dg, db, dres=torch.empty_like(x))
    ops.upcat_bwd(x, 64, 0)
    ops.maxpool_fwd(x)
torch.cuda.synchronize()
print("ok")
