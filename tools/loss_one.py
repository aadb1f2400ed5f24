"""One CE fwd+bwd (fp32 NCHW, B=16, 24 classes, 512x512) a few times: the ncu target for the loss kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops
z = torch.randn(16, 24, 512, 512, device="cuda"); t = torch.randint(0, 24, (16, 512, 512), device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "ce"
for _ in range(4):
    if mode == "ce": ops.seg_loss(z, t, ce_mode=1, use_dice=False)
    elif mode == "cedice": ops.seg_loss(z, t, ce_mode=1, use_dice=True)
torch.cuda.synchronize()
print("ok")
