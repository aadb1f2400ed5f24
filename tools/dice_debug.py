"""Debug: per-(b,c) Dice sums of the streaming pass-1 kernel vs torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import ops
torch.manual_seed(0)
B, C, H, W = 3, 24, 128, 128
z = torch.randn(B, C, H, W, device="cuda") * 3
t = torch.randint(0, C, (B, H, W), device="cuda")
out4, grad = ops.seg_loss(z, t, ce_mode=0, use_dice=True, smooth=0.5)
torch.cuda.synchronize()
ws = ops.workspace(1, z.device)
sums = ws[32:32 + B * C * 3 * 8].view(torch.float64).view(B, C, 3).cpu()
p = torch.softmax(z.double(), 1)
oh = torch.nn.functional.one_hot(t, C).permute(0, 3, 1, 2).double()
ref = torch.stack([(p * oh).sum((2, 3)), p.sum((2, 3)), oh.sum((2, 3))], -1).cpu()
d = (sums - ref)
print("max abs diff I, sum p, sum t:", d.abs().amax((0, 1)))
print("rel diff per image/class (I):", (d[..., 0] / ref[..., 0])[0])
print("rel diff (sum p):", (d[..., 1] / ref[..., 1])[0])
print("diff (sum t):", d[..., 2][0])
