#!/usr/bin/env python
"""Per-CTA phase / wait-time trace of the tcgen05 convolution kernels (experiment build: `make EXPERIMENTS=1`).
For each selected U-Net r34 shape at the benchmarked size it runs forward (with the fused BatchNorm statistics) and
dgrad once with tracing on and prints, averaged over CTAs (and for the slowest CTA), where each warp role spent its
time: TMA producer waiting for free ring slots, MMA thread waiting for operands / for a drained accumulator, epilogue
waiting for accumulators vs busy.  Clocks are SM clocks (1.965 GHz)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libuda_b200_exp.so")
from uda_aerial_semantic_segmentation_research_b200 import ops

B = int(os.environ.get("B", 16)); S = int(os.environ.get("S", 512))
s2, s4, s8, s16, s32 = S // 2, S // 4, S // 8, S // 16, S // 32
SHAPES = [("layer1", s4, 64, 64, 3, 1), ("l2.0", s4, 64, 128, 3, 2), ("layer2", s8, 128, 128, 3, 1),
          ("l3.0", s8, 128, 256, 3, 2), ("layer3", s16, 256, 256, 3, 1), ("layer4", s32, 512, 512, 3, 1),
          ("dec0.c1", s16, 768, 256, 3, 1), ("dec1.c1", s8, 384, 128, 3, 1), ("dec2.c1", s4, 192, 64, 3, 1),
          ("dec3.c1", s2, 128, 32, 3, 1), ("dec4.c2", S, 16, 16, 3, 1), ("head", S, 16, 24, 3, 1)]
want = sys.argv[1:] or ["layer1", "layer2", "layer3", "layer4", "dec0.c1", "l3.0", "dec3.c1"]
trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib = _lib.lib()
NAMES = ["entry", "pdl_wait", "prod_wait_empty", "prod_done", "mma_wait_full", "mma_wait_tempty", "mma_first", "mma_last_commit",
         "epi_wait_tfull", "epi_busy", "epi_done", "cta_exit", "tiles", "mma_wait_afull"]


def run(tag, fn):
    fn(); torch.cuda.synchronize()       # warm-up without tracing
    trace.zero_()
    lib.uda_exp_set_trace(ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.uda_exp_set_trace(ctypes.c_void_p(0))
    t = trace.view(148, 16).cpu()
    used = t[:, 11] > 0
    if not bool(used.any()):
        print(f"{tag:22s} {e0.elapsed_time(e1)*1e3:7.1f} us  (kernel not traced)")
        return
    t = t[used].double()
    slow = int(t[:, 11].argmax())
    def f(i):
        return f"{t[:, i].mean()/1e3:6.1f}k/{t[slow, i]/1e3:6.1f}k"
    print(f"{tag:22s} {e0.elapsed_time(e1)*1e3:7.1f} us  ctas {int(used.sum()):3d} tiles/cta {t[:,12].mean():4.1f}  [mean/slowest CTA, kclk]  "
          f"exit {f(11)}  pdl_wait {f(1)}  first_mma {f(6)}  last_commit {f(7)}  epi_done {f(10)} | "
          f"producer wait {f(2)}  mma wait operands {f(4)} (+A {f(13)})  mma wait epilogue {f(5)} | epi wait {f(8)} busy {f(9)}")


for name, H, Cin, Cout, k, s in SHAPES:
    if name not in want:
        continue
    p = (k - 1) // 2
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda") * 0.05).bfloat16()
    sums = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv_fwd(x, w, None, s, p)
    dy = torch.randn_like(y)
    wft = ops.weight_flip_transpose(w)
    run(f"{name} fwd+stats", lambda: ops.conv_fwd(x, w, None, s, p, bn_sums=sums))
    run(f"{name} fwd", lambda: ops.conv_fwd(x, w, None, s, p))
    run(f"{name} dgrad", lambda: ops.conv_dgrad(dy, w, x.shape, s, p, w_ft=wft))
    dw = torch.zeros(Cout, k, k, Cin, device="cuda")
    run(f"{name} wgrad", lambda: ops.conv_wgrad(dy, x, dw, s, p))
