#!/usr/bin/env python
"""Per-launch, per-CTA role traces of EVERY tcgen05 convolution launch INSIDE the captured training step (experiment
build: `make EXPERIMENTS=1`), on one time axis (%globaltimer at CTA entry + SM clocks).

trace_conv.py times single launches in isolation; the in-graph CUPTI timeline (tools/timeline.py) shows kernels of
15-18 us in-kernel time occupying 25-30 us of the main stream.  This tool shows where the difference goes: for each conv
launch of the replayed graph it prints, relative to the moment the previous traced launch's last CTA exited,

  entry    first / last CTA entry (programmatic dependent launch: usually before the predecessor ended)
  go       griddepcontrol.wait returned (predecessor complete + flushed)   [max over CTAs]
  mma0     first MMA issued            [mean]
  mmaN     last commit issued          [max]
  epi      epilogue done               [max]
  exit     last CTA exit
  and the in-CTA waits (MMA thread waiting for operands / epilogue, producer waiting for slots).
"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uda_aerial_semantic_segmentation_research_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libuda_b200_exp.so")
import uda_aerial_semantic_segmentation_research_b200 as U
from uda_aerial_semantic_segmentation_research_b200 import ops
from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
from uda_aerial_semantic_segmentation_research_b200.graph import GraphedStep

B = int(os.environ.get("B", 16)); S = int(os.environ.get("S", 512))
GHZ = float(os.environ.get("SM_GHZ", 1.965))
NS = 1200
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = _lib.lib()
buf = torch.zeros(NS * 148 * 16, dtype=torch.int64, device=dev)
model = U.Unet("resnet34", encoder_weights=None, in_channels=3, classes=24).to(dev).train()
opt = FusedAdam(model, lr=1e-3)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 3, S, S, generator=g).to(dev)
t = torch.randint(0, 24, (B, S, S), generator=g).to(dev)
lib.uda_exp_set_trace_series(ctypes.c_void_p(buf.data_ptr()), NS)
step = GraphedStep(model, CrossEntropyLoss(), opt, x, t)
left = lib.uda_exp_set_trace_series(ctypes.c_void_p(0), 0)
used = NS - left
for _ in range(3):
    step(x, t)
torch.cuda.synchronize()
buf.zero_()
step(x, t)
torch.cuda.synchronize()
tr = buf.view(NS, 148, 16)[:used].cpu()
live = [i for i in range(used) if int((tr[i, :, 11] > 0).sum()) > 0]
print(f"traced launches issued {used}, in the captured graph {len(live)}; B={B} S={S}, clk = {GHZ} GHz")
KIND = {1: "persist", 2: "halo", 3: "phalo", 4: "wgrad_big", 17: "persist+bn", 19: "phalo+bn"}
rows = []
for i in live:
    r = tr[i]
    m = r[:, 11] > 0
    r = r[m].double()
    meta = int(r[0, 15])
    kind, cout, cred, tiles = meta & 0xff, (meta >> 8) & 0xffff, (meta >> 24) & 0xffff, meta >> 40
    g0 = r[:, 14]                                  # ns
    ns = lambda col: g0 + r[:, col] / GHZ          # absolute ns of a per-CTA clock offset
    rows.append(dict(i=i, kind=KIND.get(kind, "?"), cout=cout, cred=cred, tiles=tiles, ctas=int(m.sum()),
                     entry0=float(g0.min()), entry1=float(g0.max()), go=float(ns(1).max()), mma0=float(ns(6).mean()),
                     mmaN=float(ns(7).max()), epi=float(ns(10).max()), exit=float(ns(11).max()),
                     w_ops=float(r[:, 4].mean()), w_epi=float(r[:, 5].mean()), w_prod=float(r[:, 2].mean()),
                     epi_busy=float(r[:, 9].mean()), in_kernel=float((r[:, 11] - r[:, 1]).max()),
                     p1=float(ns(5).max()), bar=float(ns(13).max()), raw=r[0].tolist()))
rows.sort(key=lambda d: d["entry0"])
t_base = rows[0]["entry0"]
print("  #   t_us kind      Cred Cout tiles ctas | rel. to prev traced exit (us): entry0 entryN    go  mma0  mmaN   epi  exit |"
      " go->exit us | in-CTA kclk: wait_ops wait_epi prod_wait epi_busy")
prev_exit = {}
for d in rows:
    stream = "w" if d["kind"] == "wgrad_big" else "m"
    pe = prev_exit.get(stream, d["entry0"])
    f = lambda v: f"{(v - pe) / 1e3:6.1f}"
    print(f"{d['i']:4d} {(d['entry0'] - t_base) / 1e3:7.1f} {d['kind']:9s} {d['cred']:4d} {d['cout']:4d} {d['tiles']:5d} {d['ctas']:4d} | "
          f"{f(d['entry0'])} {f(d['entry1'])} {f(d['go'])} {f(d['mma0'])} {f(d['mmaN'])} {f(d['epi'])} {f(d['exit'])} | "
          f"{(d['exit'] - d['go']) / 1e3:6.1f} | {d['w_ops'] / 1e3:6.1f} {d['w_epi'] / 1e3:6.1f} {d['w_prod'] / 1e3:6.1f} {d['epi_busy'] / 1e3:6.1f}")
    if d["kind"].endswith("+bn"):     # fused conv + BatchNorm: pass 1 done / barrier passed / pass 2 done, relative to go (us)
        print(f"       fused epilogue: pass1 done {(d['p1'] - d['go']) / 1e3:5.1f}  barrier passed {(d['bar'] - d['go']) / 1e3:5.1f}  "
              f"pass2 done {(d['epi'] - d['go']) / 1e3:5.1f}  (last MMA {(d['mmaN'] - d['go']) / 1e3:5.1f})")
    prev_exit[stream] = d["exit"]
mid = rows[len(rows) // 4]
print("raw record of CTA 0 of launch", mid["i"], mid["kind"], [int(v) for v in mid["raw"]])
# summary per kind: time from go to exit (in-kernel useful window) vs first-mma delay and tail
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for d in rows:
    a = agg[d["kind"]]
    a[0] += 1; a[1] += d["exit"] - d["go"]; a[2] += d["mma0"] - d["go"]; a[3] += d["mmaN"] - d["mma0"]; a[4] += d["exit"] - d["mmaN"]
print("kind, launches, sum go->exit us, sum go->first MMA us, sum first->last MMA us, sum last MMA->exit us")
for k, a in agg.items():
    print(f"{k},{a[0]},{a[1] / 1e3:.1f},{a[2] / 1e3:.1f},{a[3] / 1e3:.1f},{a[4] / 1e3:.1f}")
