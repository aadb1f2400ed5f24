"""Kernel timeline of the captured training step (CUPTI through torch.profiler — the image has no nsys).

Runs the supervised bench step (U-Net r34, B=16 @512x512, CE + Adam) as the CUDA-graph replay bench.py times and records
every kernel of a few replays with its start, duration and stream.  Output (stdout, also --out FILE):

* per-kernel-name totals IN the graph (launch count, total / mean duration) — unlike the eager CUDA-event breakdown of
  bench.py these contain no launch gaps, and unlike the ncu launch list they are warm and overlapped as in the timed run;
* the step's wall time, the time at least one kernel is running, the time two streams overlap, and the idle gaps
  (histogram + the 12 largest with the kernels on either side);
* optionally (--dump) the full ordered list.

Numbers taken under a profiler are evidence of STRUCTURE (shares, gaps, overlap), never a bench value.
"""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--dump", action="store_true")
    args = ap.parse_args()
    import torch
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    from uda_aerial_semantic_segmentation_research_b200.graph import GraphedStep
    from torch.profiler import profile, ProfilerActivity

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = U.Unet("resnet34", encoder_weights=None, in_channels=3, classes=24).to(dev).train()
    opt = FusedAdam(model, lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(args.batch, 3, args.size, args.size, generator=g).to(dev)
    t = torch.randint(0, 24, (args.batch, args.size, args.size), generator=g).to(dev)
    step = GraphedStep(model, CrossEntropyLoss(), opt, x, t)
    for _ in range(5):
        step(x, t)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step(x, t)
        torch.cuda.synchronize()
    path = "/tmp/uda_timeline_trace.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X"]
    ks.sort(key=lambda e: e["ts"])
    if not ks:
        print("no kernel records (CUPTI unavailable?)")
        return 1
    # split into steps at the input copies (two memcpy DtoD per step precede the graph)
    lines = []
    P = lines.append
    # use the LAST step only for the structural numbers
    adam = [i for i, e in enumerate(ks) if "adam" in e["name"]]
    if len(adam) >= 2:
        lo, hi = adam[-2] + 1, adam[-1] + 1
    else:
        lo, hi = 0, len(ks)
    last = ks[lo:hi]
    t0 = last[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in last)
    P(f"last step: {len(last)} device records, wall {t1 - t0:.1f} us")
    # union / overlap
    pts = []
    for e in last:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    busy = over = 0.0
    depth, prev = 0, pts[0][0]
    for tt, d in pts:
        if depth >= 1:
            busy += tt - prev
        if depth >= 2:
            over += tt - prev
        depth += d
        prev = tt
    P(f"  >=1 kernel running {busy:.1f} us, >=2 concurrent {over:.1f} us, idle {t1 - t0 - busy:.1f} us")
    streams = collections.Counter(e["args"].get("stream") for e in last)
    P(f"  records per stream: {dict(streams)}")
    # per name
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in last:
        n = e["name"].replace("void ", "").replace("(anonymous namespace)::", "").replace("uda::", "").replace("tcconv::", "")
        n = n.split("<")[0].split("(")[0]
        agg[n][0] += 1
        agg[n][1] += e["dur"]
    tot = sum(v[1] for v in agg.values())
    P(f"  sum of kernel durations {tot:.1f} us")
    P("kernel,launches,total_us,mean_us,share_of_sum_pct")
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        P(f"{n},{c},{d:.1f},{d / c:.2f},{100 * d / tot:.2f}")
    # idle gaps on the union timeline
    gaps = []
    end = last[0]["ts"] + last[0]["dur"]
    prev_e = last[0]
    for e in last[1:]:
        if e["ts"] > end:
            gaps.append((e["ts"] - end, prev_e["name"][:60], e["name"][:60]))
        if e["ts"] + e["dur"] > end:
            end = e["ts"] + e["dur"]
            prev_e = e
    hist = collections.Counter(min(int(gp[0]), 10) for gp in gaps)
    P(f"idle gaps: {len(gaps)} totalling {sum(gp[0] for gp in gaps):.1f} us; histogram (us floor -> count): {dict(sorted(hist.items()))}")
    for gp in sorted(gaps, reverse=True)[:12]:
        P(f"  gap {gp[0]:.1f} us between {gp[1]}  ->  {gp[2]}")
    if args.dump:
        P("ordered records of the last step: start_us,dur_us,stream,name")
        for e in last:
            P(f"{e['ts'] - t0:.1f},{e['dur']:.1f},{e['args'].get('stream')},{e['name'][:100]}")
    out = "\n".join(lines)
    print(out)
    if args.out:
        open(args.out, "w").write(out + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
