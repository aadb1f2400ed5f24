#!/usr/bin/env python
"""Benchmark of the segmentation-training hot path (BASELINE.json metric: UDA train images/s @512x512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload supervised|adversarial|finetune]

One step = one pass of the hot path over one batch of synthetic input:
  supervised  (BASELINE configs[1], default): U-Net r34, batch 16 @512x512 bf16, 24 classes —
              zero_grad, forward, cross-entropy, backward, Adam step   (reference src/models/train.py:336-346)
  adversarial (configs[2]): the reference's discriminator step + generator step on a source and a target
              batch (src/models/adversarial_trainer.py:76-114), 8+8 images per GPU
  finetune    (configs[3]): unsupervised target-domain fine-tuning — two views of a target batch through the
              network, FineTuningLoss (ramped symmetric-KL consistency + domain confusion,
              src/models/losses.py:256-342) + entropy minimisation, clip_grad_norm 1.0, Adam
              (src/models/unsupervised_trainer.py:99-150); 4 images per GPU (global 32 on 8 GPUs)
For N>1 the script is launched by torchrun (one rank per GPU, NCCL); each rank processes its own batch
(weak scaling) and gradients are all-reduced in ~25 MB buckets overlapped with backward — the NCCL kernels are
captured into the step's CUDA graph on a side stream.  Rank 0 prints ONE JSON line; the supervised (default) line
carries an `adversarial` sub-record (configs[2], the workload the north-star's scaling target is written for).
`--impl reference` times the reference's own CPU path (oracle port: fp32 PyTorch on all host cores).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLASSES = 24
# algorithmic conv FLOPs of U-Net r34 / 24 classes per 512x512 image (counted on the oracle)
FWD_GFLOP_PER_IMG_512 = 64.248
TRAIN_GFLOP_PER_IMG_512 = 191.51


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "which": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def synthetic_batch(B, size, seed, device=None, pinned=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, size, size, generator=g)
    # blocky label map (nearest-upsampled 16x16 grid): aerial-like regions, skewed histogram bins
    t = torch.randint(0, CLASSES, (B, 16, 16), generator=g).repeat_interleave(size // 16, 1).repeat_interleave(size // 16, 2)
    t = t.contiguous()
    if pinned:
        return x.pin_memory(), t.pin_memory()
    if device is not None:
        return x.to(device), t.to(device)
    return x, t


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.ref_unet import RefUnet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    size, B = args.size, args.ref_batch
    model = RefUnet("resnet34", classes=CLASSES).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    x, t = synthetic_batch(B, size, 1234)

    def step():
        opt.zero_grad()
        loss = crit(model(x), t)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = f"{args.steps} steps of batch {B} @{size}x{size} (bounded sample of the batch-{args.batch} workload), fp32, oneDNN"
    print(json.dumps({
        "impl": "reference", "metric": "train_images_per_s", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"supervised U-Net resnet34 @{size}x{size}, {CLASSES} classes, CE + Adam (reference CPU path)",
                   "sample_batch": B},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(size, budget_s=20.0):
    """Oracle port on the box's host cores, bounded sample (rank 0, N=1 only)."""
    import torch
    from oracle.ref_unet import RefUnet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    B = 2
    model = RefUnet("resnet34", classes=CLASSES).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x, t = synthetic_batch(B, size, 1234)

    def step():
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), t).backward()
        opt.step()

    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step(); n += 1
        if time.perf_counter() - t0 > budget_s or n >= 10:
            break
    dt = time.perf_counter() - t0
    return {"value": B * n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n} steps of batch {B} @{size}x{size} after 1 warm-up (fp32 oracle U-Net r34 + CE + Adam)"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def conv_traffic_from_profile():
    """Average DRAM bytes per tensor-core convolution launch from the newest committed ncu launch list of this workload
    (profiles/r*_launch_shares*.csv: dram__bytes_read.sum + dram__bytes_write.sum per kernel).  Returns (bytes, file)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_launch_shares*.csv")))
    for path in reversed(files):
        try:
            n, tot = 0, 0.0
            for line in open(path).read().splitlines()[1:]:
                f = line.split(",")
                if f[0].startswith("conv_tc_") and f[4] and f[5]:
                    n += int(f[1])
                    tot += (float(f[4]) + float(f[5])) * 1e6
            if n:
                return tot / n, os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


def measure_workload(args, workload, env, full):
    """Build the networks / optimizers / captured step of one workload, time K steps device-resident and end to end,
    and (``full``) profile the kernel families.  Returns the result dict (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import uda_aerial_semantic_segmentation_research_b200 as U
    from uda_aerial_semantic_segmentation_research_b200 import ops, _lib
    from uda_aerial_semantic_segmentation_research_b200.losses import CrossEntropyLoss, AdversarialLoss
    from uda_aerial_semantic_segmentation_research_b200.optim import FusedAdam
    from uda_aerial_semantic_segmentation_research_b200.discriminator import DomainDiscriminator
    from uda_aerial_semantic_segmentation_research_b200.graph import GraphedStep, GraphedFn, GraphedPhases

    world, rank, local, dev = env["world"], env["rank"], env["local"], env["dev"]
    torch.manual_seed(0)
    size, B = args.size, args.batch
    adversarial, finetune, grl = workload == "adversarial", workload == "finetune", workload == "adversarial_grl"
    if finetune and args.batch == 16:
        B = 4                                               # configs[3]: 32 images on 8 GPUs
    use_graph = not args.no_graph
    model = U.Unet("resnet34", encoder_weights=None, in_channels=3, classes=CLASSES).to(dev).train()
    nets = [model]
    crit = CrossEntropyLoss()
    opt = FusedAdam(model, lr=1e-3, capturable=use_graph and adversarial, max_grad_norm=1.0 if finetune else None)
    if grl:   # output-space discriminator behind a gradient-reversal layer: ONE backward, ONE optimizer over both networks
        from uda_aerial_semantic_segmentation_research_b200.discriminator import OutputSpaceAdversary
        disc = DomainDiscriminator(CLASSES).to(dev).train()
        nets.append(disc)
        adversary, adv = OutputSpaceAdversary(disc, alpha=0.1), AdversarialLoss(0.001)
        opt = FusedAdam([model, disc], lr=1e-3, capturable=use_graph)
    if adversarial:
        disc = DomainDiscriminator(3).to(dev).train()
        nets.append(disc)
        dopt = FusedAdam(disc, lr=1e-4, capturable=use_graph)
        adv = AdversarialLoss(0.001)
    if finetune:
        from uda_aerial_semantic_segmentation_research_b200.losses import FineTuningLoss, EntropyMinimizationLoss
        disc = DomainDiscriminator(3).to(dev).train()      # frozen critic: only its prediction enters the loss
        for q in disc.parameters():
            q.requires_grad_(False)
        ft_loss, ent_loss = FineTuningLoss(), EntropyMinimizationLoss(0.1)
    if world > 1:
        from uda_aerial_semantic_segmentation_research_b200.ddp import GradSync
        GradSync(nets)                                      # bucketed all-reduce, overlapped with backward
        # Captured steps: by default ONE flat all-reduce of the gradient buffer right after the backward kernels.
        # --nccl-in-graph captures the bucketed, backward-overlapped all-reduces instead; measured slower on 2 and 8
        # B200s (8.91 vs 8.71 ms/step at N=8, profiles/r02_bench_*_8gpu.json): the NCCL CTAs hold SMs for the duration of
        # every bucket and break the one-CTA-per-SM assumption of the persistent convolution kernels.
        if use_graph and not args.nccl_in_graph:
            for n in nets:
                n._grad_sync = None

    def flat_allreduce(net):
        if world > 1 and net._grad_sync is None:
            dist.all_reduce(net._store.grad, op=dist.ReduceOp.AVG)

    # ---- inputs ------------------------------------------------------------------------------------------------
    Bs = B // 2 if (adversarial or grl) else B
    hx, ht = synthetic_batch(Bs, size, 1234 + rank, pinned=True)
    host = [hx, ht]
    if adversarial:
        host.append(synthetic_batch(Bs, size, 4321 + rank, pinned=True)[0])
    if grl:       # source and target batch travel as one [2*Bs,...] tensor (one pass of the segmentation network)
        host = [torch.cat([hx, synthetic_batch(Bs, size, 4321 + rank)[0]]).pin_memory(), ht]
    if finetune:   # second view = flipped + jittered copy (stand-in for the strong augmentation)
        gq = torch.Generator().manual_seed(77 + rank)
        hx2 = hx.flip(-1) + 0.1 * torch.randn(hx.shape, generator=gq)
        host = [torch.cat([hx, hx2]).pin_memory(), ht] if args.ft_views == "pooled" else [hx, ht, hx2.pin_memory()]
    devin = [h.to(dev) for h in host]

    # ---- the step (reference semantics) ------------------------------------------------------------------------
    def sup_step(xs, ts):                # src/models/train.py:336-346
        opt.zero_grad()
        loss = crit(model(xs), ts)
        loss.backward()
        flat_allreduce(model)
        opt.step()
        return loss.detach()

    def adv_step(xs, ts, xtg):           # src/models/adversarial_trainer.py:84-114
        dopt.zero_grad()
        d_loss = adv.discriminator_loss(disc(xs), disc(xtg))
        d_loss.backward()
        flat_allreduce(disc)
        dopt.step()
        opt.zero_grad()
        total = crit(model(xs), ts) + adv.generator_loss(disc(xtg))
        total.backward()
        flat_allreduce(model)
        opt.step()
        return total.detach()

    class SplitViews(torch.autograd.Function):
        """logits [2n,...] -> the two halves; backward re-joins the two logit gradients with one copy (autograd's own
        slice backward would materialise two zero-padded full-size tensors and add them)."""

        @staticmethod
        def forward(ctx, p):
            n = p.shape[0] // 2
            return p[:n], p[n:]

        @staticmethod
        def backward(ctx, g1, g2):
            g1 = torch.zeros_like(g2) if g1 is None else g1
            g2 = torch.zeros_like(g1) if g2 is None else g2
            return torch.cat([g1, g2])

    def grl_step(xcat, ts):              # north-star configs[2]: output-space discriminator + gradient reversal
        opt.zero_grad()
        logits = model(xcat)             # source images first, target images second
        ls, _lt = SplitViews.apply(logits)
        dom = adversary(logits)          # D(GRL(softmax(logits))): [2*Bs, 1]
        total = crit(ls, ts) + adv.discriminator_loss(dom[:Bs], dom[Bs:])
        total.backward()
        flat_allreduce(model)
        flat_allreduce(disc)
        opt.step()
        return total.detach()

    def ft_compute(*inp):                # src/models/unsupervised_trainer.py:99-150
        opt.zero_grad()
        if args.ft_views == "pooled":    # both views in ONE pass of 2B images, split at the loss
            xv, ts = inp
            p1, p2 = SplitViews.apply(model(xv))
            xs = xv[:B]
        else:
            xs, ts, x2 = inp
            p1, p2 = model(xs), model(x2)
        with torch.no_grad():
            dpred = disc(xs)
        total = ft_loss(p1, p2, dpred, 40)["total"] + ent_loss(p1)
        total.backward()
        return total.detach()

    def ft_finish():
        flat_allreduce(model)
        opt.step()

    def ft_step(*inp):
        out = ft_compute(*inp)
        ft_finish()
        return out

    eager = ft_step if finetune else adv_step if adversarial else grl_step if grl else sup_step
    graphed, launch = None, "eager launches"
    nccl = "" if world == 1 else ("; bucketed NCCL all-reduce captured inside the graph on a side stream (overlaps backward)"
                                  if args.nccl_in_graph else "; one flat NCCL all-reduce of the gradient buffer after backward")
    if use_graph:
        if finetune:
            graphed = GraphedPhases([(ft_compute, ft_finish)], devin, [model, disc])
            launch = "cuda-graph replay of fwd+loss+bwd, then clip + fused Adam" + nccl
        elif grl:
            graphed = GraphedFn(grl_step, devin, nets)
            launch = "ONE cuda-graph replay of the whole step (fwd, CE + adversarial loss, bwd, fused Adam over both networks)" + nccl
        elif adversarial:
            # one graph for every world size: the two flat all-reduces (or, with --nccl-in-graph, the bucketed ones) are
            # captured in-stream between the backward kernels and the captured fused-Adam launches
            graphed = GraphedFn(adv_step, devin, nets)
            launch = "ONE cuda-graph replay of the whole D step + G step (both fused Adam steps captured)" + nccl
        else:
            graphed = GraphedStep(model, crit, opt, devin[0], devin[1])
            launch = "cuda-graph replay of fwd+loss+bwd, then fused Adam" + nccl
    step = graphed if graphed is not None else eager

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm -----------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(*devin)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.LAUNCHES
    ms = timed(lambda: step(*devin), args.steps)
    launches = ops.LAUNCHES - l0
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps
    sampler.stop_flag = True
    sampler.join(1.0)
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)

    # ---- end-to-end arm: pinned host inputs, H2D inside the timed region, loss read back ------------------------
    def e2e_eager():
        return step(*[h.to(dev, non_blocking=True) for h in host]).item()

    def e2e_pipelined(steps):
        # same bytes per step; the H2D copy of batch i+1 runs on a side stream under the compute of batch i
        graphed.stage(*host)
        for i in range(steps):
            loss = graphed()
            if i + 1 < steps:
                graphed.stage(*host)
            loss.item()

    if graphed is not None:
        e2e_pipelined(2)
        ms_e2e = timed(lambda: e2e_pipelined(args.steps), 1)
    else:
        for _ in range(2):
            e2e_eager()
        ms_e2e = timed(e2e_eager, args.steps)
    h2d = sum(h.numel() * h.element_size() for h in host)
    e2e = {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    wl = (f"supervised source-domain training, U-Net resnet34, batch {B}/GPU @{size}x{size}, {CLASSES} classes, "
          "CE loss + Adam (BASELINE configs[1])") if not (adversarial or finetune or grl) else \
         (f"unsupervised target-domain fine-tuning (two views, FineTuningLoss consistency + domain confusion + entropy "
          f"minimisation, clip 1.0, Adam), U-Net resnet34, {B} images/GPU @{size}x{size} (BASELINE configs[3]); "
          + ("both views run as ONE pass of 2B images split at the loss (BatchNorm statistics pooled over the two views); "
             if args.ft_views == "pooled" else "two separate passes of B images; ")
          + "deviation from src/models/unsupervised_trainer.py:116-122: its third, unused segmentation forward of the "
            "un-augmented batch is not run, the critic is frozen") if finetune else \
         (f"adversarial UDA with an OUTPUT-SPACE discriminator behind a gradient-reversal layer (north-star configs[2]): "
          f"U-Net resnet34 on {Bs} source + {Bs} target images/GPU @{size}x{size} in one pass, CE(source) + "
          f"BCE(D(GRL(softmax(logits)))), one backward, one fused Adam over both networks; beyond the reference, which "
          f"defines the gradient-reversal layer and the discriminator but never wires them (SURVEY T3)") if grl else \
         (f"adversarial UDA step (D step + G step), U-Net resnet34 + image discriminator, {Bs}+{Bs} images/GPU "
          f"@{size}x{size} (BASELINE configs[2])")
    res = {"value": value, "ms_per_step": ms / args.steps, "e2e": e2e, "gpu_launches": launches,
           "clocks": sampler.result(), "workload": wl, "launch": launch, "global_batch": B * world}
    if not full:
        del model, nets, graphed, step
        torch.cuda.empty_cache()
        return res if rank == 0 else None

    # ---- per-kernel CUDA-event profile of a few steps (rank 0): roofline of the dominant kernel -----------------
    roof, breakdown = None, {}
    prof_steps = 3
    if rank == 0:
        _lib.PROFILE = {}
    sync()
    f0 = ops.TC_FLOPS
    ws_saved, ops.WGRAD_STREAM = ops.WGRAD_STREAM, False   # one stream: event pairs then bracket ONE kernel family each
    for _ in range(prof_steps):                 # every rank runs them (they contain the gradient all-reduce)
        torch.cuda._sleep(int(60e-3 * 1.9e9))   # let the host run ahead: event pairs then bracket pure device time
        eager(*devin)
    torch.cuda.synchronize()
    ops.WGRAD_STREAM = ws_saved
    if rank == 0:
        prof, _lib.PROFILE = _lib.PROFILE, None
        for name, evs in prof.items():
            tot = sum(a.elapsed_time(b) for a, b in evs)
            breakdown[name] = {"ms_per_step": tot / prof_steps, "launches_per_step": len(evs) / prof_steps}
            if os.environ.get("UDA_B200_PROFILE_DUMP"):   # per-call durations (us) of the last profiled step
                per = len(evs) // prof_steps
                last = [round(a.elapsed_time(b) * 1e3, 1) for a, b in evs[-per:]]
                print(f"[profile] {name}: {last}", file=sys.stderr)
        pk = peaks()
        # every entry point whose FLOPs ops._tc_account() counts (the two halves of the split decoder conv1 included)
        # conv2d_tc_fwd_bn_act = convolution + BatchNorm + activation in one launch: its WHOLE time is charged to the
        # convolutions (the launch contains the grid barrier and the normalisation pass) - conservative, not flattering
        tc_keys = ("conv2d_tc_fwd", "conv2d_tc_fwd_add", "conv2d_tc_fwd_fused", "conv2d_tc_fwd_bn_act", "upconv_tc_fwd",
                   "conv2d_tc_dgrad", "conv2d_tc_dgrad_bnstats", "conv2d_tc_wgrad", "stem_tc_fwd", "stem_tc_fwd_act",
                   "stem_tc_wgrad")
        tc_ms = sum(breakdown.get(k, {}).get("ms_per_step", 0.0) for k in tc_keys)
        tc_n = sum(breakdown.get(k, {}).get("launches_per_step", 0.0) for k in tc_keys)
        tc_gflop = (ops.TC_FLOPS - f0) / prof_steps / 1e9
        top = max(breakdown.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if breakdown else None
        if tc_ms > 0:
            ach = tc_gflop / tc_ms  # GFLOP / ms == TFLOP/s
            traffic, tfile = conv_traffic_from_profile()
            roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM convolution family (conv_tc_persist / conv_tc_halo / "
                                                 "conv_tc_phalo / conv_tc_wgrad[_big|_halo] kernels: fwd + dgrad + wgrad launches)",
                    "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic,
                    "traffic_note": f"average DRAM bytes per conv launch (dram__bytes_read.sum + dram__bytes_write.sum) from the "
                                    f"committed ncu launch list {tfile} (cold cache, same command); not re-measured in this run",
                    "peak_source": pk["which"] + " (sustained: kernels timed inside a long step)",
                    "gflop_per_step": tc_gflop, "ms_per_step": tc_ms, "launches_per_step": tc_n,
                    "fused_bn_launches_ms_per_step": breakdown.get("conv2d_tc_fwd_bn_act", {}).get("ms_per_step", 0.0),
                    "timing": "CUDA events around every entry point of eager single-stream replays of the step (the timed "
                              "region itself is a graph replay with the wgrad launches on a second stream)",
                    "whole_step_frac": tc_gflop / (ms / args.steps) / pk["bf16_tflops_sustained"],
                    "top_entry_point_by_time": top}
    hbm_kernels = None
    if rank == 0:
        P = float(Bs * size * size)
        # algorithmic bytes per launch, fp32 NCHW logits as the reference passes them
        alg = {"seg_loss_fwd_bwd": 2.0 * P * CLASSES * 4 + 8.0 * P,      # CE: logits read, targets read, gradient written
               "consistency_fwd_bwd": 4.0 * P * CLASSES * 4,             # two logit tensors read, two gradients written
               "entropy_fwd_bwd": 2.0 * P * CLASSES * 4}
        for name, nbytes in alg.items():
            if name in breakdown and breakdown[name]["ms_per_step"] > 0:
                n_launch = max(breakdown[name]["launches_per_step"], 1.0)
                t_ms = breakdown[name]["ms_per_step"] / n_launch
                gbs = nbytes / (t_ms * 1e-3) / 1e9
                hbm_kernels = hbm_kernels or {}
                hbm_kernels[name] = {"bound": "hbm", "achieved": gbs, "peak": peaks()["hbm_gbs"], "unit": "GB/s",
                                     "frac": gbs / peaks()["hbm_gbs"], "frac_of_8TBs_nominal": gbs / 8000.0,
                                     "algorithmic_mb_per_launch": nbytes / 1e6, "ms_per_launch": t_ms,
                                     "note": "includes the finalize / rescale launches of the entry point"}
    res.update({"roofline": roof, "hbm_kernels": hbm_kernels,
                "kernel_breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in
                                                 sorted(breakdown.items(), key=lambda kv: -kv[1]["ms_per_step"])}})
    del model, nets, graphed, step
    torch.cuda.empty_cache()
    return res if rank == 0 else None


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # NCCL collectives are captured into the CUDA graphs (PyTorch's rule for that)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    env = {"world": world, "rank": rank, "local": local, "dev": dev}
    size = args.size
    head = measure_workload(args, args.workload, env, full=True)
    subs = {}
    if args.workload == "supervised" and not args.no_sub:
        # the north-star's scaling target is written for the adversarial UDA step: measure it in the same line, both as
        # the reference's D-step / G-step iteration and as the output-space + gradient-reversal variant
        for name in ("adversarial", "adversarial_grl"):
            subs[name] = measure_workload(args, name, env, full=False)
    if rank == 0:
        out = {
            "metric": "train_images_per_s", "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": head["workload"], "global_batch": head["global_batch"], "image_size": size,
                       "classes": CLASSES, "parallelism": f"dp{world}",
                       "l2": "inputs+activations per step (>3 GB) exceed the 126 MB L2", "launch": head["launch"]},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
            "roofline": head["roofline"], "hbm_kernels": head["hbm_kernels"],
            "kernel_breakdown_ms_per_step": head["kernel_breakdown_ms_per_step"],
            "fraction_of_flop_roofline": head["value"] / world / (peaks()["bf16_tflops"] * 1e3 / TRAIN_GFLOP_PER_IMG_512 * (512 / size) ** 2),
            # every UDA_B200_* switch present in the environment of this run (none = the shipped defaults)
            "env": {k: v for k, v in sorted(os.environ.items()) if k.startswith("UDA_B200_")},
        }
        for name, sub in subs.items():
            if sub is not None:
                out[name] = {"metric": "train_images_per_s", "value": sub["value"], "unit": "images/s",
                             "ms_per_step": sub["ms_per_step"], "e2e": sub["e2e"], "gpu_launches": sub["gpu_launches"],
                             "n_gpus": world, "scaling": "weak",
                             "config": {"workload": sub["workload"], "global_batch": sub["global_batch"],
                                        "launch": sub["launch"]}}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(size)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="supervised", choices=["supervised", "adversarial", "adversarial_grl", "finetune"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--ref-batch", type=int, default=2, help="bounded sample batch of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-sub", action="store_true", help="skip the adversarial sub-record of the supervised line")
    ap.add_argument("--nccl-in-graph", action="store_true",
                    help="A/B switch (N>1): capture the bucketed, backward-overlapped all-reduces inside the step graph "
                         "instead of one flat all-reduce after the backward kernels")
    ap.add_argument("--ft-views", default="pooled", choices=["pooled", "separate"],
                    help="finetune workload: both views as one 2B pass (default) or two B passes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
