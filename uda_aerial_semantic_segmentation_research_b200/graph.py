"""CUDA-graph capture of the training step.

Data-parallel runs capture the gradient all-reduce INSIDE the graph: the engine's per-parameter "gradient final"
notifications make ``ddp.GradSync`` fork a communication stream after the last wgrad of every ~25 MB bucket and issue
``ncclAllReduce`` there (stream fork / join become graph dependencies), so on replay the NCCL kernels run under the
remaining backward kernels and only the last bucket is exposed.  (PyTorch's rules for capturing NCCL apply: NCCL >=
2.9.6 and ``TORCH_NCCL_ASYNC_ERROR_HANDLING=0`` set before ``init_process_group``.)

One training step of the hot path is ~900 kernel launches; issued one by one from Python the host needs
~13 ms per step — as long as the B200 needs to execute them.  ``GraphedStep`` captures
``zero_grad -> forward -> loss -> backward`` once into a CUDA graph (all kernels, memsets and TMA
descriptors of the step are baked in; shapes and buffer addresses are static) and replays it with a single
launch; the gradient all-reduce (``ddp.GradSync``, when given) and the fused optimizer step run right
after the replay.

    step = GraphedStep(model, criterion, optimizer, example_images, example_masks)
    loss = step(images, masks)          # images/masks: host (pinned) or device tensors of the captured shape

The step function has exactly the reference's semantics (``src/models/train.py:336-346``).
"""
import torch
import torch.distributed as dist


def _capture_kwargs(with_nccl):
    """Arguments of ``torch.cuda.graph``.  NCCL's watchdog thread may touch the CUDA API while a capture is open: relax
    the capture error mode to the capturing thread when collectives are captured.  (Capturing the step on a
    high-priority stream, so that the dgrad -> BatchNorm-backward chain wins SMs over the side-stream weight-gradient
    kernels, was measured SLOWER: 8.81 vs 8.50 ms/step — the weight gradients then all run after the chain.)"""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    return {"capture_error_mode": "thread_local"} if (with_nccl or multi) else {}


class _Staging:
    """Double-buffered input staging shared by the captured-step classes: ``stage(*batch)`` starts the asynchronous
    host->device copy of the NEXT batch (pinned host or device tensors) on a side stream while the current step runs;
    calling the step without arguments consumes the staged batch."""

    def _init_staging(self, inputs):
        self._in = inputs
        dev = inputs[0].device
        self._stg = [torch.empty_like(t) for t in inputs]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staged_ready, self.staged_free = torch.cuda.Event(), torch.cuda.Event()
        self.staged_free.record()
        self._staged = False

    def stage(self, *batch):
        if len(batch) != len(self._stg):
            raise ValueError(f"stage(): expected {len(self._stg)} tensors")
        self.copy_stream.wait_event(self.staged_free)
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self._stg, batch):
                dst.copy_(src, non_blocking=True)
            self.staged_ready.record()
        self._staged = True

    def _load_inputs(self, batch):
        """Copy ``batch`` (or, when empty, the staged batch) into the captured input buffers."""
        staged = len(batch) == 0 or batch[0] is None
        if staged:
            if not self._staged:
                raise RuntimeError("captured step called without a batch and none staged")
            torch.cuda.current_stream().wait_event(self.staged_ready)
            batch = self._stg
        for dst, src in zip(self._in, batch):
            dst.copy_(src, non_blocking=True)
        if staged:
            self.staged_free.record()
            self._staged = False


class GraphedStep(_Staging):
    def __init__(self, model, criterion, optimizer, example_images, example_masks, warmup=3, extra_models=(),
                 sync_in_graph=True):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs the model on a CUDA device")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.models = [model] + list(extra_models)
        self.x = example_images.to(dev, non_blocking=True).clone()
        self.t = example_masks.to(dev, non_blocking=True).clone()
        self.sync = getattr(model, "_grad_sync", None)
        # sync_in_graph: the bucketed, backward-overlapped all-reduce is captured with the step (module docstring);
        # otherwise the hooks are detached and ONE all-reduce of the flat gradient buffer runs after the replay
        self.sync_in_graph = bool(sync_in_graph and self.sync is not None and self.sync.overlap)
        saved = [(m, m._grad_sync) for m in self.models]
        if not self.sync_in_graph:
            for m in self.models:
                m._grad_sync = None
        try:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):           # allocator / workspace / attribute warm-up outside the capture
                    self._fwd_bwd()
                    self._finish()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            for m in self.models:                 # make the captured step refresh the bf16 shadow weights itself
                m._store.shadow_version = None
                m._store.shadow_ft_version = None
            self.graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            from . import ops
            l0 = ops.LAUNCHES
            with torch.cuda.graph(self.graph, **_capture_kwargs(self.sync_in_graph)):
                self.loss = self._fwd_bwd()
            self.launches_per_step = ops.LAUNCHES - l0 + 1      # captured launches + the fused Adam launch
        finally:
            for m, gs in saved:
                m._grad_sync = gs
        # staging buffers: the next batch is copied host->device on a side stream while this step runs
        self._init_staging([self.x, self.t])

    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model(self.x), self.t)
        loss.backward()
        return loss.detach()

    def _finish(self):
        if self.sync is not None and not self.sync_in_graph:
            for m in self.models:
                if m._store.grad is not None:
                    dist.all_reduce(m._store.grad, op=dist.ReduceOp.AVG, group=self.sync.group)
        self.optimizer.step()

    def __call__(self, images=None, masks=None):
        self._load_inputs(() if images is None else (images, masks))
        self.graph.replay()
        for m in self.models:
            m._store.grad_dropped = False    # the replayed backward refilled the flat gradient buffers
        self._finish()
        return self.loss


class GraphedFn(_Staging):
    """Capture an arbitrary training-step function (e.g. the adversarial D/G step of
    ``src/models/adversarial_trainer.py:84-114``) into one CUDA graph.

    ``fn(*inputs)`` must do everything of the step on the device - zero_grad, forward, losses, backward and
    ``FusedAdam(..., capturable=True).step()`` - and return a tensor (or tuple of tensors).  Networks that carry a
    ``ddp.GradSync`` get their bucketed gradient all-reduce captured with the step (overlapped with backward, see the
    module docstring), so the same class serves single-GPU and data-parallel runs.  ``networks`` are the uda_b200
    networks used, so that the captured step refreshes their bf16 shadow weights itself.
    """

    def __init__(self, fn, example_inputs, networks, warmup=3):
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedFn needs CUDA example inputs")
        with_nccl = any(getattr(m, "_grad_sync", None) is not None for m in networks)
        self.inputs = [t.clone() for t in example_inputs]
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for m in networks:
            m._store.shadow_version = None
            m._store.shadow_ft_version = None
        from . import ops
        l0 = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, **_capture_kwargs(with_nccl)):
            self.out = fn(*self.inputs)
        self.launches_per_step = ops.LAUNCHES - l0
        self._init_staging(self.inputs)

    def __call__(self, *inputs):
        self._load_inputs(inputs)
        self.graph.replay()
        return self.out


class GraphedPhases(_Staging):
    """A training step made of several captured compute phases with eager glue in between — the data-parallel form
    of the adversarial step (``src/models/adversarial_trainer.py:84-114``): D-step compute graph -> all-reduce of the
    discriminator gradients + its optimizer step -> G-step compute graph -> all-reduce of the segmentation network's
    gradients + its optimizer step.  ``phases`` = [(compute_fn, finish_fn), ...]; every ``compute_fn(*inputs)`` does
    zero_grad / forward / loss / backward on the device and returns a tensor; ``finish_fn()`` runs right after the
    replay (collectives, optimizer).  The networks must not carry an overlapping ``GradSync`` (the bucketed hooks
    are host logic that a graph cannot replay): reduce ``net._store.grad`` in ``finish_fn`` instead.
    """

    def __init__(self, phases, example_inputs, networks, warmup=2):
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedPhases needs CUDA example inputs")
        # networks that still carry a GradSync get their bucketed all-reduce captured inside the phase graphs
        with_nccl = any(getattr(m, "_grad_sync", None) is not None for m in networks)
        self.inputs = [t.clone() for t in example_inputs]
        self.networks = list(networks)
        self.finish = [f for _, f in phases]
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for compute, finish in phases:
                    compute(*self.inputs)
                    finish()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import ops
        self.graphs, self.outs, self.launches_per_step = [], [], 0
        for compute, finish in phases:
            for m in networks:                 # every phase refreshes the bf16 shadow weights it reads
                m._store.shadow_version = None
                m._store.shadow_ft_version = None
            g = torch.cuda.CUDAGraph()
            l0 = ops.LAUNCHES
            with torch.cuda.graph(g, **_capture_kwargs(with_nccl)):
                out = compute(*self.inputs)
            self.launches_per_step += ops.LAUNCHES - l0 + 1
            self.graphs.append(g)
            self.outs.append(out)
            g.replay()                         # capturing does not execute: run the phase once for real so that the
            finish()                           # next phase is captured (and later replayed) on a consistent trajectory
        self._init_staging(self.inputs)

    def __call__(self, *inputs):
        self._load_inputs(inputs)
        for g, finish in zip(self.graphs, self.finish):
            g.replay()
            for m in self.networks:
                m._store.grad_dropped = False
            finish()
        return self.outs[-1]
