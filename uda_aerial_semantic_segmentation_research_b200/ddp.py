"""Data-parallel gradient synchronisation: one process per GPU, identical replicas,
``all_reduce(AVG)`` of the flat gradient buffer in reverse-execution-order buckets, issued on a side
stream from the engine's per-parameter "gradient final" notifications so that the NCCL kernels overlap
the remaining backward.  BatchNorm statistics stay per rank (the reference has no SyncBN).

The reference has no distributed layer at all; this is the layer the north-star adds.  Works with any
``torch.distributed`` backend (NCCL on the B200 box; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class GradSync:
    """Bucketed, overlapped all-reduce over the flat gradient buffers of uda_b200 networks."""

    def __init__(self, networks, bucket_mb=25.0, process_group=None, broadcast_parameters=True, overlap=True):
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("GradSync requires an initialised torch.distributed process group")
        if not isinstance(networks, (list, tuple)):
            networks = [networks]
        self.networks = list(networks)
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_elems = max(int(bucket_mb * (1 << 20) / 4), 1)
        self.overlap = overlap
        self._plans = {}
        self._state = {}
        self._comm_stream = None
        self.producer_stream = None   # extra stream whose work (side-stream wgrads) a bucket must also wait for
        for net in self.networks:
            net._grad_sync = self
        if broadcast_parameters:
            self.broadcast_parameters()

    # -- setup ------------------------------------------------------------------------------
    def broadcast_parameters(self, src=0):
        """Make every replica start from rank ``src``'s parameters and buffers."""
        for net in self.networks:
            dev = next(net.parameters()).device
            net._store.ensure_flat(dev)
            dist.broadcast(net._store.flat, src, group=self.group)
            for b in net.buffers():
                dist.broadcast(b, src, group=self.group)
            net._store.shadow_version = None

    def _plan(self, st):
        """Buckets as [start, end) ranges of the flat buffer, last range first (backward order)."""
        plan = self._plans.get(id(st))
        if plan is None:
            order = sorted(st.params, key=lambda p: st.offsets[id(p)])
            buckets, cur_end, cur_start = [], st.total, st.total
            members, cur = [], []
            for p in reversed(order):
                off = st.offsets[id(p)]
                cur.append(id(p))
                cur_start = off
                if cur_end - cur_start >= self.bucket_elems:
                    buckets.append((cur_start, cur_end)); members.append(cur)
                    cur, cur_end = [], cur_start
            if cur:
                buckets.append((cur_start, cur_end)); members.append(cur)
            owner = {}
            for bi, ids in enumerate(members):
                for i in ids:
                    owner[i] = bi
            plan = (buckets, [len(m) for m in members], owner)
            self._plans[id(st)] = plan
        return plan

    # -- engine callbacks -------------------------------------------------------------------
    def begin(self, st):
        buckets, counts, _ = self._plan(st)
        self._state[id(st)] = {"left": list(counts), "works": [], "launched": [False] * len(buckets)}

    def param_done(self, st, p):
        state = self._state.get(id(st))
        if state is None:
            return
        buckets, _, owner = self._plan(st)
        bi = owner[id(p)]
        state["left"][bi] -= 1
        if state["left"][bi] == 0 and self.overlap:
            self._launch(st, state, bi)

    def _launch(self, st, state, bi):
        if state["launched"][bi]:
            return
        state["launched"][bi] = True
        start, end = self._plan(st)[0][bi]
        view = st.grad[start:end]
        if view.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=view.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._comm_stream.wait_event(ev)
            if self.producer_stream is not None:
                self._comm_stream.wait_stream(self.producer_stream)
            with torch.cuda.stream(self._comm_stream):
                work = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            view.record_stream(self._comm_stream)
        else:
            work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        state["works"].append(work)

    def end(self, st):
        state = self._state.pop(id(st), None)
        if state is None:
            return
        for bi in range(len(state["launched"])):
            self._launch(st, state, bi)
        for w in state["works"]:
            w.wait()
        if st.grad.is_cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        else:
            st.grad.div_(self.world)  # gloo has no AVG


def allreduce_confusion_matrix(hist, process_group=None):
    """Evaluation: windows/images shard across ranks; the int64 [C,C] histogram is summed (bit-exact in
    any order)."""
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=process_group)
    return hist
