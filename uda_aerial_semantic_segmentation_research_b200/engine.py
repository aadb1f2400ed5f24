"""Layer engine: NHWC forward / manual backward of the conv-BN-activation networks on the CUDA kernels.

The reference reaches these layers through ``nn.Module`` calls into ATen/cuDNN; here a
network forward is a straight-line sequence of C-ABI launches (``ops``) recorded on a small tape, and
the backward replays the tape in reverse.  One ``torch.autograd.Function`` per network (see
``unet.py`` / ``discriminator.py``) makes the whole thing a single autograd node, so the reference's
``loss.backward(); optimizer.step()`` call sites work unchanged.

Gradient accumulation for tensors with several consumers (residual identity, encoder skips) is
folded into the consumer kernels through their ``addend`` argument instead of separate add passes.
"""
import os

import torch
import torch.nn as nn

from . import ops


class Var:
    """An activation (NHWC tensor) plus its gradient slot on the tape.  A network input that feeds a
    tensor-core stem keeps the raw fp32 NCHW image in ``nchw`` instead (``t`` is None)."""
    __slots__ = ("t", "g", "nchw", "sums", "uses", "bn_req", "bwd_sums")

    def __init__(self, t, nchw=None):
        self.t = t
        self.g = None
        self.nchw = nchw
        self.sums = None   # BatchNorm batch statistics emitted by the producing conv epilogue (double[2C])
        self.uses = 0      # consumers still to contribute to .g during backward (counted at forward time)
        self.bn_req = None     # (z tensor or None, slope): this Var is a = act(BN(z) (+res)); set by bn_act
        self.bwd_sums = None   # BatchNorm-backward sums of .g, filled by the dgrad that made the LAST contribution


def input_var(x, dtype, cp, need_grad):
    """Network input: NHWC copy, or (Cin=3 stem on the tensor cores, no input gradient needed) the raw image."""
    x = x.contiguous().float()
    B, C, H, W = x.shape
    if (dtype == torch.bfloat16 and not need_grad and x.is_cuda
            and ops.stem_supported(B, H, W, C, cp.out_channels, cp.kernel_size, cp.stride, cp.padding)):
        return Var(None, nchw=x)
    return Var(ops.nchw_to_nhwc(x, dtype))


class Tape:
    def __init__(self):
        self.fns = []

    def push(self, fn):
        self.fns.append(fn)

    def backward(self):
        for fn in reversed(self.fns):
            fn()
        self.fns = []


class ConvParams(nn.Module):
    """Parameter holder with ``nn.Conv2d``'s state_dict keys (``weight`` [O,I,KH,KW], ``bias``).

    The weight is stored channels_last, i.e. physically OHWI — the layout the kernels read.
    Forward is executed by the fused engine, never through cuDNN.
    """

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding = k, stride, padding
        w = torch.empty(cout, cin, k, k).contiguous(memory_format=torch.channels_last)
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

    def forward(self, *a, **k):
        raise RuntimeError("ConvParams is executed by the uda_b200 engine (call the owning network)")

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, padding={self.padding}, bias={self.bias is not None}")


class BNParams(nn.Module):
    """Parameter/buffer holder with ``nn.BatchNorm2d``'s state_dict keys (eps 1e-5, momentum 0.1)."""

    def __init__(self, c, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = c, eps, momentum
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, *a, **k):
        raise RuntimeError("BNParams is executed by the uda_b200 engine (call the owning network)")

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


class LinearParams(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.in_features, self.out_features = cin, cout
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout))

    def forward(self, *a, **k):
        raise RuntimeError("LinearParams is executed by the uda_b200 engine (call the owning network)")


class ParamStore:
    """Flat fp32 parameter buffer (+ bf16 shadow) and per-backward flat gradient buffer of a network.

    ``nn.Parameter`` objects stay the public handles (state_dict / optimizers / DDP see them), but
    their storage is re-pointed into one contiguous fp32 buffer so that the bf16 shadow refresh, the
    fused Adam and the gradient all-reduce are single launches over one range.
    """

    def __init__(self, module):
        self.module = module
        self.params = [p for p in module.parameters()]
        self.offsets = {}
        off = 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) // 4 * 4  # keep every slice 16-byte aligned
        self.total = off
        self.flat = None
        self.shadow = None
        self.shadow_version = None
        self.grad = None
        self.grad_dropped = False   # zero_grad() since the last backward (FusedAdam refuses to step on stale gradients)
        self.stats_epoch = 0        # training forwards so far (running statistics are updated by kernels)
        self._fold_cache = {}

    # -- parameters -------------------------------------------------------------------------
    def _view_like(self, flat, p):
        off = self.offsets[id(p)]
        seg = flat[off:off + p.numel()]
        if p.dim() == 4:
            O, I, KH, KW = p.shape
            return seg.view(O, KH, KW, I).permute(0, 3, 1, 2)  # logical OIHW, physical OHWI
        return seg.view(p.shape)

    def ensure_flat(self, device):
        if self.flat is not None and self.flat.device == device:
            base = self.flat.data_ptr()
            ok = all(p.data_ptr() == base + 4 * self.offsets[id(p)] for p in self.params)
            if ok:
                return
        flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for p in self.params:
                v = self._view_like(flat, p)
                v.copy_(p.detach().to(device))
                p.data = v
        self.flat = flat
        self.shadow = torch.empty(self.total, dtype=torch.bfloat16, device=device)
        self.shadow_ft = torch.empty(self.total, dtype=torch.bfloat16, device=device)
        self.shadow_version = None
        self.shadow_ft_version = None
        self._ft_table = None
        self._ft_event = None

    def param_version(self):
        """Freshness key of the bf16 shadow copies.  The parameters are re-pointed into the flat buffer with
        ``p.data = view``, so every ``nn.Parameter`` keeps its OWN version counter: in-place writes through the
        parameter (``torch.optim.Adam``, ``load_state_dict``, ``p.copy_``) bump ``p._version`` and never
        ``flat._version``.  The key is therefore the sum over the parameters (plus the flat buffer's, for writes
        through it); the fused Adam kernel refreshes the shadow itself and calls ``mark_shadow_fresh``."""
        return self.flat._version + sum(p._version for p in self.params)

    def refresh_shadow(self, force=False):
        ver = self.param_version()
        if force or self.shadow_version != ver:
            ops.cast_f32(self.flat, self.shadow)
            self.shadow_version = ver
            self.shadow_ft_version = None

    def mark_shadow_fresh(self):
        self.shadow_version = self.param_version()
        self.shadow_ft_version = None

    def w_ft(self, p):
        """[Cin][KH][KW][Cout] flipped/transposed bf16 copy of conv weight ``p`` (dgrad on the tensor cores);
        all copies are refreshed together, lazily, the first time a backward needs them after an update."""
        ver = self.param_version()
        if self.shadow_ft_version != ver or self.shadow_ft_version is None:
            self._flip_transpose_all()
            self.shadow_ft_version = ver
            self._ft_event = None
        elif self._ft_event is not None:     # refreshed on the side stream during the forward (prefetch_w_ft)
            torch.cuda.current_stream().wait_event(self._ft_event)
            self._ft_event = None
        off = self.offsets[id(p)]
        O, I, KH, KW = p.shape
        return self.shadow_ft[off:off + p.numel()].view(I, KH, KW, O)

    def _flip_transpose_all(self):
        if self._ft_table is None:
            rows = [[self.offsets[id(q)]] + [q.shape[0], q.shape[1], q.shape[2], q.shape[3]]
                    for q in self.params if q.dim() == 4]
            self._ft_table = torch.tensor(rows, dtype=torch.int32, device=self.flat.device).contiguous()
        ops.weight_flip_transpose_batch(self.shadow, self.shadow_ft, self._ft_table)

    def prefetch_w_ft(self, side):
        """Refresh the dgrad weight copies on ``side`` while the forward runs (they are first needed in the backward and
        depend only on the bf16 shadow): ~80 us off the step's critical chain.  ``w_ft`` waits for the event."""
        ver = self.param_version()
        if self.shadow_ft_version == ver:
            return False
        side.wait_stream(torch.cuda.current_stream(self.flat.device))      # after the shadow refresh / earlier dgrads
        with torch.cuda.stream(side):
            self._flip_transpose_all()
            self._ft_event = torch.cuda.Event()
            self._ft_event.record(side)
        self.shadow_ft_version = ver
        return True

    def w(self, p, dtype):
        """OHWI kernel view of conv weight ``p`` in ``dtype`` (bf16 -> shadow copy)."""
        off = self.offsets[id(p)]
        O, I, KH, KW = p.shape
        src = self.shadow if dtype == torch.bfloat16 else self.flat
        return src[off:off + p.numel()].view(O, KH, KW, I)

    # -- eval-mode BatchNorm folding ----------------------------------------------------------------
    def fold_key(self):
        """Freshness key of the folded (conv x BatchNorm) inference weights: parameters, BatchNorm buffers (in-place
        loads bump their versions) and the number of training forwards (the kernels update the running statistics
        without torch noticing)."""
        return (self.param_version(), sum(b._version for b in self.module.buffers()), self.stats_epoch)

    def folded(self, key, cp, bn):
        """(bf16 OHWI weights, fp32 bias) of ``cp`` with ``bn`` folded in; cached until ``key`` changes."""
        hit = self._fold_cache.get(id(cp))
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        w, b = ops.bn_fold_conv(self.w(cp.weight, torch.float32), cp.bias, bn.weight, bn.bias, bn.running_mean,
                                bn.running_var, bn.eps)
        self._fold_cache[id(cp)] = (key, w, b)
        return w, b

    # -- gradients --------------------------------------------------------------------------
    def new_grad(self):
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=self.flat.device)
        self.grad_dropped = False
        return self.grad

    def g(self, p):
        off = self.offsets[id(p)]
        seg = self.grad[off:off + p.numel()]
        if p.dim() == 4:
            O, I, KH, KW = p.shape
            return seg.view(O, KH, KW, I)
        return seg.view(p.shape)

    def grad_views(self):
        """Gradients shaped like the parameters (zero-copy views of the flat buffer)."""
        return [self._view_like(self.grad, p) for p in self.params]


_SIDE_STREAMS = {}


def _side_stream(device):
    """One persistent side stream per device (persistent so that CUDA-graph captures see a stable stream)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    s = _SIDE_STREAMS.get(key)
    if s is None:
        # UDA_B200_WGRAD_PRIORITY=1: high-priority side stream (A/B switch)
        prio = -1 if os.environ.get("UDA_B200_WGRAD_PRIORITY", "0") == "1" else 0
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=prio)
    return s


class Ctx:
    """Per-forward execution context."""

    def __init__(self, store, dtype, training, tape):
        self.store, self.dtype, self.training, self.tape = store, dtype, training, tape
        self.sync = None  # optional gradient-sync object (ddp.GradSync): param_done(store, param)
        self.side = None          # second stream carrying the wgrad launches of this backward (ops.WGRAD_STREAM)
        self._keep = []           # tensors the side stream still reads: freed only after the join
        if tape is not None:
            tape.push(self._join_side)      # first pushed = last executed in backward
        self.fold_key = None                    # eval mode: freshness key of the folded weights, taken once per forward
        self._pool, self._pool_used = None, 0   # zero-filled float64 scratch for the fused BatchNorm statistics
        self._trackers = []                     # num_batches_tracked buffers to bump at the end of the forward
        if (tape is not None and dtype == torch.bfloat16 and ops.USE_TC and ops.WGRAD_STREAM and ops.PREFETCH_WFT
                and store.flat is not None and store.flat.is_cuda):
            side = _side_stream(store.flat.device)
            if store.prefetch_w_ft(side):
                self.side = side                # joined at the end of the backward (_join_side)

    def stats_slot(self, n, device):
        """``n`` zeroed float64 values for a conv epilogue's BatchNorm sums: slices of ONE zero-filled pool per
        forward instead of one ``torch.zeros`` launch per layer."""
        n_al = (n + 15) // 16 * 16
        if self._pool is None or self._pool_used + n_al > self._pool.numel():
            self._pool = torch.zeros(max(1 << 16, n_al), dtype=torch.float64, device=device)
            self._pool_used = 0
        out = self._pool[self._pool_used:self._pool_used + n]
        self._pool_used += n_al
        return out

    def finish_forward(self):
        """One fused increment of every BatchNorm's ``num_batches_tracked`` (instead of one launch per layer)."""
        if self._trackers:
            torch._foreach_add_(self._trackers, 1)
            self._trackers = []
            self.store.stats_epoch += 1

    def wgrad_stream(self, *tensors):
        """Context manager: run the enclosed wgrad launches on the side stream, after everything issued so far."""
        dev = tensors[0].device
        if self.side is None:
            self.side = _side_stream(dev)
        self.side.wait_stream(torch.cuda.current_stream(dev))
        self._keep.extend(tensors)
        return torch.cuda.stream(self.side)

    def _join_side(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
            self.side = None
        self._keep = []

    def done(self, *params):
        """Tell the data-parallel layer that the gradients of ``params`` are final for this backward."""
        if self.sync is not None and self.side is not None:
            self.sync.producer_stream = self.side       # the bucket's all-reduce must also wait for the wgrad stream
        if self.sync is not None:
            for p in params:
                if p is not None:
                    self.sync.param_done(self.store, p)


# ------------------------------------------------------------------------------------------------
# tape ops
# ------------------------------------------------------------------------------------------------
def _fused_stats_ok(ctx, xin, cp):
    """Can the tensor-core conv of this layer emit the BatchNorm statistics from its epilogue?"""
    if not (ctx.training and ctx.dtype == torch.bfloat16 and ops.USE_TC and ops.FUSE_BN_STATS):
        return False
    if xin.t is None:
        return True
    B, H, W, Cin = xin.t.shape
    return ops.tc_supported(0, B, H, W, Cin, cp.out_channels, cp.kernel_size, cp.kernel_size, cp.stride, cp.padding)


def conv(ctx, xin, cp, nchw_out=False, bn=None, _pre=None):
    """y = conv(x) (+bias).  Returns Var (NHWC) or, with nchw_out, a raw fp32 NCHW tensor + grad hook.
    ``bn``: the BatchNorm that consumes the output — its batch statistics are then accumulated by the conv
    epilogue (``out.sums``) instead of a separate pass over the output.  ``_pre``: the output was already computed
    by a fused launch (``_conv_bn_act_fused``); only the backward is recorded."""
    st = ctx.store
    w = st.w(cp.weight, ctx.dtype)
    sums = None
    if _pre is None and bn is not None and _fused_stats_ok(ctx, xin, cp):
        sums = ctx.stats_slot(2 * cp.out_channels, w.device)
    if xin.t is None:   # Cin = 3 stem on the tensor cores (see ops.stem_*)
        H, W = xin.nchw.shape[2:]
        K, pad = cp.kernel_size, cp.padding
        xs = ops.stem_pack_input(xin.nchw, pad)
        out = Var(ops.stem_fwd(xs, ops.stem_pack_weight(w), cp.bias, H, W, K, pad, bn_sums=sums))
        out.sums = sums
        if ctx.tape is not None:
            def bwd_stem():
                dy = out.g
                out.g = None
                if dy is None:
                    return
                if cp.bias is not None:
                    ops.colsum(dy, st.g(cp.bias))
                ops.stem_wgrad(dy, xs, st.g(cp.weight), H, W, K, pad)
                ctx.done(cp.weight, cp.bias)
            ctx.tape.push(bwd_stem)
        return out
    y = _pre if _pre is not None else ops.conv_fwd(xin.t, w, cp.bias, cp.stride, cp.padding, nchw_out=nchw_out,
                                                   bn_sums=sums)
    out = Var(y)
    out.sums = sums
    if ctx.tape is not None:
        xin.uses += 1

        def bwd():
            dy = out.g
            out.g = None
            xin.uses -= 1
            if dy is None:
                return
            if ops.WGRAD_STREAM and dy.is_cuda:
                with ctx.wgrad_stream(dy, xin.t):
                    if cp.bias is not None:
                        ops.colsum(dy, st.g(cp.bias))
                    ops.conv_wgrad(dy, xin.t, st.g(cp.weight), cp.stride, cp.padding)
            else:
                if cp.bias is not None:
                    ops.colsum(dy, st.g(cp.bias))
                ops.conv_wgrad(dy, xin.t, st.g(cp.weight), cp.stride, cp.padding)
            if xin.g is not False:  # False marks "no gradient needed" (network input)
                wft = st.w_ft(cp.weight) if (ctx.dtype == torch.bfloat16 and ops.USE_TC) else None
                stats = None
                # last contribution to dL/da of a BatchNorm output: its epilogue also reduces the BN-backward sums
                if (xin.uses == 0 and xin.bn_req is not None
                        and ops.dgrad_bnstats_supported(xin.t.shape, w.shape, cp.stride, cp.padding, ctx.dtype)):
                    z, slope = xin.bn_req
                    xin.bwd_sums = ctx.stats_slot(2 * xin.t.shape[-1], xin.t.device)
                    stats = (xin.t, z, slope, xin.bwd_sums)
                xin.g = ops.conv_dgrad(dy, w, xin.t.shape, cp.stride, cp.padding, addend=xin.g, w_ft=wft, bn_stats=stats)
            ctx.done(cp.weight, cp.bias)
        ctx.tape.push(bwd)
    return out


def conv_padded_cin(ctx, xin, cp):
    """y = conv(x) + bias for an input tensor whose channel dimension is zero-padded beyond ``cp.in_channels`` (the
    24-class probability map padded to 32 channels: a tensor-core channel atom).  The weights are padded on the fly
    (a few thousand elements), the weight gradient is un-padded into the flat gradient buffer."""
    st = ctx.store
    cpad = xin.t.shape[-1]
    w = st.w(cp.weight, ctx.dtype)
    wp = ops.pad_channels(w, cpad)
    y = ops.conv_fwd(xin.t, wp, cp.bias, cp.stride, cp.padding)
    out = Var(y)
    if ctx.tape is not None:
        xin.uses += 1

        def bwd():
            dy = out.g
            out.g = None
            xin.uses -= 1
            if dy is None:
                return
            if cp.bias is not None:
                ops.colsum(dy, st.g(cp.bias))
            dwp = torch.zeros(wp.shape, dtype=torch.float32, device=wp.device)
            ops.conv_wgrad(dy, xin.t, dwp, cp.stride, cp.padding)
            ops.unpad_channels_add(dwp, st.g(cp.weight))
            if xin.g is not False:
                xin.g = ops.conv_dgrad(dy, wp, xin.t.shape, cp.stride, cp.padding, addend=xin.g)
            ctx.done(cp.weight, cp.bias)
        ctx.tape.push(bwd)
    return out


def _upconv_ok(ctx, xin, skip, cp):
    if not (ctx.dtype == torch.bfloat16 and ops.USE_TC and ops.TC_PERSIST and ops.FUSE_UPCAT and xin.t is not None
            and cp.kernel_size == 3 and cp.stride == 1 and cp.padding == 1 and cp.bias is None):
        return False
    B, h, w, C1 = xin.t.shape
    H, W, O = 2 * h, 2 * w, cp.out_channels
    # measured per block at B=16, 512x512 (tools/upconv_bench.py, profiles/r02_upconv_bench.txt): the split form wins
    # from decoder block 1 on (-10 % .. -36 % forward + backward); on block 0 (512 upsampled channels on 16x16 pixels:
    # four parity classes of 256-pixel work each) the dgrad / wgrad launches of the two halves cost more than the copy
    if C1 > 256:
        return False
    ok = all(ops.tc_supported(op, B, H, W, O, C1, 4, 4, 2, 1) for op in (0, 1, 2))      # the 4x4 stride-2 convolution
    if skip is not None:
        C2 = skip.t.shape[-1]
        ok = ok and all(ops.tc_supported(op, B, H, W, C2, O, 3, 3, 1, 1) for op in (0, 1, 2))
    return ok


def upconv_bn_act(ctx, xin, skip, cp, bn, slope=0.0):
    """Decoder conv1: a = act(BN(conv3x3(cat(upsample2x(x), skip)))) WITHOUT the upsampled / concatenated tensor:

        conv3x3(cat(up2(x), skip), W) = conv_transpose4x4_s2_p1(x, W4) + conv3x3(skip, Ws)

    (on the nearest-upsampled image the nine taps of an output pixel fall on 2 x 2 pixels of x: per output parity the
    3x3 kernel collapses to 2x2, together a 4x4 stride-2 transposed convolution = the dgrad launch of the tensor-core
    path; 16/36 of the FLOPs of the x channels).  The skip half accumulates onto it and emits the BatchNorm statistics
    of the sum.  Backward: dx = conv4x4_s2(dz, W4) at LOW resolution (no 2x2 gradient-sum pass), dskip = dgrad3x3(dz,
    Ws), weight gradients of both halves merged into the [O,3,3,C1+C2] parameter gradient.  Falls back to the
    materialising ``upcat`` + ``conv_bn_act`` when a shape is not on the tensor-core path."""
    if not _upconv_ok(ctx, xin, skip, cp):
        return conv_bn_act(ctx, upcat(ctx, xin, skip), cp, bn, slope=slope)
    st = ctx.store
    C1 = xin.t.shape[-1]
    fold = not ctx.training and ctx.tape is None and ops.FOLD_BN_EVAL
    if fold:    # inference: BatchNorm folded into the weights, bias + activation in the epilogue of the last half
        if ctx.fold_key is None:
            ctx.fold_key = st.fold_key()
        w, b = st.folded(ctx.fold_key, cp, bn)
        wx, ws = ops.upconv_split_weights(w, C1)
        if skip is None:
            return Var(ops.upconv_fwd(xin.t, wx, bias=b, act_slope=slope))
        zx = ops.upconv_fwd(xin.t, wx)
        return Var(ops.conv_fwd_fused(skip.t, ws, b, slope, addend=zx))
    w = st.w(cp.weight, ctx.dtype)
    if ctx.tape is not None:
        wx, ws, w4, ws_ft = ops.upconv_split_weights(w, C1, backward=True)
    else:
        wx, ws = ops.upconv_split_weights(w, C1)
    sums, pre = None, None
    zx = ops.upconv_fwd(xin.t, wx) if skip is not None else None
    if skip is not None and _bn_fuse_ok(ctx, skip, cp):
        # skip half + BatchNorm + activation as ONE launch (the x half is its addend) when the tiles fit one wave's TMEM
        r = ops.conv_bn_act_fused(skip.t, ws, ctx.stats_slot(2 * cp.out_channels + 1, w.device), bn.weight, bn.bias,
                                  bn.running_mean, bn.running_var, bn.eps, bn.momentum, slope, 1, 1, addend=zx)
        if r is not None:
            z, pre = r[0], r[1:]
    if pre is None:
        sums = ctx.stats_slot(2 * cp.out_channels, w.device) if (ctx.training and ops.FUSE_BN_STATS) else None
        if skip is None:
            z = ops.upconv_fwd(xin.t, wx, bn_sums=sums)
        else:
            z = ops.conv_fwd_add(skip.t, ws, zx, bn_sums=sums)
    zv = Var(z)
    zv.sums = sums
    if ctx.tape is not None:
        xin.uses += 1
        if skip is not None:
            skip.uses += 1

        def bwd():
            dz = zv.g
            zv.g = None
            xin.uses -= 1
            if skip is not None:
                skip.uses -= 1
            if dz is None:
                return
            O = cp.out_channels

            def wgrads():
                C2 = skip.t.shape[-1] if skip is not None else 0
                buf = torch.zeros(O * (16 * C1 + 9 * C2), dtype=torch.float32, device=dz.device)     # one memset
                dw4 = buf[:O * 16 * C1].view(C1, 4, 4, O)
                ops.conv_wgrad(xin.t, dz, dw4, 2, 1)            # roles: "dy" = x (low resolution), "x" = dz
                dws = None
                if skip is not None:
                    dws = buf[O * 16 * C1:].view(O, 3, 3, C2)
                    ops.conv_wgrad(dz, skip.t, dws, 1, 1)
                ops.upconv_merge_wgrad(dw4, dws, st.g(cp.weight), C1)

            if ops.WGRAD_STREAM:
                with ctx.wgrad_stream(dz, xin.t, *([skip.t] if skip is not None else [])):
                    wgrads()
            else:
                wgrads()
            if isinstance(xin.g, torch.Tensor):
                raise RuntimeError("upconv: the low-resolution input is expected to have no earlier gradient on the tape")
            if xin.g is not False:
                xin.g = ops.conv_fwd(dz, w4, None, 2, 1)     # dx at low resolution
            if skip is not None and skip.g is not False:
                skip.g = ops.conv_dgrad(dz, ws, skip.t.shape, 1, 1, addend=skip.g, w_ft=ws_ft)
            ctx.done(cp.weight)
        ctx.tape.push(bwd)
    return bn_act(ctx, zv, bn, slope=slope, _pre=pre)


def conv_bn_act(ctx, xin, cp, bn, slope=0.0, residual=None):
    """a = act(BN(conv(x)) (+ residual)).  Training: convolution with the BatchNorm statistics in its epilogue, then the
    normalise + activation pass.  Eval mode on the tensor cores (inference, ``src/models/predict.py:113-130``): the
    BatchNorm is FOLDED into the weights / bias and the residual add + activation run in the convolution's epilogue —
    one launch, no BatchNorm pass, no intermediate tensor."""
    if (not ctx.training and ctx.tape is None and ctx.dtype == torch.bfloat16 and ops.USE_TC and ops.TC_PERSIST
            and ops.FOLD_BN_EVAL):
        st = ctx.store
        if ctx.fold_key is None:
            ctx.fold_key = st.fold_key()
        if xin.t is None:   # Cin = 3 stem
            H, W = xin.nchw.shape[2:]
            w, b = st.folded(ctx.fold_key, cp, bn)
            xs = ops.stem_pack_input(xin.nchw, cp.padding)
            return Var(ops.stem_fwd(xs, ops.stem_pack_weight(w), b, H, W, cp.kernel_size, cp.padding, act_slope=slope))
        B, H, W, Cin = xin.t.shape
        if ops.tc_supported(0, B, H, W, Cin, cp.out_channels, cp.kernel_size, cp.kernel_size, cp.stride, cp.padding):
            w, b = st.folded(ctx.fold_key, cp, bn)
            return Var(ops.conv_fwd_fused(xin.t, w, b, slope, addend=residual.t if residual is not None else None,
                                          stride=cp.stride, pad=cp.padding))
    if _bn_fuse_ok(ctx, xin, cp):
        r = ops.conv_bn_act_fused(xin.t, ctx.store.w(cp.weight, ctx.dtype), ctx.stats_slot(2 * cp.out_channels + 1, xin.t.device),
                                  bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, slope,
                                  cp.stride, cp.padding, residual=residual.t if residual is not None else None)
        if r is not None:   # ONE launch: conv, batch statistics, grid barrier, normalise (+ residual) + activation
            return bn_act(ctx, conv(ctx, xin, cp, _pre=r[0]), bn, slope=slope, residual=residual, _pre=r[1:])
    return bn_act(ctx, conv(ctx, xin, cp, bn=bn), bn, slope=slope, residual=residual)


def _bn_fuse_ok(ctx, xin, cp):
    """Training-mode conv + BatchNorm + activation as one launch (``ops.conv_bn_act_fused``): tensor-core path, no conv
    bias; whether the layer's tiles fit the tensor memory of one wave is decided by the launch itself."""
    return (ctx.training and ctx.dtype == torch.bfloat16 and ops.USE_TC and ops.TC_PERSIST and ops.FUSE_BN_STATS
            and ops.FUSE_BN_APPLY and xin.t is not None and cp.bias is None and xin.t.is_cuda)


def bn_act(ctx, zin, bn, slope=0.0, residual=None, _pre=None):
    """a = act(BN(z) (+ residual)); train mode uses batch statistics and updates the running ones.  ``_pre``: (a, mean,
    rstd, scale, shift) already computed by a fused conv + BatchNorm launch; only the backward is recorded."""
    st = ctx.store
    res_t = residual.t if residual is not None else None
    if _pre is not None:
        a, mean, rstd, scale, shift = _pre
        ctx._trackers.append(bn.num_batches_tracked)
    elif ctx.training and zin.sums is not None:
        a, mean, rstd, scale, shift = ops.bn_apply_fused(zin.t, zin.sums, bn.weight, bn.bias, bn.running_mean,
                                                         bn.running_var, bn.eps, bn.momentum, res_t, slope)
        ctx._trackers.append(bn.num_batches_tracked)
    else:
        if ctx.training:
            mean, rstd, scale, shift = ops.bn_stats(zin.t, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                    bn.eps, bn.momentum)
            ctx._trackers.append(bn.num_batches_tracked)
        else:
            scale, shift = ops.bn_eval_coeffs(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
            mean = rstd = None
        a = ops.bn_apply(zin.t, scale, shift, res_t, slope)
    out = Var(a)
    if ctx.tape is not None:
        if not ctx.training:
            raise RuntimeError("uda_b200: backward through eval-mode BatchNorm is not supported "
                               "(the reference never trains in eval mode)")
        if residual is not None:
            residual.uses += 1
        # the dgrad that completes dL/da may reduce this layer's backward sums in its epilogue: residual layers need
        # z for xhat (a - res is not the BN output), the others recover the pre-activation from a alone
        if slope != 1.0 or residual is None:
            out.bn_req = (zin.t if residual is not None else None, slope)

        def bwd():
            da = out.g
            out.g = None
            if residual is not None:
                residual.uses -= 1
            if da is None:
                return
            dres = None
            acc = False
            if residual is not None:
                if residual.g is None:
                    dres = torch.empty_like(residual.t)
                else:
                    dres, acc = residual.g, True
            # activation mask: non-residual layers recompute it from z (no read of `a`); residual ones need `a`
            use_a = slope != 1.0 and residual is not None
            zm = slope != 1.0 and residual is None
            if out.bwd_sums is not None:
                zin.g = ops.bn_bwd_apply_fused(da, zin.t, a if use_a else None, out.bwd_sums, residual is not None,
                                               bn.weight, bn.bias, mean, rstd, slope, st.g(bn.weight), st.g(bn.bias),
                                               dres=dres, dres_accumulate=acc,
                                               scale=scale if zm else None, shift=shift if zm else None)
                out.bwd_sums = None
            else:
                zin.g = ops.bn_bwd(da, zin.t, a if use_a else None, bn.weight, mean, rstd, slope,
                                   st.g(bn.weight), st.g(bn.bias), dres=dres, dres_accumulate=acc,
                                   scale=scale if zm else None, shift=shift if zm else None)
            if residual is not None:
                residual.g = dres
            ctx.done(bn.weight, bn.bias)
        ctx.tape.push(bwd)
    return out


def bias_act(ctx, zin, slope):
    """a = act(z) for a conv whose bias was already added in its epilogue (discriminator layer 1)."""
    a = ops.bias_act(zin.t, None, slope)
    out = Var(a)
    if ctx.tape is not None:
        zin.uses += 1

        def bwd():
            zin.uses -= 1
            if out.g is not None:
                zin.g = ops.act_bwd(out.g, a, slope)
            out.g = None
        ctx.tape.push(bwd)
    return out


def conv_bn_act_maxpool(ctx, xin, cp, bn, slope=0.0):
    """The ResNet stem: ``a = act(BN(conv(x)))`` and ``y = maxpool3x3s2(a)``; returns both (``a`` is the first skip
    connection).  Training on the bf16 path with the statistics from the conv epilogue: normalise + activation + pooling
    as ONE pass over z (``ops.bn_apply_maxpool_fused``); the backward is the two ops' own."""
    if ctx.training and ctx.dtype == torch.bfloat16 and ops.FUSE_BN_POOL and ops.FUSE_BN_STATS:
        zin = conv(ctx, xin, cp, bn=bn)
        r = None
        if zin.sums is not None:
            r = ops.bn_apply_maxpool_fused(zin.t, zin.sums, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                           bn.eps, bn.momentum, slope)
        if r is None:
            f = bn_act(ctx, zin, bn, slope=slope)
            return f, maxpool(ctx, f)
        f = bn_act(ctx, zin, bn, slope=slope, _pre=r[:5])
        return f, maxpool(ctx, f, _pre=r[5:])
    f = conv_bn_act(ctx, xin, cp, bn, slope=slope)
    return f, maxpool(ctx, f)


def maxpool(ctx, xin, _pre=None):
    y, idx = _pre if _pre is not None else ops.maxpool_fwd(xin.t)
    out = Var(y)
    if ctx.tape is not None:
        xin.uses += 1

        def bwd():
            xin.uses -= 1
            if out.g is not None:
                xin.g = ops.maxpool_bwd(out.g, idx, xin.t.shape, addend=xin.g)
            out.g = None
        ctx.tape.push(bwd)
    return out


def upcat(ctx, xin, skip):
    y = ops.upcat_fwd(xin.t, skip.t if skip is not None else None)
    out = Var(y)
    if ctx.tape is not None:
        C1 = xin.t.shape[-1]
        C2 = skip.t.shape[-1] if skip is not None else 0
        xin.uses += 1
        if skip is not None:
            skip.uses += 1

        def bwd():
            xin.uses -= 1
            if skip is not None:
                skip.uses -= 1
            if out.g is None:
                return
            dx, dskip = ops.upcat_bwd(out.g, C1, C2)
            out.g = None
            # `False` marks an input that needs no gradient (detached features given to the stand-alone decoder)
            if isinstance(xin.g, torch.Tensor) or (skip is not None and isinstance(skip.g, torch.Tensor)):
                raise RuntimeError("upcat: inputs are expected to have no earlier gradient on the tape")
            if xin.g is not False:
                xin.g = dx
            if skip is not None and skip.g is not False:
                skip.g = dskip
        ctx.tape.push(bwd)
    return out
