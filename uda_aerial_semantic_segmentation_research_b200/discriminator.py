"""Image-level ``DomainDiscriminator`` of the reference (``src/models/discriminator.py:4-55``) on the
uda_b200 kernels: conv4x4 s2 (+bias) -> LeakyReLU(0.2); three conv4x4 s2 (+bias) + BatchNorm +
LeakyReLU(0.2); global average pool; Linear(512,1); Sigmoid.  Same constructor, state_dict keys
(``features.{0,2,3,5,6,8,9}``, ``classifier.2``), default PyTorch initialisation, ``[B,1]`` output.
"""
import math
import torch
import torch.nn as nn

from . import ops
from . import engine as E
from .engine import ConvParams, BNParams, LinearParams, Var, Tape, Ctx, ParamStore


class _Holder(nn.Sequential):
    def forward(self, *a, **k):
        raise RuntimeError("executed by the uda_b200 engine (call the owning network)")


def _default_conv_init(conv):
    # nn.Conv2d.reset_parameters: kaiming_uniform_(a=sqrt(5)) + uniform bias in +-1/sqrt(fan_in)
    nn.init.kaiming_uniform_(conv.weight, a=math.sqrt(5))
    if conv.bias is not None:
        fan_in = conv.in_channels * conv.kernel_size * conv.kernel_size
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(conv.bias, -bound, bound)


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, record, x, *params):
        y, tape, xin, state = net._run(x, record)
        ctx.net, ctx.tape, ctx.xin, ctx.state = net, tape, xin, state
        return y

    @staticmethod
    def backward(ctx, gy):
        net, tape, st = ctx.net, ctx.tape, ctx.net._store
        if tape is None:
            raise RuntimeError("DomainDiscriminator: backward called on a graph recorded without gradients")
        st.new_grad()
        if net._grad_sync is not None:
            net._grad_sync.begin(st)
        need_x = ctx.needs_input_grad[2]
        if not need_x:
            ctx.xin.g = False
        feat, y, pooled = ctx.state
        lin = net.classifier[2]
        feat.g = ops.gap_linear_sigmoid_bwd(gy.contiguous().float(), y, pooled, st.w2d(lin.weight), st.g(lin.weight),
                                            st.g(lin.bias), feat.t.shape, feat.t.dtype)
        if net._grad_sync is not None:
            net._grad_sync.param_done(st, lin.weight)
            net._grad_sync.param_done(st, lin.bias)
        tape.backward()
        if net._grad_sync is not None:
            net._grad_sync.end(st)
        gx = ops.nhwc_to_nchw(ctx.xin.g) if (need_x and isinstance(ctx.xin.g, torch.Tensor)) else None
        ctx.tape = ctx.xin = ctx.state = None
        return (None, None, gx) + tuple(st.grad_views())


class DomainDiscriminator(nn.Module):
    def __init__(self, input_channels=3, compute_dtype=torch.bfloat16):
        super().__init__()
        self.compute_dtype = compute_dtype
        layers = [ConvParams(input_channels, 64, 4, 2, 1, bias=True), nn.Identity()]
        for cin in (64, 128, 256):
            layers += [ConvParams(cin, cin * 2, 4, 2, 1, bias=True), BNParams(cin * 2), nn.Identity()]
        self.features = _Holder(*layers)
        self.classifier = _Holder(nn.Identity(), nn.Identity(), LinearParams(512, 1), nn.Identity())
        for m in self.features:
            if isinstance(m, ConvParams):
                _default_conv_init(m)
        lin = self.classifier[2]
        nn.init.kaiming_uniform_(lin.weight, a=math.sqrt(5))
        nn.init.uniform_(lin.bias, -1 / math.sqrt(512), 1 / math.sqrt(512))
        self._store = _DiscStore(self)
        self._grad_sync = None

    def _run(self, x, record, xin=None):
        """``xin``: an already prepared channels-last input Var (output-space use: softmax probabilities, channel
        dimension possibly zero-padded to a tensor-core atom) instead of the NCHW image ``x``."""
        dtype = self.compute_dtype
        tape = Tape() if record else None
        ctx = Ctx(self._store, dtype, self.training, tape)
        ctx.sync = self._grad_sync
        f = self.features
        if xin is None:
            xin = E.input_var(x, dtype, f[0], record and x.requires_grad)
            z0 = E.conv(ctx, xin, f[0])
        elif xin.t.shape[-1] != f[0].in_channels:
            z0 = E.conv_padded_cin(ctx, xin, f[0])
        else:
            z0 = E.conv(ctx, xin, f[0])
        y = E.bias_act(ctx, z0, 0.2)
        for ci, bi in ((2, 3), (5, 6), (8, 9)):
            y = E.conv_bn_act(ctx, y, f[ci], f[bi], slope=0.2)
        ctx.finish_forward()
        lin = self.classifier[2]
        out, pooled = ops.gap_linear_sigmoid_fwd(y.t, self._store.w2d(lin.weight), lin.bias)
        return out, tape, xin, (y, out, pooled)

    def _prepare(self, device):
        if device.type != "cuda":
            raise RuntimeError("uda_b200.DomainDiscriminator runs on CUDA (sm_100a) only — no CPU fallback")
        self._store.ensure_flat(device)
        if self.compute_dtype == torch.bfloat16:
            self._store.refresh_shadow()

    def forward(self, x):
        self._prepare(x.device)
        record = torch.is_grad_enabled() and (any(p.requires_grad for p in self._store.params) or x.requires_grad)
        return _DiscFn.apply(self, record, x, *self._store.params)


class _OutputSpaceFn(torch.autograd.Function):
    """logits -> D(GRL(softmax(logits))) as one autograd node.  Forward: fused softmax + NCHW->NHWC bf16 pack, then the
    discriminator's layers.  Backward: the discriminator's tape, then ONE pass that applies the softmax Jacobian and
    the gradient-reversal factor -alpha (reference ``GradientReverseFunction``, src/models/uda.py:103-112) and writes
    the fp32 NCHW logit gradient."""

    @staticmethod
    def forward(ctx, net, record, alpha, logits, *params):
        z = logits.contiguous().float()
        C = z.shape[1]
        cpad = 8 if C <= 8 else (16 if C <= 16 else 32)
        probs = ops.softmax_nhwc(z, cpad)
        xin = Var(probs)
        y, tape, xin, state = net._run(None, record, xin=xin)
        ctx.net, ctx.tape, ctx.xin, ctx.state = net, tape, xin, state
        ctx.alpha, ctx.C = float(alpha), C
        return y

    @staticmethod
    def backward(ctx, gy):
        net, tape, st = ctx.net, ctx.tape, ctx.net._store
        if tape is None:
            raise RuntimeError("OutputSpaceAdversary: backward called on a graph recorded without gradients")
        st.new_grad()
        if net._grad_sync is not None:
            net._grad_sync.begin(st)
        need_x = ctx.needs_input_grad[3]
        if not need_x:
            ctx.xin.g = False
        feat, y, pooled = ctx.state
        lin = net.classifier[2]
        feat.g = ops.gap_linear_sigmoid_bwd(gy.contiguous().float(), y, pooled, st.w2d(lin.weight), st.g(lin.weight),
                                            st.g(lin.bias), feat.t.shape, feat.t.dtype)
        if net._grad_sync is not None:
            net._grad_sync.param_done(st, lin.weight)
            net._grad_sync.param_done(st, lin.bias)
        tape.backward()
        if net._grad_sync is not None:
            net._grad_sync.end(st)
        gz = None
        if need_x and isinstance(ctx.xin.g, torch.Tensor):
            gz = ops.softmax_bwd_grl(ctx.xin.t, ctx.xin.g, -ctx.alpha, ctx.C)
        ctx.tape = ctx.xin = ctx.state = None
        return (None, None, None, gz) + tuple(st.grad_views())


class OutputSpaceAdversary(nn.Module):
    """Output-space domain adversary behind a gradient-reversal layer — the north-star's adversarial variant
    (BASELINE configs[2]: "output-space discriminator, gradient reversal"), i.e. the composition

        DomainDiscriminator(num_classes)(gradient_reverse_layer(softmax(logits, dim=1), alpha))

    of the reference's own building blocks (``src/models/uda.py:99-112``, ``src/models/discriminator.py:4-55``; the
    reference defines both and never wires them together).  One backward pass trains the discriminator
    to tell the domains apart and hands the segmentation network the REVERSED gradient, scaled by ``alpha``."""

    def __init__(self, discriminator, alpha=1.0):
        super().__init__()
        if discriminator.compute_dtype != torch.bfloat16:
            raise NotImplementedError("OutputSpaceAdversary runs on the bf16 tensor-core path")
        self.discriminator = discriminator
        self.alpha = alpha

    def forward(self, logits):
        d = self.discriminator
        if not logits.is_cuda:
            raise RuntimeError("OutputSpaceAdversary runs on CUDA (sm_100a) only — no CPU fallback")
        if logits.shape[1] != d.features[0].in_channels:
            raise ValueError(f"discriminator expects {d.features[0].in_channels} classes, logits have {logits.shape[1]}")
        d._prepare(logits.device)
        record = torch.is_grad_enabled() and (any(p.requires_grad for p in d._store.params) or logits.requires_grad)
        return _OutputSpaceFn.apply(d, record, self.alpha, logits, *d._store.params)


class _DiscStore(ParamStore):
    def w2d(self, p):
        off = self.offsets[id(p)]
        return self.flat[off:off + p.numel()].view(p.shape)
