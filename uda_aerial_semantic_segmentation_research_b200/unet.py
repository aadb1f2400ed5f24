"""B200-native U-Net behind the reference's model-creation entry point.

The reference creates its model with ``smp.Unet(encoder_name=..., encoder_weights=..., in_channels=3,
classes=C)`` (``src/test_system.py:90-95``, ``src/models/train.py:572-577``, ``src/models/uda.py:42-48``)
and then only uses ``.to/.train/.eval/.parameters/.state_dict/.load_state_dict``, ``forward(x)``,
``.encoder(x)`` -> 6 feature maps, ``.encoder.out_channels``, ``.decoder(*features)`` and
``.segmentation_head``.  ``Unet`` below keeps exactly that surface and smp's
state_dict keys, while its forward/backward is one autograd node that runs the hand-written sm_100a
kernels on NHWC bf16 (or fp32 "parity mode") activations.
"""
import math
import torch
import torch.nn as nn

from . import ops
from . import engine as E
from .engine import ConvParams, BNParams, Var, Tape, Ctx, ParamStore

_CFG = {
    "resnet18": ("basic", (2, 2, 2, 2), (3, 64, 64, 128, 256, 512)),
    "resnet34": ("basic", (3, 4, 6, 3), (3, 64, 64, 128, 256, 512)),
    "resnet50": ("bottleneck", (3, 4, 6, 3), (3, 64, 256, 512, 1024, 2048)),
}


class _Seq(nn.Sequential):
    """Sequential container used purely for smp-compatible key names (never called as a layer)."""

    def forward(self, *a, **k):
        raise RuntimeError("executed by the uda_b200 engine (call the owning network)")


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = ConvParams(cin, planes, 3, stride, 1)
        self.bn1 = BNParams(planes)
        self.conv2 = ConvParams(planes, planes, 3, 1, 1)
        self.bn2 = BNParams(planes)
        self.downsample = None
        if stride != 1 or cin != planes:
            self.downsample = _Seq(ConvParams(cin, planes, 1, stride, 0), BNParams(planes))

    def run(self, ctx, x):
        idt = x
        if self.downsample is not None:
            idt = E.conv_bn_act(ctx, x, self.downsample[0], self.downsample[1], slope=1.0)
        y = E.conv_bn_act(ctx, x, self.conv1, self.bn1, slope=0.0)
        return E.conv_bn_act(ctx, y, self.conv2, self.bn2, slope=0.0, residual=idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = ConvParams(cin, planes, 1, 1, 0)
        self.bn1 = BNParams(planes)
        self.conv2 = ConvParams(planes, planes, 3, stride, 1)
        self.bn2 = BNParams(planes)
        self.conv3 = ConvParams(planes, planes * 4, 1, 1, 0)
        self.bn3 = BNParams(planes * 4)
        self.downsample = None
        if stride != 1 or cin != planes * 4:
            self.downsample = _Seq(ConvParams(cin, planes * 4, 1, stride, 0), BNParams(planes * 4))

    def run(self, ctx, x):
        idt = x
        if self.downsample is not None:
            idt = E.conv_bn_act(ctx, x, self.downsample[0], self.downsample[1], slope=1.0)
        y = E.conv_bn_act(ctx, x, self.conv1, self.bn1, slope=0.0)
        y = E.conv_bn_act(ctx, y, self.conv2, self.bn2, slope=0.0)
        return E.conv_bn_act(ctx, y, self.conv3, self.bn3, slope=0.0, residual=idt)


class ResNetEncoder(nn.Module):
    """ResNet feature extractor with smp's key names (``conv1, bn1, layer1..4``) and ``out_channels``."""

    def __init__(self, name="resnet34", in_channels=3):
        super().__init__()
        if name not in _CFG:
            raise ValueError(f"unsupported encoder {name!r}; available: {sorted(_CFG)}")
        kind, layers, out_channels = _CFG[name]
        block = BasicBlock if kind == "basic" else Bottleneck
        self.out_channels = (in_channels,) + tuple(out_channels[1:])
        self.conv1 = ConvParams(in_channels, 64, 7, 2, 3)
        self.bn1 = BNParams(64)
        cin = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            blocks = []
            for bi in range(n):
                blocks.append(block(cin, planes, 2 if (bi == 0 and li > 1) else 1))
                cin = planes * block.expansion
            setattr(self, f"layer{li}", _Seq(*blocks))
        for m in self.modules():  # torchvision ResNet default initialisation
            if isinstance(m, ConvParams):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def run(self, ctx, x):
        f1, y = E.conv_bn_act_maxpool(ctx, x, self.conv1, self.bn1, slope=0.0)
        feats = [x, f1]
        for li in range(1, 5):
            for blk in getattr(self, f"layer{li}"):
                y = blk.run(ctx, y)
            feats.append(y)
        return feats

    def forward(self, x):
        """Stand-alone use (``model.encoder(x)``, reference src/models/domain_model.py:52-53, uda.py:64,78):
        returns the 6 feature maps as fp32 NCHW tensors."""
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise RuntimeError("encoder must be owned by a uda_b200 Unet")
        return owner()._encoder_standalone(x)


class DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Seq(ConvParams(cin + cskip, cout, 3, 1, 1), BNParams(cout), nn.Identity())
        self.conv2 = _Seq(ConvParams(cout, cout, 3, 1, 1), BNParams(cout), nn.Identity())

    def run(self, ctx, x, skip):
        y = E.upconv_bn_act(ctx, x, skip, self.conv1[0], self.conv1[1], slope=0.0)
        return E.conv_bn_act(ctx, y, self.conv2[0], self.conv2[1], slope=0.0)


class UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.blocks = nn.ModuleList(
            [DecoderBlock(i, s, o) for i, s, o in zip(in_ch, skip_ch, decoder_channels)])
        for m in self.modules():  # smp decoder initialisation
            if isinstance(m, ConvParams):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")

    def run(self, ctx, feats):
        feats = feats[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk.run(ctx, x, skips[i] if i < len(skips) else None)
        return x

    def forward(self, *features):
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise RuntimeError("decoder must be owned by a uda_b200 Unet")
        return owner()._decoder_standalone(*features)


class _UnetFn(torch.autograd.Function):
    """Network-level autograd node: (inputs..., *params) -> outputs (NCHW fp32 at the edge)."""

    @staticmethod
    def forward(ctx, net, part, n_in, record, *args):
        inputs = args[:n_in]
        outs, tape, in_vars, out_vars = net._run(part, inputs, record=record)
        ctx.net, ctx.tape, ctx.in_vars, ctx.out_vars = net, tape, in_vars, out_vars
        ctx.n_in = n_in
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        net, tape = ctx.net, ctx.tape
        if tape is None:
            raise RuntimeError("uda_b200 Unet: backward called on a graph recorded without gradients")
        st = net._store
        st.new_grad()
        dtype = net.compute_dtype
        in_need = ctx.needs_input_grad[4:4 + ctx.n_in]
        for v, g in zip(ctx.out_vars, grads):
            if g is not None and v is not None:
                v.g = ops.nchw_to_nhwc(g.contiguous().float(), dtype)
        for v, need in zip(ctx.in_vars, in_need):
            if v is not None and not need:
                v.g = False
        if net._grad_sync is not None:
            net._grad_sync.begin(st)
        tape.backward()
        if net._grad_sync is not None:
            net._grad_sync.end(st)  # waits for the bucketed all-reduces issued during the tape replay
        gin = []
        for v, need in zip(ctx.in_vars, in_need):
            if need and v is not None and isinstance(v.g, torch.Tensor):
                gin.append(ops.nhwc_to_nchw(v.g))
            else:
                gin.append(None)
        ctx.tape = ctx.in_vars = ctx.out_vars = None
        return (None, None, None, None) + tuple(gin) + tuple(st.grad_views())


class SegmentationHead(_Seq):
    def __init__(self, cin, classes):
        conv = ConvParams(cin, classes, 3, 1, 1, bias=True)
        nn.init.xavier_uniform_(conv.weight)
        super().__init__(conv, nn.Identity(), nn.Identity())

    def forward(self, x):
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise RuntimeError("segmentation_head must be owned by a uda_b200 Unet")
        return owner()._head_standalone(x)


class Unet(nn.Module):
    """``smp.Unet``-compatible model (resnet18/34/50 encoders, nearest-upsampling decoder, 3x3 head).

    ``compute_dtype``: ``torch.bfloat16`` (tcgen05 tensor-core convolutions, fp32 accumulate; logits
    within 2e-2 of the fp32 oracle) or ``torch.float32`` (FP32-pipe parity mode, 1e-4).
    """

    def __init__(self, encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
                 decoder_use_batchnorm=True, decoder_channels=(256, 128, 64, 32, 16),
                 decoder_attention_type=None, in_channels=3, classes=1, activation=None, aux_params=None,
                 compute_dtype=torch.bfloat16):
        super().__init__()
        if encoder_weights is not None:
            raise ValueError("pretrained encoder weights are not available offline; pass encoder_weights=None "
                             "and load a checkpoint with load_state_dict (smp key names are kept)")
        if encoder_depth != 5 or not decoder_use_batchnorm or decoder_attention_type or activation or aux_params:
            raise NotImplementedError("only the configuration the reference uses is implemented "
                                      "(depth 5, batch-norm decoder, no attention/activation/aux head)")
        self.compute_dtype = compute_dtype
        self.encoder = ResNetEncoder(encoder_name, in_channels)
        self.decoder = UnetDecoder(self.encoder.out_channels, decoder_channels)
        self.segmentation_head = SegmentationHead(decoder_channels[-1], classes)
        self.classes = classes
        self.name = f"u-{encoder_name}"
        self._store = ParamStore(self)
        self._grad_sync = None
        import weakref
        ref = weakref.ref(self)
        for m in (self.encoder, self.decoder, self.segmentation_head):
            object.__setattr__(m, "_owner", ref)

    # -- execution --------------------------------------------------------------------------
    def _prepare(self, device):
        if device.type != "cuda":
            raise RuntimeError("uda_b200.Unet runs on CUDA (sm_100a) only — there is no CPU fallback; "
                               "move the model and its inputs to a B200 device")
        self._store.ensure_flat(device)
        if self.compute_dtype == torch.bfloat16:
            self._store.refresh_shadow()

    def _run(self, part, inputs, record):
        """part: 'full' (x -> logits) | 'encoder' (x -> 6 feats) | 'decoder' (6 feats -> dec) | 'head'."""
        dtype = self.compute_dtype
        tape = Tape() if record else None
        ctx = Ctx(self._store, dtype, self.training, tape)
        ctx.sync = self._grad_sync
        if part in ("full", "encoder"):
            x = inputs[0]
            B, C, H, W = x.shape
            if H % 32 or W % 32:
                raise RuntimeError(f"Unet: input height and width must be divisible by 32, got {H}x{W}")
            xin = E.input_var(x, dtype, self.encoder.conv1, record and x.requires_grad)
            in_vars = [xin]
            feats = self.encoder.run(ctx, xin)
            if part == "encoder":
                ctx.finish_forward()
                outs = [ops.nhwc_to_nchw(f.t) for f in feats[1:]]
                return outs, tape, in_vars, feats[1:]
            dec = self.decoder.run(ctx, feats)
            logits = E.conv(ctx, dec, self.segmentation_head[0], nchw_out=True)
            ctx.finish_forward()
            return [logits.t], tape, in_vars, [logits]
        if part == "decoder":
            in_vars = [None] + [Var(ops.nchw_to_nhwc(f.contiguous().float(), dtype)) for f in inputs[1:]]
            dec = self.decoder.run(ctx, in_vars)
            ctx.finish_forward()
            return [ops.nhwc_to_nchw(dec.t)], tape, in_vars, [dec]
        if part == "head":
            xin = Var(ops.nchw_to_nhwc(inputs[0].contiguous().float(), dtype))
            logits = E.conv(ctx, xin, self.segmentation_head[0], nchw_out=True)
            return [logits.t], tape, [xin], [logits]
        raise ValueError(part)

    def _call(self, part, *inputs):
        self._prepare(inputs[-1].device)
        record = torch.is_grad_enabled() and (any(p.requires_grad for p in self._store.params)
                                              or any(t.requires_grad for t in inputs))
        return _UnetFn.apply(self, part, len(inputs), record, *inputs, *self._store.params)

    def forward(self, x):
        return self._call("full", x)

    def _encoder_standalone(self, x):
        return [x] + list(self._call("encoder", x))

    def _decoder_standalone(self, *features):
        return self._call("decoder", *features)

    def _head_standalone(self, x):
        return self._call("head", x)

    @torch.no_grad()
    def predict(self, x):
        """smp's ``model.predict``: eval-mode forward without gradients."""
        was = self.training
        self.eval()
        try:
            return self.forward(x)
        finally:
            self.train(was)


def create_model(model_name="Unet", encoder_name="resnet34", encoder_weights=None, in_channels=3, classes=24, **kw):
    """``getattr(smp, Config.MODEL_NAME)(...)`` drop-in (reference src/models/train.py:572-577)."""
    if model_name != "Unet":
        raise NotImplementedError(f"only Unet is on the reference's hot path, got {model_name!r}")
    return Unet(encoder_name=encoder_name, encoder_weights=encoder_weights, in_channels=in_channels,
                classes=classes, **kw)
