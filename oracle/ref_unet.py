"""Oracle restatement of ``smp.Unet(encoder_name, encoder_weights=None, in_channels, classes)``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  fp32 PyTorch, runs on CPU.

The reference creates its model through the third-party package
``segmentation_models_pytorch`` (reference ``requirements.txt:3`` ``>=0.3.0``;
call sites ``src/test_system.py:90-95``, ``src/models/train.py:572-577``,
``src/models/uda.py:42-48``).  That package is not vendored and not installed, so
the published architecture is restated here:

  encoder  = torchvision ResNet without avgpool/fc, returning 6 feature maps
             [x, relu(bn1(conv1 x)), layer1(maxpool .), layer2, layer3, layer4]
  decoder  = 5 blocks: nearest x2 upsample -> cat(skip) -> (conv3x3 no-bias, BN, ReLU) x2
             in/skip/out = 512/256/256, 256/128/128, 128/64/64, 64/64/32, 32/0/16  (resnet34)
  head     = conv3x3(16 -> classes) with bias, no activation

Structure, BN eps/momentum, bias presence and the nearest upsample are the ones
visible in the TorchScript graphs the reference logged under ``test_logs/``
(SURVEY.md T1/T2).  state_dict keys follow smp naming.
**Parity unpinned** numerically: the reference holds no golden value for the U-Net.
"""
import math
import torch
import torch.nn as nn
import torch.nn.functional as F

_CFG = {
    # name: (block, layers, out_channels)
    "resnet18": ("basic", (2, 2, 2, 2), (3, 64, 64, 128, 256, 512)),
    "resnet34": ("basic", (3, 4, 6, 3), (3, 64, 64, 128, 256, 512)),
    "resnet50": ("bottleneck", (3, 4, 6, 3), (3, 64, 256, 512, 1024, 2048)),
}


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or cin != planes:
            self.downsample = nn.Sequential(
                nn.Conv2d(cin, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or cin != planes * 4:
            self.downsample = nn.Sequential(
                nn.Conv2d(cin, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + idt)


class RefResNetEncoder(nn.Module):
    def __init__(self, name="resnet34", in_channels=3):
        super().__init__()
        kind, layers, out_channels = _CFG[name]
        block = BasicBlock if kind == "basic" else Bottleneck
        self.out_channels = (in_channels,) + tuple(out_channels[1:])
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        cin = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            blocks = []
            for bi in range(n):
                stride = 2 if (bi == 0 and li > 1) else 1
                blocks.append(block(cin, planes, stride))
                cin = planes * block.expansion
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        # torchvision ResNet default init
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        f0 = x
        f1 = self.relu(self.bn1(self.conv1(x)))
        f2 = self.layer1(self.maxpool(f1))
        f3 = self.layer2(f2)
        f4 = self.layer3(f3)
        f5 = self.layer4(f4)
        return [f0, f1, f2, f3, f4, f5]


def _conv_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class RefDecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _conv_bn_relu(cin + cskip, cout)
        self.conv2 = _conv_bn_relu(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class RefUnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        head = enc[0]
        in_ch = [head] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.blocks = nn.ModuleList(
            [RefDecoderBlock(i, s, o) for i, s, o in zip(in_ch, skip_ch, decoder_channels)])
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, *features):
        feats = features[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class RefUnet(nn.Module):
    """fp32 restatement of smp.Unet; see module docstring."""

    def __init__(self, encoder_name="resnet34", encoder_weights=None, in_channels=3, classes=24,
                 decoder_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        if encoder_weights is not None:
            raise ValueError("oracle: pretrained encoder weights are unavailable offline")
        self.encoder = RefResNetEncoder(encoder_name, in_channels)
        self.decoder = RefUnetDecoder(self.encoder.out_channels, decoder_channels)
        head = nn.Conv2d(decoder_channels[-1], classes, 3, padding=1)
        nn.init.xavier_uniform_(head.weight)
        nn.init.constant_(head.bias, 0)
        self.segmentation_head = nn.Sequential(head, nn.Identity(), nn.Identity())

    def forward(self, x):
        return self.segmentation_head(self.decoder(*self.encoder(x)))


def emulate_bf16(model):
    """Return a copy of ``model`` that rounds to bfloat16 at the points where the bf16 compute path
    stores tensors (network input, conv weights, every conv output, every post-activation tensor, the
    downsample-branch BatchNorm output), while all arithmetic stays fp32 — i.e. the reference PyTorch
    path "in bf16".  BatchNorm statistics are therefore taken from the rounded conv outputs, exactly as
    the CUDA path does.  The head's logits stay fp32 (the CUDA head writes fp32 accumulators).

    Random-init train-mode U-Nets amplify storage rounding strongly (bf16-emulated vs fp32 logits differ
    by ~1e-1 at batch 2), so the 2e-2 bf16 parity gate is checked against THIS model, and the distance
    to the fp32 oracle is reported alongside (DESIGN.md "bf16 parity").
    """
    import copy
    m = copy.deepcopy(model)

    def rnd(_mod, _inp, out):
        return out.bfloat16().float()

    def rnd_in(_mod, inp):
        return tuple(t.bfloat16().float() for t in inp)

    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                mod.weight.copy_(mod.weight.bfloat16().float())
    head = m.segmentation_head[0]
    for name, mod in m.named_modules():
        if isinstance(mod, nn.Conv2d) and mod is not head:
            mod.register_forward_hook(rnd)
        elif isinstance(mod, nn.ReLU):
            mod.register_forward_hook(rnd)
        elif isinstance(mod, nn.BatchNorm2d) and name.endswith("downsample.1"):
            mod.register_forward_hook(rnd)
    m.encoder.register_forward_pre_hook(rnd_in)
    return m
