"""Restatement of ``segmentation_models_pytorch.Unet(encoder_name="mobilenet_v2",
encoder_weights=None, in_channels=C, classes=1, activation=None)``.

The reference builds this network in ``starcop/models/model_module.py:238-251``; the
arithmetic lives in the third-party package segmentation_models_pytorch (unpinned in
``requirements.txt:9``; 0.3.x era) on top of ``torchvision.models.MobileNetV2``.
Module / attribute names are kept identical so state_dict keys match
(``encoder.features.*``, ``decoder.blocks.{i}.conv{1,2}.{0,1}.*``,
``segmentation_head.0.*`` -- SURVEY.md Appendix A.4).

Test infrastructure only (see ``oracle/__init__.py``).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# torchvision.models.mobilenetv2: inverted_residual_setting (t, c, n, s)
MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
DECODER_CHANNELS = (256, 128, 64, 32, 16)


class ConvBNReLU6(nn.Sequential):
    """torchvision ``Conv2dNormActivation(norm=BatchNorm2d, act=ReLU6)``: keys 0=conv, 1=bn."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, groups=1):
        pad = (kernel_size - 1) // 2
        super().__init__(nn.Conv2d(cin, cout, kernel_size, stride, pad, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    """torchvision ``InvertedResidual``: self.conv = [expand?] + depthwise + project conv + bn."""

    def __init__(self, inp, oup, stride, expand_ratio):
        super().__init__()
        hidden = int(round(inp * expand_ratio))
        self.use_res_connect = stride == 1 and inp == oup
        layers = []
        if expand_ratio != 1:
            layers.append(ConvBNReLU6(inp, hidden, kernel_size=1))
        layers.extend([ConvBNReLU6(hidden, hidden, stride=stride, groups=hidden),
                       nn.Conv2d(hidden, oup, 1, 1, 0, bias=False),
                       nn.BatchNorm2d(oup)])
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv(x) if self.use_res_connect else self.conv(x)


class MobileNetV2Encoder(nn.Module):
    """smp ``MobileNetV2Encoder`` (torchvision MobileNetV2 minus classifier, depth 5)."""

    def __init__(self, in_channels=3):
        super().__init__()
        feats = [ConvBNReLU6(3, 32, stride=2)]
        cin = 32
        for t, c, n, s in MBV2_SETTING:
            for i in range(n):
                feats.append(InvertedResidual(cin, c, s if i == 0 else 1, t))
                cin = c
        feats.append(ConvBNReLU6(cin, 1280, kernel_size=1))
        self.features = nn.Sequential(*feats)
        # torchvision init
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        self.out_channels = (in_channels, 16, 24, 32, 96, 1280)
        if in_channels != 3:
            # smp ``patch_first_conv(pretrained=False)``: fresh weight + reset_parameters()
            conv = self.features[0][0]
            conv.in_channels = in_channels
            conv.weight = nn.Parameter(torch.empty(conv.out_channels, in_channels, 3, 3))
            conv.reset_parameters()

    def stages(self):
        f = self.features
        return [nn.Identity(), f[:2], f[2:4], f[4:7], f[7:14], f[14:]]

    def forward(self, x):
        out = []
        for st in self.stages():
            x = st(x)
            out.append(x)
        return out


class Conv2dReLU(nn.Sequential):
    """smp ``modules.Conv2dReLU`` with use_batchnorm=True: keys 0=conv(bias=False), 1=bn, 2=relu."""

    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True))


class DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout)
        self.conv2 = Conv2dReLU(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)          # upsampled first, skip second
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=DECODER_CHANNELS):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]        # (1280, 96, 32, 24, 16)
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.blocks = nn.ModuleList(DecoderBlock(i, s, o)
                                    for i, s, o in zip(in_ch, skip_ch, decoder_channels))
        for m in self.modules():                      # smp initialize_decoder
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, *features):
        features = features[1:][::-1]
        x, skips = features[0], features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class Unet(nn.Module):
    """``smp.Unet(mobilenet_v2)``; ``forward`` requires H, W divisible by 32 like smp."""

    def __init__(self, encoder_name="mobilenet_v2", encoder_weights=None, in_channels=3,
                 classes=1, activation=None):
        super().__init__()
        assert encoder_name == "mobilenet_v2" and activation is None
        assert encoder_weights is None, "no network here: imagenet weights are unavailable"
        self.encoder = MobileNetV2Encoder(in_channels)
        self.decoder = UnetDecoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1),
                                               nn.Identity(), nn.Identity())
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)   # smp initialize_head
        nn.init.constant_(self.segmentation_head[0].bias, 0)

    def forward(self, x):
        h, w = x.shape[-2:]
        if h % 32 or w % 32:
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected image height and "
                               f"width divisible by 32.")
        return self.segmentation_head(self.decoder(*self.encoder(x)))


def count_parameters(m):
    return sum(p.numel() for p in m.parameters())
