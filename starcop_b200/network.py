"""Parameter container of the HyperSTARCOP network with the state_dict layout of
``segmentation_models_pytorch.Unet(encoder_name="mobilenet_v2")``, so
checkpoints written by the reference (``model.pt`` / Lightning ``.ckpt``) load unchanged.

The modules below hold parameters and buffers only -- they never run: ``forward`` hands the whole
network to the CUDA engine (``engine.UNetEngine``).  Construction order and initialisers follow
torchvision's MobileNetV2 and smp's ``initialize_decoder`` / ``initialize_head``.
"""
import torch
import torch.nn as nn

from .engine import DECODER_CHANNELS, MBV2_SETTING


def _conv_bn(cin, cout, k=3, stride=1, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class _InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, t):
        super().__init__()
        hidden = int(round(inp * t))
        layers = [] if t == 1 else [_conv_bn(inp, hidden, 1)]
        layers += [_conv_bn(hidden, hidden, 3, stride, groups=hidden), nn.Conv2d(hidden, oup, 1, 1, 0, bias=False),
                   nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*layers)


class _Encoder(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        feats = [_conv_bn(3, 32, 3, 2)]
        cin = 32
        for t, c, n, s in MBV2_SETTING:
            for i in range(n):
                feats.append(_InvertedResidual(cin, c, s if i == 0 else 1, t))
                cin = c
        feats.append(_conv_bn(cin, 1280, 1))
        self.features = nn.Sequential(*feats)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if in_channels != 3:      # smp patch_first_conv without pretrained weights
            conv = self.features[0][0]
            conv.in_channels = in_channels
            conv.weight = nn.Parameter(torch.empty(conv.out_channels, in_channels, 3, 3))
            conv.reset_parameters()


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        enc = (1280, 96, 32, 24, 16)
        in_ch = [enc[0]] + list(DECODER_CHANNELS[:-1])
        skip = list(enc[1:]) + [0]
        self.blocks = nn.ModuleList(_DecoderBlock(i + s, o) for i, s, o in zip(in_ch, skip, DECODER_CHANNELS))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class UnetParameters(nn.Module):
    """encoder / decoder / segmentation_head parameter tree (names == smp.Unet's)."""

    def __init__(self, in_channels, classes=1):
        super().__init__()
        assert classes == 1, "the STARCOP hot path is binary segmentation (num_classes: 1)"
        self.encoder = _Encoder(in_channels)
        self.decoder = _Decoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1),
                                               nn.Identity(), nn.Identity())
        nn.init.xavier_uniform_(self.segmentation_head[0].weight)
        nn.init.constant_(self.segmentation_head[0].bias, 0)
        self.in_channels = in_channels
