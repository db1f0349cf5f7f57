"""MobileNetV2 layer graph at 32x32 (reference models/mobilenet.py): unquantized 3x3/s2 stem, 17
inverted-residual blocks (quantized pointwise 1x1 expand -> quantized depthwise 3x3 -> quantized
pointwise 1x1 project, ReLU6 after the first two), unquantized 1x1 head conv to 1280, pool, linear.
Supplies the depthwise / pointwise shapes of BASELINE.json configs[2]."""
import math

import torch.nn as nn

# (expansion t, out channels c, repeats n, first stride s) -- Sandler et al. 2018, table 2
_STAGES = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def _bn_act(c, act=True, fused=False):
    """[norm, ReLU6]; with the fused norm the ReLU6 happens inside the norm kernel and an nn.Identity
    keeps the reference's nn.Sequential indices (and with them the state_dict keys)"""
    if fused:
        from po2_quantization_b200 import FusedSyncBatchNorm
        return [FusedSyncBatchNorm(c, act="relu6" if act else None)] + ([nn.Identity()] if act else [])
    return [nn.SyncBatchNorm(c)] + ([nn.ReLU6(inplace=True)] if act else [])


class _InvRes(nn.Module):
    def __init__(self, conv_cls, c_in, c_out, stride, t, quantize_fn, bits, fused=False):
        super().__init__()
        hid = round(c_in * t)
        self.identity = stride == 1 and c_in == c_out
        q = dict(bias=False, quantize_fn=quantize_fn, bits=bits)
        layers = []
        if t != 1:
            layers += [conv_cls(c_in, hid, 1, 1, 0, **q)] + _bn_act(hid, fused=fused)
        layers += [conv_cls(hid, hid, 3, stride, 1, groups=hid, **q)] + _bn_act(hid, fused=fused)
        layers += [conv_cls(hid, c_out, 1, 1, 0, **q)] + _bn_act(c_out, act=False, fused=fused)
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv(x) if self.identity else self.conv(x)


class MobileNetV2Cifar(nn.Module):
    def __init__(self, conv_cls, num_classes, quantize_fn, bits, fused=False):
        super().__init__()
        c_in = 32
        feats = [nn.Sequential(nn.Conv2d(3, c_in, 3, 2, 1, bias=False), *_bn_act(c_in, fused=fused))]
        for t, c, n, s in _STAGES:
            for i in range(n):
                feats.append(_InvRes(conv_cls, c_in, c, s if i == 0 else 1, t, quantize_fn, bits, fused))
                c_in = c
        self.features = nn.Sequential(*feats)
        self.conv = nn.Sequential(nn.Conv2d(c_in, 1280, 1, 1, 0, bias=False), *_bn_act(1280, fused=fused))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.classifier = nn.Linear(1280, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    def forward(self, x):
        x = self.conv(self.features(x))
        return self.classifier(self.avgpool(x).flatten(1))


def mobilenet_v2_cifar(num_classes: int = 10, quantize_fn=None, bits: int = 4, conv_cls=None, fused_norm=None):
    """fused_norm: FusedSyncBatchNorm (norm + ReLU6 in one kernel) instead of nn.SyncBatchNorm + nn.ReLU6;
    default: on with this repo's conv class, off when another implementation's class is passed"""
    if fused_norm is None:
        fused_norm = conv_cls is None
    if conv_cls is None:
        from po2_quantization_b200 import QuantizedConv2d as conv_cls
    return MobileNetV2Cifar(conv_cls, num_classes, quantize_fn, bits, fused_norm)
