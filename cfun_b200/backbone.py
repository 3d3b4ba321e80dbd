"""Pseudo-3D ResNet backbone (stages C1..C3) on the cfun_b200 CUDA ops.

Same class / factory names, constructor signatures and state_dict keys as reference backbone.py
(conv_S:14, conv_T:20, Bottleneck:26-114, P3D:117-158, P3D19:161); compute goes to ops.conv3d (+ fused frozen-BN /
ReLU / residual pass and the 2x2x2 max-pool kernel)."""
import math
import torch.nn as nn

from . import ops
from .layers import Conv3d, FrozenBatchNorm3d, Slot


def conv_S(in_planes, out_planes, stride=1, padding=1):
    """spatial 1x3x3 conv (backbone.py:14)"""
    return Conv3d(in_planes, out_planes, kernel_size=(1, 3, 3), stride=stride, padding=padding)


def conv_T(in_planes, out_planes, stride=1, padding=1):
    """temporal 3x1x1 conv (backbone.py:20)"""
    return Conv3d(in_planes, out_planes, kernel_size=(3, 1, 1), stride=stride, padding=padding)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, block, expand=False, stride=1, ST_structure=('A', 'B', 'C')):
        super().__init__()
        self.stride = stride
        self.expand = expand
        self.conv1 = Conv3d(inplanes, planes, kernel_size=1, stride=stride)
        self.bn1 = FrozenBatchNorm3d(planes)
        self.ST = list(ST_structure)[(block - 1) % len(ST_structure)]
        self.conv2 = conv_S(planes, planes, stride=1, padding=(0, 1, 1))
        self.bn2 = FrozenBatchNorm3d(planes)
        self.conv3 = conv_T(planes, planes, stride=1, padding=(1, 0, 0))
        self.bn3 = FrozenBatchNorm3d(planes)
        if expand:
            self.conv4 = Conv3d(planes, planes * 4, kernel_size=1)
            self.bn4 = FrozenBatchNorm3d(planes * 4)
            self.downsample = nn.Sequential(Conv3d(inplanes, planes * 4, kernel_size=1, stride=2),
                                            FrozenBatchNorm3d(planes * 4))
        else:
            self.conv4 = Conv3d(planes, inplanes, kernel_size=1)
            self.bn4 = FrozenBatchNorm3d(inplanes)
        self.relu = Slot("ReLU (fused into the BN pass)")

    def forward(self, x):
        out = self.bn1(self.conv1(x), relu=True)
        if self.ST == 'A':          # S then T (backbone.py:58)
            out = self.bn2(self.conv2(out), relu=True)
            out = self.bn3(self.conv3(out), relu=True)
        elif self.ST == 'B':        # S and T in parallel, summed (backbone.py:69)
            s = self.bn2(self.conv2(out), relu=True)
            t = self.bn3(self.conv3(out), relu=True)
            out = t + s
        else:                       # S, then S + T(S) (backbone.py:80)
            s = self.bn2(self.conv2(out), relu=True)
            t = self.bn3(self.conv3(s), relu=True)
            out = s + t
        residual = x
        if self.expand:
            residual = self.downsample[1](self.downsample[0](x))
        # relu(bn4(conv4(out)) + residual) in one pass
        return self.bn4(self.conv4(out), relu=True, residual=residual)


class Stem(nn.Sequential):
    """C1 = Conv3d(3x7x7, s2) -> BN -> ReLU -> MaxPool3d(2) with the reference's child indices (backbone.py:123-128)."""

    def forward(self, x):
        return ops.maxpool2(self[1](self[0](x), relu=True))


class P3D(nn.Module):
    def __init__(self, block, layers, input_channel=1, config=None):
        super().__init__()
        self.inplanes = config.BACKBONE_CHANNELS[0]
        stem_k = getattr(config, "BACKBONE_STEM_KERNEL", (3, 7, 7))
        self.C1 = Stem(
            Conv3d(input_channel, config.BACKBONE_CHANNELS[0], kernel_size=stem_k, stride=2,
                   padding=tuple(k // 2 for k in stem_k)),
            FrozenBatchNorm3d(config.BACKBONE_CHANNELS[0]),
            Slot("ReLU"),
            Slot("MaxPool3d(2, 2)"))
        self.C2 = self._make_layer(block, config.BACKBONE_CHANNELS[0], layers[0], stride=2)
        self.C3 = self._make_layer(block, config.BACKBONE_CHANNELS[1], layers[1], stride=2)
        for m in self.modules():   # backbone.py:133-139 (overwritten later by MaskRCNN.initialize_weights)
            if isinstance(m, nn.Conv3d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        layers = [block(self.inplanes, planes, 1, True, stride)]
        self.inplanes = planes * block.expansion
        for i in range(2, blocks + 1):
            layers.append(block(self.inplanes, planes, i, False))
        return nn.Sequential(*layers)

    def forward(self, x):
        return self.C3(self.C2(self.C1(x)))

    def stages(self):
        return [self.C1, self.C2, self.C3]


def P3D19(**kwargs):
    return P3D(Bottleneck, [2, 3], **kwargs)


def P3D35(**kwargs):
    """LiTS variant (reference LiTS_2017/backbone.py:172)."""
    return P3D(Bottleneck, [4, 5], **kwargs)
