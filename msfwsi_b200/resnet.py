"""ResNet-18/34 encoder with pooled pyramid features -- the producer of the hot path's inputs.

Out of scope for the CUDA work (BASELINE.json:north_star: "the backbone convolutions stay on
PyTorch/cuDNN"); written here only so the drop-in module is self-contained.  It keeps the
reference's state-dict key contract (`conv1, bn1, layer{1..4}.{i}.{conv1,bn1,conv2,bn2,downsample.{0,1}}, fc`;
src/models/resnet.py:145-205, consumed by tools/ssl_finetune.py:153-170) and its
`return_features=True` behaviour: global-average-pooled layer1..4 outputs of widths
64/128/256/512 (src/models/resnet.py:244-254).
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, layers, num_classes: int = 1000, zero_init_residual: bool = False, return_features: bool = False):
        super().__init__()
        self.return_features = return_features
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def _make_layer(self, planes: int, blocks: int, stride: int) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        seq = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        seq += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = self.layer4(x3)
        out = self.fc(torch.flatten(self.avgpool(x4), 1))
        if not self.return_features:
            return out
        pooled = tuple(torch.flatten(self.avgpool(t), 1) for t in (x1, x2, x3))
        return (*pooled, out)


def _build(layers, pretrained, **kw):
    model = ResNet(layers, **kw)
    if pretrained:
        path = os.environ.get("MSFWSI_PRETRAINED_RESNET")
        if path and os.path.exists(path):
            model.load_state_dict(torch.load(path, map_location="cpu"))
        else:
            # the reference downloads ImageNet weights here (src/models/resnet.py:271-274); no network in this
            # environment, so point MSFWSI_PRETRAINED_RESNET at a local torchvision state_dict to get them
            warnings.warn("pretrained=True requested but MSFWSI_PRETRAINED_RESNET is not set: random init", stacklevel=2)
    return model


def resnet18(pretrained: bool = False, **kw) -> ResNet:
    return _build([2, 2, 2, 2], pretrained, **kw)


def resnet34(pretrained: bool = False, **kw) -> ResNet:
    return _build([3, 4, 6, 3], pretrained, **kw)
