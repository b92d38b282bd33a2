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


class FusedBatchNorm2d(nn.Module):
    """BatchNorm2d with the element-wise work around it fused in (same parameters / buffers / state-dict keys as
    nn.BatchNorm2d): ``act`` = "none" | "relu" | "relu_pool" (the stem's bn -> relu -> 3x3/2 max-pool), and an optional
    residual added before the ReLU (BasicBlock's ``relu(bn2(.) + identity)``).

    On CUDA in train mode it runs on this repo's channels-last kernels (``ops.bn_act2d``: 3 streaming passes forward,
    2 backward, 64-bit indexing -- ATen's kernels fall onto a ~20x slower path above 2^31 elements, which the stem of a
    4096-image batch exceeds).  Deliberately NOT a ``_BatchNorm`` subclass: ``convert_sync_batchnorm`` leaves it alone
    and it reduces its statistics over the default process group itself (SyncBatchNorm semantics, tools/ssl_train.py:160)
    whenever torch.distributed is initialised with more than one rank.  Eval mode and CPU tensors use the plain ATen
    sequence (not a hot path)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, act: str = "none"):
        super().__init__()
        if act not in ("none", "relu", "relu_pool"):
            raise ValueError(f"act must be none | relu | relu_pool, got {act!r}")
        self.num_features, self.eps, self.momentum, self.act = num_features, eps, momentum, act
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}, act={self.act}"

    def forward(self, x, residual=None, want_mean=False):
        """``want_mean=True`` returns ``(out, out.mean((2, 3)))``: the spatial mean rides along so that its gradient is
        folded into the fused backward (only for the relu(bn(x) + residual) form)."""
        if self.training and x.is_cuda:
            import torch.distributed as dist
            from . import ops
            sync = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            self.num_batches_tracked.add_(1)
            fused_mean = want_mean and residual is not None and self.act == "relu"
            out = ops.bn_act2d(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                               relu=self.act != "none", residual=residual, pool=self.act == "relu_pool",
                               sync_group=dist.group.WORLD if sync else None, want_mean=fused_mean)
            if want_mean and not fused_mean:
                return out, out.mean(dim=(2, 3))
            return out
        if self.training:
            self.num_batches_tracked.add_(1)
        out = nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training, self.momentum, self.eps)
        if residual is not None:
            out = out + residual
        if self.act != "none":
            out = nn.functional.relu(out)
        if self.act == "relu_pool":
            out = nn.functional.max_pool2d(out, 3, 2, 1)
        return (out, out.mean(dim=(2, 3))) if want_mean else out


def ops_stem_weight(w):
    from . import ops
    return ops.stem_s2d_weight(w)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = FusedBatchNorm2d(planes, act="relu")
        self.relu = nn.ReLU(inplace=True)  # kept for module-tree compatibility; the ReLUs run inside bn1 / bn2
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = FusedBatchNorm2d(planes, act="relu")
        self.downsample = downsample

    def forward(self, x, want_mean=False):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.bn1(self.conv1(x))                       # relu(bn1(conv1 x))
        return self.bn2(self.conv2(out), idt, want_mean)    # relu(bn2(conv2 out) + identity) [, its spatial mean]


class ResNet(nn.Module):
    def __init__(self, layers, num_classes: int = 1000, zero_init_residual: bool = False, return_features: bool = False):
        super().__init__()
        self.return_features = return_features
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = FusedBatchNorm2d(64, act="relu_pool")  # bn1 -> relu -> maxpool in one pass (resnet.py:244-247 of the reference)
        self.relu = nn.ReLU(inplace=True)     # kept for module-tree compatibility (no parameters)
        self.maxpool = nn.MaxPool2d(3, 2, 1)  # idem
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, FusedBatchNorm2d)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def _make_layer(self, planes: int, blocks: int, stride: int) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), FusedBatchNorm2d(planes))
        seq = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        seq += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    def _stem_conv(self, x):
        """conv1 (7x7, stride 2, padding 3 over 3 channels).  On CUDA it is handed to cuDNN in its equivalent
        space-to-depth form -- a 4x4 stride-1 convolution over the 16-channel pixel-unshuffled input (``ops.stem_s2d``)
        with the same weights rearranged (``ops.stem_s2d_weight``) -- which cuDNN runs ~3x faster forward and backward
        than the 3-channel form; the parameter, its gradient and the state dict keep the (64, 3, 7, 7) shape."""
        c = self.conv1
        w = c.weight
        if x.shape[1] == 16 and tuple(w.shape[1:]) == (3, 7, 7):
            # already in the space-to-depth layout (ops.view_crops_s2d writes the views like this straight from the uint8 tiles)
            return nn.functional.conv2d(x, ops_stem_weight(w))
        if (x.is_cuda and not x.requires_grad and tuple(w.shape[1:]) == (3, 7, 7) and tuple(c.stride) == (2, 2) and tuple(c.padding) == (3, 3)
                and tuple(c.dilation) == (1, 1) and c.groups == 1 and c.bias is None and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0):
            from . import ops
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else w.dtype
            return nn.functional.conv2d(ops.stem_s2d(x, dt), ops.stem_s2d_weight(w))
        return c(x)

    def forward(self, x):
        x = self.bn1(self._stem_conv(x))
        # the global average pool of each layer's output (src/models/resnet.py:244-254) is produced by the layer's last
        # block together with the feature map, so the fused backward sees both gradients at once
        pooled = []
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer[:-1]:
                x = blk(x)
            x, m = layer[-1](x, want_mean=True)
            pooled.append(m)
        out = self.fc(pooled[3])
        if not self.return_features:
            return out
        return (*pooled[:3], out)


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
