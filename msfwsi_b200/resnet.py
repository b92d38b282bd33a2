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


_INT32_MAX = 2 ** 31 - 1


def _batch_chunks(x: torch.Tensor):
    """Split along the batch so that every piece has fewer than 2^31 elements (ATen's batch-norm / pooling kernels drop
    to a ~20x slower 64-bit-index path above that; at 4096 x 64 x 112 x 112 the stem crosses it)."""
    n = x.shape[0]
    per = max(1, _INT32_MAX // max(1, x[0].numel()))
    return [x] if n <= per else list(x.split(per, dim=0))


class _StemBNFn(torch.autograd.Function):
    """Train-mode batch norm with exact full-batch (and cross-rank) statistics, evaluated chunk by chunk with the same
    ATen primitives torch.nn.SyncBatchNorm uses (batch_norm_stats / gather_stats_with_counts / elemt and the two
    backward halves)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group):
        chunks = _batch_chunks(x)
        means, invstds, counts = [], [], []
        for c in chunks:
            m, iv = torch.batch_norm_stats(c, eps)
            means.append(m)
            invstds.append(iv)
            counts.append(c.numel() // c.shape[1])
        mean_all, invstd_all = torch.stack(means), torch.stack(invstds)
        count_all = torch.tensor(counts, dtype=mean_all.dtype, device=x.device)
        world = torch.distributed.get_world_size(group) if group is not None else 1
        if world > 1:
            packed = torch.cat([mean_all, invstd_all, count_all[:, None]], dim=1)
            gathered = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=x.device)
            torch.distributed.all_gather_into_tensor(gathered, packed, group=group)
            gathered = gathered.flatten(0, 1)
            ch = x.shape[1]
            mean_all, invstd_all, count_all = gathered[:, :ch].contiguous(), gathered[:, ch:2 * ch].contiguous(), gathered[:, 2 * ch].contiguous()
        mean, invstd = torch.batch_norm_gather_stats_with_counts(chunks[0], mean_all, invstd_all, running_mean, running_var,
                                                                 momentum, eps, count_all)
        out = torch.empty_like(x)
        for c, o in zip(chunks, _batch_chunks(out)):
            o.copy_(torch.batch_norm_elemt(c, weight, bias, mean, invstd, eps))
        ctx.save_for_backward(x, weight, mean, invstd, count_all.to(torch.int32))
        ctx.group, ctx.world = group, world
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, mean, invstd, count = ctx.saved_tensors
        grad_out = grad_out.contiguous(memory_format=torch.channels_last) if x.is_contiguous(memory_format=torch.channels_last) else grad_out.contiguous()
        xs, gs = _batch_chunks(x), _batch_chunks(grad_out)
        sum_dy = sum_dy_xmu = gw = gb = None
        for c, g in zip(xs, gs):
            a, b2, w_, b_ = torch.batch_norm_backward_reduce(g, c, mean, invstd, weight, True, True, True)
            sum_dy = a if sum_dy is None else sum_dy + a
            sum_dy_xmu = b2 if sum_dy_xmu is None else sum_dy_xmu + b2
            gw = w_ if gw is None else gw + w_
            gb = b_ if gb is None else gb + b_
        if ctx.world > 1:
            packed = torch.cat([sum_dy, sum_dy_xmu])
            torch.distributed.all_reduce(packed, group=ctx.group)
            sum_dy, sum_dy_xmu = packed[:sum_dy.numel()], packed[sum_dy.numel():]
        grad_in = torch.empty_like(x)
        for c, g, o in zip(xs, gs, _batch_chunks(grad_in)):
            o.copy_(torch.batch_norm_backward_elemt(g, c, mean, invstd, weight, sum_dy, sum_dy_xmu, count))
        return grad_in, gw.to(weight.dtype), gb.to(weight.dtype), None, None, None, None, None


class StemBatchNorm2d(nn.Module):
    """BatchNorm2d of the stem (same parameters / buffers / state-dict keys as nn.BatchNorm2d).  Deliberately NOT a
    ``_BatchNorm`` subclass: ``convert_sync_batchnorm`` leaves it alone and it synchronises its statistics across the
    default process group itself (same math as SyncBatchNorm), while keeping every ATen call below 2^31 elements."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def forward(self, x):
        import torch.distributed as dist
        sync = self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if not self.training or not x.is_cuda or (x.numel() <= _INT32_MAX and not sync):
            if self.training:
                self.num_batches_tracked.add_(1)
            return nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training,
                                            self.momentum, self.eps)
        self.num_batches_tracked.add_(1)
        return _StemBNFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                               dist.group.WORLD if sync else None)


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
        self.bn1 = StemBatchNorm2d(64)
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
            elif isinstance(m, (nn.BatchNorm2d, StemBatchNorm2d)):
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
        x = self.relu(self.bn1(self.conv1(x)))
        x = torch.cat([self.maxpool(c) for c in _batch_chunks(x)], dim=0) if x.numel() > _INT32_MAX else self.maxpool(x)
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
