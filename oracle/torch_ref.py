"""Plain-PyTorch MODULE restatement of the reference step (TEST / BASELINE INFRASTRUCTURE ONLY).

Where ``msf_oracle.py`` restates the arithmetic functionally in fp64, this file restates the reference as the
``nn.Module`` graph it is -- ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` / ``nn.BatchNorm1d`` /
``nn.CosineSimilarity`` / ``torch.optim.Adam`` -- so that it can be run under CUDA autocast exactly as
``tools/ssl_train.py`` runs the reference.  It is

  * the **GPU-eager baseline** of bench.py (SURVEY 8d "the performance bar": the same step through cuBLAS / cuDNN / ATen
    on the same B200 in the same job), and the encoder of the CPU arm (so no product code runs in a baseline leg);
  * the **training-parity oracle** of tests/test_training_parity_gpu.py (multi-step bf16 trajectories).

Nothing here is imported by the product package.  Structure and state-dict keys follow the reference:
  ResNet (torchvision layout, pooled pyramid features)        src/models/resnet.py:145-256
  make_projector / make_predictor                             src/models/backbone.py:12-31
  RefMSFWSI.__init__ / forward                                src/models/backbone.py:39-104, 129-222
  ref_ssl_loss                                                tools/ssl_train.py:422, 448-466
Checked against the unmodified reference through tests/golden (tests/test_oracle_golden.py::test_torch_ref_*).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

WIDTHS = (64, 128, 256, 512)  # backbone.py:67


# ---------------------------------------------------------------------------------------------------------------
# encoder: ResNet-18/34 with BasicBlocks, torchvision key layout, return_features (resnet.py:232-256)
# ---------------------------------------------------------------------------------------------------------------
class Block(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        skip = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(y)) + skip)


class PlainResNet(nn.Module):
    def __init__(self, depths=(2, 2, 2, 2), zero_init_residual: bool = False, return_features: bool = False, **_):
        super().__init__()
        self.return_features = return_features
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (w, n) in enumerate(zip(WIDTHS, depths), start=1):
            blocks = [Block(cin, w, 1 if i == 1 else 2)] + [Block(w, w, 1) for _ in range(n - 1)]
            setattr(self, f"layer{i}", nn.Sequential(*blocks))
            cin = w
        self.fc = nn.Linear(512, 1000)
        for m in self.modules():  # resnet.py:190-205
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Block):
                    nn.init.zeros_(m.bn2.weight)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        feats = []
        for i in range(1, 5):
            x = getattr(self, f"layer{i}")(x)
            feats.append(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))
        out = self.fc(feats[3])
        return (feats[0], feats[1], feats[2], out) if self.return_features else out


def plain_resnet18(**kw) -> PlainResNet:
    kw.pop("pretrained", None)  # no network: random init (the reference downloads ImageNet weights, resnet.py:271-274)
    return PlainResNet((2, 2, 2, 2), **kw)


# ---------------------------------------------------------------------------------------------------------------
# heads (backbone.py:12-31)
# ---------------------------------------------------------------------------------------------------------------
def make_projector(din: int, dout: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(din, din, bias=False), nn.BatchNorm1d(din), nn.ReLU(inplace=True),
                         nn.Linear(din, din, bias=False), nn.BatchNorm1d(din), nn.ReLU(inplace=True),
                         nn.Linear(din, dout, bias=False), nn.BatchNorm1d(dout, affine=False))


def make_predictor(din: int, hidden: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(din, hidden, bias=False), nn.BatchNorm1d(hidden), nn.ReLU(inplace=True), nn.Linear(hidden, din))


class RefMSFWSI(nn.Module):
    """backbone.py:34-222 with stock torch modules; same attribute names and state-dict keys."""

    def __init__(self, base_encoder=plain_resnet18, scale: int = 4, dim: int = 2048, pred_dim: int = 512, mask_ratio: float = 0.5,
                 use_checkpoint: bool = False):
        super().__init__()
        self.K = int(scale ** 2)
        self.n_keep = int(self.K * (1 - mask_ratio))
        self.context_encoder = base_encoder(zero_init_residual=True, pretrained=True, return_features=True)
        self.target_encoder = base_encoder(zero_init_residual=True, pretrained=True, return_features=True)
        self.context_encoder.fc = nn.Identity()
        self.target_encoder.fc = nn.Identity()
        fused = [w * (self.n_keep + 1) for w in WIDTHS]
        self.context_projector = nn.ModuleList(make_projector(w, w) for w in WIDTHS)
        self.target_projector = nn.ModuleList(make_projector(w, w) for w in WIDTHS)
        self.inter_projector = nn.ModuleList(make_projector(w, w) for w in fused)
        self.context_predictor = nn.ModuleList(make_predictor(w, w // 4) for w in WIDTHS)
        self.target_predictor = nn.ModuleList(make_predictor(w, w // 4) for w in WIDTHS)
        self.inter_predictor = nn.ModuleList(make_predictor(w, w // 4) for w in fused)

    def heads(self, cf1, cf2, tf1, tf2, jigsaw_idx):
        B, K = cf1[0].shape[0], self.K
        dev = cf1[0].device
        split1 = [t.reshape(B, K, -1) for t in tf1]                        # backbone.py:147-148
        split2 = [t.reshape(B, K, -1) for t in tf2]
        bidx = torch.arange(B, device=dev)[:, None].expand(B, K)           # :151 (built on the device here)
        r1, r2 = (torch.as_tensor(r).to(dev) for r in jigsaw_idx)
        sort1 = [t[bidx, r1].flatten(0, 1) for t in split1]                # :153-158
        sort2 = [t[bidx, r2].flatten(0, 1) for t in split2]
        run = lambda mods, xs: [m(x) for m, x in zip(mods, xs)]
        cz1, cz2 = run(self.context_projector, cf1), run(self.context_projector, cf2)   # :161-172
        tz1, tz2 = run(self.target_projector, sort1), run(self.target_projector, sort2)
        cp1, cp2 = run(self.context_predictor, cz1), run(self.context_predictor, cz2)   # :175-186
        tp1, tp2 = run(self.target_predictor, tz1), run(self.target_predictor, tz2)
        ms1 = [torch.cat((c, s[:, :self.n_keep].flatten(1)), 1) for c, s in zip(cf1, split1)]  # :195-202 (shuffled order)
        ms2 = [torch.cat((c, s[:, :self.n_keep].flatten(1)), 1) for c, s in zip(cf2, split2)]
        mz1, mz2, mp1, mp2 = [], [], [], []
        for i in range(len(WIDTHS)):                                        # :205-212
            mz1.append(self.inter_projector[i](ms1[i]))
            mz2.append(self.inter_projector[i](ms2[i]))
            mp1.append(self.inter_predictor[i](mz1[i]))
            mp2.append(self.inter_predictor[i](mz2[i]))
        det = lambda ts: tuple(t.detach() for t in ts)                      # :188-191, 214-215
        return ((tuple(cp1), tuple(cp2), det(cz1), det(cz2)), (tuple(tp1), tuple(tp2), det(tz1), det(tz2)),
                (tuple(mp1), tuple(mp2), det(mz1), det(mz2)))

    def forward(self, x1, x2, jigsaw_idx=None):
        cf1, cf2 = self.context_encoder(x1[0]), self.context_encoder(x2[0])  # :140-145
        tf1, tf2 = self.target_encoder(x1[1]), self.target_encoder(x2[1])
        return self.heads(cf1, cf2, tf1, tf2, jigsaw_idx)


def ref_ssl_loss(outputs, fuser_weights: Sequence[float] = (0.1, 0.4, 0.7, 1.0), mode: str = "cosine", tau: float = 0.07):
    """tools/ssl_train.py:448-466 (mode="cosine"); mode="infonce" = the same pairs through F.cross_entropy over the
    normalised similarity matrix (the extension's torch expression, single process)."""
    cos = nn.CosineSimilarity(dim=1)
    loss = 0
    for branch in outputs:
        for i, (p1, p2, z1, z2) in enumerate(zip(*branch)):
            if mode == "cosine":
                loss = loss + (-(cos(p1, z2).mean() + cos(p2, z1).mean()) * 0.5) * fuser_weights[i]
            else:
                for p, z in ((p1, z2), (p2, z1)):
                    logits = F.normalize(p.float(), dim=1, eps=1e-8) @ F.normalize(z.float(), dim=1, eps=1e-8).t() / tau
                    loss = loss + 0.5 * fuser_weights[i] * F.cross_entropy(logits, torch.arange(p.shape[0], device=p.device))
    return loss
