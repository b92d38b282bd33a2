"""CPU port of one full SSL pre-training step (TEST / BASELINE INFRASTRUCTURE ONLY).

What bench.py times as `cpu_baseline` (kind "port") and as the `--impl reference` arm: the reference's
step restated on the host cores -- two ResNet-18 encoders (plain PyTorch modules with the reference's
layout, src/models/resnet.py), the head path of src/models/backbone.py:147-222 through
`oracle.msf_oracle.heads_forward`, the loss block of tools/ssl_train.py:448-466 (or the InfoNCE
extension), backward and an Adam step, fp32, all host threads.  The unmodified reference cannot travel to
the GPU box (/root/reference does not exist there), hence a port; it was checked against the reference
itself through tests/golden (tests/test_oracle_golden.py)."""
from __future__ import annotations

import os
import time

import torch

from . import msf_oracle as O


class CpuReferenceStep:
    def __init__(self, batch: int, img: int = 224, loss: str = "cosine", tau: float = 0.07, seed: int = 3407, threads: int | None = None):
        from msfwsi_b200.resnet import resnet18  # plain torch module (runs on CPU); the backbone is not a CUDA target
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        g = torch.Generator().manual_seed(seed)
        self.B, self.K, self.loss_mode, self.tau = batch, 16, loss, tau
        self.enc = [resnet18(zero_init_residual=True, return_features=True) for _ in range(2)]
        for e in self.enc:
            e.fc = torch.nn.Identity()
            e.train()
        self.heads = {k: v.clone().requires_grad_(True) for k, v in O.closed_form_head_params().items()}
        params = [p for e in self.enc for p in e.parameters()] + list(self.heads.values())
        self.opt = torch.optim.Adam(params, lr=1e-3)
        self.ctx = [torch.randn(batch, 3, img, img, generator=g) for _ in range(2)]
        self.tgt = [torch.randn(batch * self.K, 3, img, img, generator=g) for _ in range(2)]
        self.rev = [torch.stack([O.jigsaw_indices(g, self.K)[1] for _ in range(batch)]) for _ in range(2)]

    def step(self) -> float:
        t0 = time.perf_counter()
        cf = [self.enc[0](x) for x in self.ctx]
        tf = [self.enc[1](x) for x in self.tgt]
        out = O.heads_forward(cf[0], cf[1], tf[0], tf[1], self.rev[0], self.rev[1], self.heads, self.K, 8)
        if self.loss_mode == "cosine":
            loss = O.ssl_loss_block(out)
        else:
            w = (0.1, 0.4, 0.7, 1.0)
            loss = 0.0
            for br in out:
                for l, (p1, p2, z1, z2) in enumerate(zip(*br)):
                    loss = loss + 0.5 * w[l] * (O.infonce_loss(p1, z2, self.tau)[0] + O.infonce_loss(p2, z1, self.tau)[0])
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.last_loss = float(loss.item())
        return time.perf_counter() - t0
