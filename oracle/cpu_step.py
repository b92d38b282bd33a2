"""The reference's full SSL pre-training step as a stock-PyTorch module graph (TEST / BASELINE INFRASTRUCTURE ONLY).

What bench.py times as `cpu_baseline` (kind "port"), as the `--impl reference` arm (on the host cores) and as
`gpu_eager_baseline` (the same graph on the GPU through cuBLAS / cuDNN / ATen -- SURVEY 8d "the performance bar"):
`oracle.torch_ref.RefMSFWSI` (two plain ResNet-18 encoders + the heads of src/models/backbone.py:12-31, 129-222 built from
nn.Linear / nn.BatchNorm1d), the loss block of tools/ssl_train.py:448-466, backward and torch.optim.Adam over the three
learning-rate groups (ssl_train.py:281-309).  No product code (msfwsi_b200) runs here.  The unmodified reference cannot
travel to the GPU box (/root/reference does not exist there), hence a port; oracle/torch_ref.py is pinned to the
reference through tests/golden (tests/test_oracle_golden.py::test_torch_ref_*)."""
from __future__ import annotations

import os
import time

import torch

from . import torch_ref as R


class ReferenceStep:
    """device="cpu": fp32 on all host threads.  device="cuda:N": what tools/ssl_train.py runs -- autocast (bf16),
    channels_last is NOT applied (the reference does not), SyncBatchNorm + DDP when torch.distributed is initialised."""

    def __init__(self, batch: int, img: int = 224, loss: str = "cosine", tau: float = 0.07, seed: int = 3407, threads: int | None = None,
                 device: str = "cpu", autocast_dtype: torch.dtype | None = None, lr: float = 1e-3):
        import torch.distributed as dist
        self.device = torch.device(device)
        self.threads = threads or os.cpu_count() or 1
        if self.device.type == "cpu":
            torch.set_num_threads(self.threads)
        g = torch.Generator().manual_seed(seed)
        self.B, self.K, self.loss_mode, self.tau, self.autocast_dtype = batch, 16, loss, tau, autocast_dtype
        model = R.RefMSFWSI(R.plain_resnet18, 4, 2048, 512, 0.5, False)
        self.ddp = self.device.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.ddp:
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)  # ssl_train.py:160
        model = model.to(self.device).train()
        self.model = model
        self.step_mod = _LossStep(model, loss, tau)
        if self.ddp:
            self.step_mod = torch.nn.parallel.DistributedDataParallel(self.step_mod, device_ids=[self.device.index])  # :170
        groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
        self.opt = torch.optim.Adam(groups, lr=lr)  # :281-309
        self.ctx = [torch.randn(batch, 3, img, img, generator=g).to(self.device) for _ in range(2)]
        self.tgt = [torch.randn(batch * self.K, 3, img, img, generator=g).to(self.device) for _ in range(2)]
        self.rev = [torch.stack([torch.randperm(self.K, generator=g).argsort() for _ in range(batch)]) for _ in range(2)]  # stay on the CPU, like the reference's

    def step(self) -> float:
        """One step (ssl_train.py:441-474).  Returns host seconds for CPU; on CUDA the caller times with events."""
        t0 = time.perf_counter()
        if self.autocast_dtype is not None:
            with torch.autocast(self.device.type, dtype=self.autocast_dtype):
                loss = self.step_mod((self.ctx[0], self.tgt[0]), (self.ctx[1], self.tgt[1]), self.rev)
        else:
            loss = self.step_mod((self.ctx[0], self.tgt[0]), (self.ctx[1], self.tgt[1]), self.rev)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.last = loss
        if self.device.type == "cpu":
            self.last_loss = float(loss.item())
        return time.perf_counter() - t0


class _LossStep(torch.nn.Module):
    def __init__(self, model, loss, tau):
        super().__init__()
        self.model, self.loss, self.tau = model, loss, tau

    def forward(self, x1, x2, rev):
        return R.ref_ssl_loss(self.model(x1, x2, rev), (0.1, 0.4, 0.7, 1.0), self.loss, self.tau)


CpuReferenceStep = ReferenceStep  # round-1 name
