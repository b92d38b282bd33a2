"""Drop-in for the reference's ``MSFWSI`` module (src/models/backbone.py:34-222) and the loss block of
``train`` (tools/ssl_train.py:448-466).

Same constructor, same ``forward(x1, x2, jigsaw_idx)`` signature and nested output, same
parameter / buffer names (so ``convert_sync_batchnorm``, DDP, the ``context_/target_/inter_`` prefix
filter of the optimizer at ssl_train.py:281-307 and checkpoints keep working).  What changes is the
hot path after the encoder calls: the inverse-jigsaw gather + fuser concat run as one CUDA launch
(``ops.gather_concat``) and the loss block as one fused launch (``ops.cosine_loss``) or the
flash-style InfoNCE kernel (``ops.infonce_loss``).  Under bf16 autocast the head Linears (forward, dX, dW) run
on this repo's persistent tcgen05 GEMM (``TCLinear`` -> ``ops.linear_tc``) and the heads' BatchNorm1d + ReLU on this
repo's batch-norm kernels (``FusedBatchNorm1d`` -> ``ops.bn_act2d``), which also carry the cross-rank statistics
(SyncBatchNorm semantics) in one fp64 all-reduce per layer and direction.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from . import ops

PYRAMID_WIDTHS = (64, 128, 256, 512)  # pooled layer1..4 widths of the ResNet-18/34 encoders
DEFAULT_FUSER_WEIGHTS = (0.1, 0.4, 0.7, 1.0)  # --fuser_weights default, tools/ssl_train.py:623-625


class TCLinear(nn.Linear):
    """nn.Linear whose 16-bit-autocast CUDA forward/backward run on this repo's tcgen05 GEMM (ops.linear_tc).
    Same parameters and state-dict keys as nn.Linear.

    The GEMM reads a 16-bit copy of the fp32 master weight (what autocast's cast cache holds for F.linear).  The copy
    is refreshed whenever the parameter changed: ATen in-place writes move ``weight._version``; raw-pointer writers
    (msf_adam_multi, msf_ema_multi) bump ``_lib.param_epoch`` instead -- unless the copy is *maintained*, i.e. registered
    with ``FusedAdam.attach_shadows`` (see :func:`bind_optimizer`), in which case the optimizer kernel rewrites it in the
    same pass as the parameter and no cast runs per step."""
    use_tc = True

    def lowp_weight(self, dtype: torch.dtype) -> torch.Tensor:
        from . import _lib
        w = self.weight
        st = self.__dict__.setdefault("_lowp", {}).get(dtype)
        key = (w._version, w.data_ptr(), w.device)
        # maintained = the optimizer bound to THIS storage rewrites the copy itself (a deepcopy / .to() of the module is not)
        if st is not None and st["key"] == key and (st["maintained_ptr"] == w.data_ptr() or st["epoch"] == _lib.param_epoch):
            return st["t"]
        with torch.no_grad():
            if st is not None and st["t"].device == w.device and st["t"].shape == w.shape:
                st["t"].copy_(w)  # in place: a maintained shadow keeps its address (the optimizer's table points at it)
            else:
                st = {"t": w.detach().to(dtype), "maintained_ptr": None}
                self._lowp[dtype] = st
        st["key"], st["epoch"] = key, _lib.param_epoch
        return st["t"]

    def forward(self, x):
        if TCLinear.use_tc and x.is_cuda and x.dim() == 2 and torch.is_autocast_enabled() and self.in_features % 8 == 0 and self.out_features % 8 == 0:
            dt = torch.get_autocast_dtype("cuda")
            if dt == torch.bfloat16:
                return ops.linear_tc(x, self.weight, self.bias, self.lowp_weight(dt))
        return super().forward(x)


def bind_optimizer(model: nn.Module, optimizer, dtype: torch.dtype = torch.bfloat16) -> int:
    """Register the 16-bit operand copies of every TCLinear weight in ``model`` with a ``FusedAdam`` so that
    msf_adam_multi rewrites them in the same pass as the fp32 masters (2 extra bytes per head parameter instead of a
    separate cast of the 99.8 M head parameters every step).  Returns the number of weights bound.  Optional: without it
    the copies are re-cast after every optimizer step (``_lib.param_epoch``)."""
    stepped = {id(p) for g in optimizer.param_groups for p in g["params"]}
    pairs = []
    for m in model.modules():
        if isinstance(m, TCLinear) and m.weight.is_cuda and id(m.weight) in stepped:
            sh = m.lowp_weight(dtype)
            m._lowp[dtype]["maintained_ptr"] = m.weight.data_ptr()
            pairs.append((m.weight, sh))
    optimizer.attach_shadows(pairs)
    return len(pairs)


class FusedBatchNorm1d(nn.Module):
    """BatchNorm1d (+ fused ReLU) of the heads on this repo's batch-norm kernels (``ops.bn_act2d`` over a [rows][C]
    matrix).  Same parameters / buffers / state-dict keys as nn.BatchNorm1d.  Like the encoders' ``FusedBatchNorm2d`` it
    is NOT a ``_BatchNorm`` subclass: ``convert_sync_batchnorm`` (tools/ssl_train.py:160) leaves it alone and it reduces
    its statistics over the default process group itself -- one fp64 all-reduce per direction instead of
    torch.nn.SyncBatchNorm's all_gather + Python-side recombination (1.5 ms of host time per call at these sizes).
    CPU tensors and eval mode take the plain ATen path."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, affine: bool = True, act: str = "none"):
        super().__init__()
        if act not in ("none", "relu"):
            raise ValueError(f"act must be none | relu, got {act!r}")
        self.num_features, self.eps, self.momentum, self.affine, self.act = num_features, eps, momentum, affine, act
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}, affine={self.affine}, act={self.act}"

    def forward(self, x):
        vec = 8 if x.dtype in (torch.bfloat16, torch.float16) else 4
        if self.training and x.is_cuda and x.dim() == 2 and self.num_features % vec == 0:
            import torch.distributed as dist
            sync = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if x.shape[0] * (dist.get_world_size() if sync else 1) <= 1:  # what nn.BatchNorm1d / SyncBatchNorm raise
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
            self.num_batches_tracked.add_(1)
            return ops.bn_act2d(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                                relu=self.act == "relu", sync_group=dist.group.WORLD if sync else None)
        if self.training:
            self.num_batches_tracked.add_(1)
        out = nn.functional.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, self.training, self.momentum, self.eps)
        return nn.functional.relu(out) if self.act == "relu" else out


class FusedAway(nn.Identity):
    """Placeholder that keeps the reference's nn.Sequential indices: the ReLU at this position runs inside the preceding
    FusedBatchNorm1d."""


def make_projector(in_dim: int, out_dim: int) -> nn.Sequential:
    """3-layer projector, indices 0..7 as in backbone.py:12-22 (last BN has no affine)."""
    layers = []
    for width_out, affine, act in ((in_dim, True, True), (in_dim, True, True), (out_dim, False, False)):
        layers.append(TCLinear(in_dim, width_out, bias=False))
        layers.append(FusedBatchNorm1d(width_out, affine=affine, act="relu" if act else "none"))
        if act:
            layers.append(FusedAway())
    return nn.Sequential(*layers)


def make_predictor(in_dim: int, hidden_dim: int) -> nn.Sequential:
    """2-layer bottleneck predictor, indices 0..3 as in backbone.py:25-31."""
    return nn.Sequential(TCLinear(in_dim, hidden_dim, bias=False), FusedBatchNorm1d(hidden_dim, act="relu"), FusedAway(),
                         TCLinear(hidden_dim, in_dim))


def ssl_loss(outputs, fuser_weights: Sequence[float] = DEFAULT_FUSER_WEIGHTS, mode: str = "cosine", tau: float = 0.07,
             group=None) -> torch.Tensor:
    """The loss of ``train`` over the module's nested output.

    mode="cosine"  -- reference-exact SimSiam negative cosine (ssl_train.py:448-466), ONE fused launch over
                      all branches x levels x directions.
    mode="infonce" -- extension: same (p, z) pairs, positives on the diagonal, global negatives (keys
                      all-gathered over ``group``), temperature ``tau``; weights/0.5 factors as in cosine mode."""
    ps, zs, coefs = [], [], []
    for branch in outputs:
        for lvl, (p1, p2, z1, z2) in enumerate(zip(*branch)):
            ps += [p1, p2]
            zs += [z2, z1]
            coefs += [fuser_weights[lvl]] * 2
    if mode == "cosine":
        return ops.cosine_loss(ps, zs, [-0.5 * c for c in coefs])
    if mode == "infonce":
        total = None
        for p, z, c in zip(ps, zs, coefs):
            term = ops.infonce_loss(p, z, tau=tau, group=group) * (0.5 * c)
            total = term if total is None else total + term
        return total
    raise ValueError(f"unknown loss mode {mode!r} (expected 'cosine' or 'infonce')")


class MSFWSI(nn.Module):
    """Multi-scale SSL model: two encoders + context / target / inter (fuser) heads."""

    def __init__(self, base_encoder, scale, dim=2048, pred_dim=512, mask_ratio=0.5, use_checkpoint=False):
        # ``dim`` / ``pred_dim`` are accepted and unused, exactly as in the reference (backbone.py:43-44)
        super().__init__()
        self.K = int(scale ** 2)
        self.n_keep = int(self.K * (1 - mask_ratio))
        self.context_encoder = base_encoder(zero_init_residual=True, pretrained=True, return_features=True)
        self.target_encoder = base_encoder(zero_init_residual=True, pretrained=True, return_features=True)
        self.context_encoder.fc = nn.Identity()
        self.target_encoder.fc = nn.Identity()

        self.inter_dim = torch.as_tensor(PYRAMID_WIDTHS)
        self.ms_inter_dim = self.inter_dim * (self.n_keep + 1)
        single = [int(d) for d in self.inter_dim]
        fused = [int(d) for d in self.ms_inter_dim]
        self.context_projector = nn.ModuleList(make_projector(d, d) for d in single)
        self.target_projector = nn.ModuleList(make_projector(d, d) for d in single)
        self.inter_projector = nn.ModuleList(make_projector(d, d) for d in fused)
        self.context_predictor = nn.ModuleList(make_predictor(d, d // 4) for d in single)
        self.target_predictor = nn.ModuleList(make_predictor(d, d // 4) for d in single)
        self.inter_predictor = nn.ModuleList(make_predictor(d, d // 4) for d in fused)
        self.validate_indices = False  # True: raise IndexError like the reference (costs a host sync)
        if use_checkpoint:
            self._apply_checkpoint()

    def _apply_checkpoint(self):
        from functools import partial

        from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import (CheckpointImpl, apply_activation_checkpointing,
                                                                                  checkpoint_wrapper)
        # backbone.py:106-119 (its `offload_to_cpu=False` keyword no longer exists in current torch and would be forwarded
        # to Conv2d.forward; omitted here).  The reference then replaces both stem convolutions by fresh, unwrapped
        # layers (backbone.py:121-127) -- here conv1 is simply left unwrapped, keeping its weights.
        wrap = partial(checkpoint_wrapper, checkpoint_impl=CheckpointImpl.NO_REENTRANT)
        stems = {id(self.context_encoder.conv1), id(self.target_encoder.conv1)} if hasattr(self.context_encoder, "conv1") else set()
        apply_activation_checkpointing(self, checkpoint_wrapper_fn=wrap,
                                       check_fn=lambda m: isinstance(m, (nn.Conv2d, nn.Linear)) and id(m) not in stems)

    # ---- hot path ----------------------------------------------------------------------------
    def heads(self, context_f1, context_f2, target_f1, target_f2, jigsaw_idx):
        """Everything after the encoder calls (backbone.py:147-222)."""
        B, nl = context_f1[0].shape[0], len(context_f1)
        dev = context_f1[0].device
        if jigsaw_idx is None or len(jigsaw_idx) != 2:
            raise AssertionError("jigsaw_idx must be [rev_view1, rev_view2]")
        rev = [torch.as_tensor(r).to(device=dev, non_blocking=True) for r in jigsaw_idx]
        for r in rev:
            assert tuple(r.shape) == (B, self.K), "batch_idx.shape == jigsaw_idx shape"  # backbone.py:152
        # one launch: un-shuffle 16 target vectors per sample + build the fuser inputs, 4 levels x 2 views
        ctx_all = list(context_f1) + list(context_f2)
        tgt_all = list(target_f1) + list(target_f2)
        rev_all = [rev[0]] * nl + [rev[1]] * nl
        sorted_all, ms_all = ops.gather_concat(ctx_all, tgt_all, rev_all, self.K, self.n_keep, self.validate_indices)
        target_f1_sort, target_f2_sort = sorted_all[:nl], sorted_all[nl:]
        ms_f1, ms_f2 = ms_all[:nl], ms_all[nl:]

        def run(heads_, feats):
            return tuple(h(f) for h, f in zip(heads_, feats))

        context_z1, context_z2 = run(self.context_projector, context_f1), run(self.context_projector, context_f2)
        target_z1, target_z2 = run(self.target_projector, target_f1_sort), run(self.target_projector, target_f2_sort)
        context_p1, context_p2 = run(self.context_predictor, context_z1), run(self.context_predictor, context_z2)
        target_p1, target_p2 = run(self.target_predictor, target_z1), run(self.target_predictor, target_z2)
        ms_z1, ms_z2, ms_p1, ms_p2 = [], [], [], []
        for i in range(nl):  # same call order as the reference so BN running stats see view 1 then view 2
            ms_z1.append(self.inter_projector[i](ms_f1[i]))
            ms_z2.append(self.inter_projector[i](ms_f2[i]))
            ms_p1.append(self.inter_predictor[i](ms_z1[i]))
            ms_p2.append(self.inter_predictor[i](ms_z2[i]))
        det = lambda ts: tuple(t.detach() for t in ts)  # keys never receive gradient
        return ((context_p1, context_p2, det(context_z1), det(context_z2)),
                (target_p1, target_p2, det(target_z1), det(target_z2)),
                (tuple(ms_p1), tuple(ms_p2), det(ms_z1), det(ms_z2)))

    def forward(self, x1, x2, jigsaw_idx=None):
        context_f1, context_f2 = self.context_encoder(x1[0]), self.context_encoder(x2[0])
        target_f1, target_f2 = self.target_encoder(x1[1]), self.target_encoder(x2[1])
        return self.heads(context_f1, context_f2, target_f1, target_f2, jigsaw_idx)

    def forward_loss(self, x1, x2, jigsaw_idx, fuser_weights: Sequence[float] = DEFAULT_FUSER_WEIGHTS, mode: str = "cosine",
                     tau: float = 0.07, group=None) -> torch.Tensor:
        """``forward`` + the loss block of ssl_train.py:448-466 fused on the device (no .item() sync)."""
        return ssl_loss(self.forward(x1, x2, jigsaw_idx), fuser_weights, mode, tau, group)
