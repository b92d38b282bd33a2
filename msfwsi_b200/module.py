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
    """nn.Linear on this repo's GEMM kernels (ops.linear_tc): tcgen05 under bf16 / fp16 autocast, the exact-fp32 SIMT
    kernel otherwise.  Same parameters and state-dict keys as nn.Linear; no cuBLAS, no CPU path.

    The GEMM reads a 16-bit copy of the fp32 master weight (what autocast's cast cache holds for F.linear).  The copy
    is refreshed whenever the parameter changed: ATen in-place writes move ``weight._version``; raw-pointer writers
    (msf_adam_multi, msf_ema_multi) bump ``_lib.param_epoch`` instead -- unless the copy is *maintained*, i.e. registered
    with ``FusedAdam.attach_shadows`` (see :func:`bind_optimizer`), in which case the optimizer kernel rewrites it in the
    same pass as the parameter and no cast runs per step."""
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
        """Stand-alone call (the training step goes through ``MSFWSI.head_stage``, which runs all heads at once)."""
        if not x.is_cuda:
            raise RuntimeError("msfwsi_b200.TCLinear runs on CUDA tensors only (there is no CPU fallback)")
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else self.weight.dtype
        x2 = x.reshape(-1, x.shape[-1])
        w_op = self.lowp_weight(dt) if dt in (torch.bfloat16, torch.float16) else None
        y = ops.linear_tc(x2 if w_op is not None or x2.dtype == self.weight.dtype else x2.to(self.weight.dtype), self.weight, self.bias, w_op)
        return y.reshape(*x.shape[:-1], self.out_features)


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
    Eval mode normalises with the running statistics on the same kernels; CPU tensors raise."""

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
        """Stand-alone call (the training step goes through ``MSFWSI.head_stage``)."""
        vec = 8 if x.dtype in (torch.bfloat16, torch.float16) else 4
        if not x.is_cuda:
            raise RuntimeError("msfwsi_b200.FusedBatchNorm1d runs on CUDA tensors only (there is no CPU fallback)")
        if x.dim() != 2 or self.num_features % vec:
            raise ValueError(f"FusedBatchNorm1d: expected (rows, {self.num_features}) with the width a multiple of {vec}, got {tuple(x.shape)}")
        if self.training:
            import torch.distributed as dist
            sync = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if x.shape[0] * (dist.get_world_size() if sync else 1) <= 1:  # what nn.BatchNorm1d / SyncBatchNorm raise
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
            self.num_batches_tracked.add_(1)
            return ops.bn_act2d(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum,
                                relu=self.act == "relu", sync_group=dist.group.WORLD if sync else None)
        if torch.is_grad_enabled() and (x.requires_grad or (self.weight is not None and self.weight.requires_grad)):
            raise RuntimeError("FusedBatchNorm1d: eval-mode backward is only available through MSFWSI.head_stage")
        return ops.bn_eval_apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, relu=self.act == "relu")


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
    def _head_refs(self):
        """The 12 heads in stage order: context levels 0-3, target levels 0-3, inter (fuser) levels 0-3."""
        from .heads import HeadRefs
        refs = self.__dict__.get("_refs")
        if refs is None:
            refs = [HeadRefs(pj, pd) for pjs, pds in ((self.context_projector, self.context_predictor), (self.target_projector, self.target_predictor),
                                                      (self.inter_projector, self.inter_predictor)) for pj, pd in zip(pjs, pds)]
            self.__dict__["_refs"] = refs
        return refs

    def _apply(self, fn, *args, **kwargs):  # .to() / .cuda() / .double() replace parameter tensors: rebuild the references lazily
        self.__dict__.pop("_refs", None)
        return super()._apply(fn, *args, **kwargs)

    def head_stage(self, context_f1, context_f2, target_f1, target_f2, jigsaw_idx, want_keys: bool = False, want_rowsq: bool = False):
        """Everything after the encoder calls (backbone.py:147-222) on two-view stacks.  Returns (p, z, extras): lists over the
        12 heads (context levels, target levels, inter levels) of (2, rows, dim) tensors; z is detached."""
        from . import heads as H
        B, nl = context_f1[0].shape[0], len(context_f1)
        dev = context_f1[0].device
        if jigsaw_idx is None or len(jigsaw_idx) != 2:
            raise AssertionError("jigsaw_idx must be [rev_view1, rev_view2]")
        rev = [torch.as_tensor(r).to(device=dev, non_blocking=True) for r in jigsaw_idx]
        for r in rev:
            assert tuple(r.shape) == (B, self.K), "batch_idx.shape == jigsaw_idx shape"  # backbone.py:152
        # one launch: un-shuffle 16 target vectors per sample + build the fuser inputs, 4 levels x 2 views, written as
        # two-view stacks (the layout the head stage consumes)
        ctx_all = list(context_f1) + list(context_f2)
        tgt_all = list(target_f1) + list(target_f2)
        dt = tgt_all[0].dtype
        ctx_all = [t if t.dtype == dt else t.to(dt) for t in ctx_all]
        rev_all = [rev[0]] * nl + [rev[1]] * nl
        sorted_st, ms_st = ops.gather_concat(ctx_all, tgt_all, rev_all, self.K, self.n_keep, self.validate_indices, stacked=True)
        ctx_st = [torch.stack((a, b)) for a, b in zip(ctx_all[:nl], ctx_all[nl:])]
        refs = self._head_refs()
        import torch.distributed as dist
        sync = self.training and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.training:
            # every head BatchNorm1d is applied to view 1 and to view 2 (nn.BatchNorm1d counts each call)
            torch._foreach_add_([bn.num_batches_tracked for r in refs for bn in r.bn], 2)
        return H.head_stage(ctx_st + sorted_st + ms_st, refs, self.training, dist.group.WORLD if sync else None,
                            want_keys=want_keys, want_rowsq=want_rowsq)

    def heads(self, context_f1, context_f2, target_f1, target_f2, jigsaw_idx):
        """The reference's nested output ((ctx p1,p2,z1,z2), (tgt ...), (inter ...)), each a tuple over the pyramid levels."""
        p, z, _ = self.head_stage(context_f1, context_f2, target_f1, target_f2, jigsaw_idx)
        nl = len(context_f1)
        out = []
        for b in range(3):
            hs = range(b * nl, (b + 1) * nl)
            out.append((tuple(p[h][0] for h in hs), tuple(p[h][1] for h in hs), tuple(z[h][0] for h in hs), tuple(z[h][1] for h in hs)))
        return tuple(out)

    def forward(self, x1, x2, jigsaw_idx=None):
        context_f1, context_f2 = self.context_encoder(x1[0]), self.context_encoder(x2[0])
        target_f1, target_f2 = self.target_encoder(x1[1]), self.target_encoder(x2[1])
        return self.heads(context_f1, context_f2, target_f1, target_f2, jigsaw_idx)

    def heads_loss(self, context_f1, context_f2, target_f1, target_f2, jigsaw_idx, fuser_weights: Sequence[float] = DEFAULT_FUSER_WEIGHTS,
                   mode: str = "cosine", tau: float = 0.07, group=None) -> torch.Tensor:
        """Head stage + the loss block of ssl_train.py:448-466 on the two-view stacks (no per-view slicing in between)."""
        nl = len(context_f1)
        if mode == "cosine":
            p, z, _ = self.head_stage(context_f1, context_f2, target_f1, target_f2, jigsaw_idx)
            return ops.cosine_loss_stacked(p, z, [-0.5 * fuser_weights[h % nl] for h in range(len(p))])
        if mode == "infonce":
            if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16 and context_f1[0].is_cuda:
                # fused front end: the head stage leaves the normalised keys (one buffer, one all-gather) and the row norms
                # of p; ONE grouped call computes all 24 terms
                p, z, ex = self.head_stage(context_f1, context_f2, target_f1, target_f2, jigsaw_idx, want_keys=True, want_rowsq=True)
                return ops.infonce_grouped(p, ex, [0.5 * fuser_weights[h % nl] for h in range(len(p))], tau)
            p, z, _ = self.head_stage(context_f1, context_f2, target_f1, target_f2, jigsaw_idx)
            total = None
            for h in range(len(p)):
                for v in range(2):
                    term = ops.infonce_loss(p[h][v], z[h][1 - v], tau=tau, group=group) * (0.5 * fuser_weights[h % nl])
                    total = term if total is None else total + term
            return total
        raise ValueError(f"unknown loss mode {mode!r} (expected 'cosine' or 'infonce')")

    def forward_loss(self, x1, x2, jigsaw_idx, fuser_weights: Sequence[float] = DEFAULT_FUSER_WEIGHTS, mode: str = "cosine",
                     tau: float = 0.07, group=None) -> torch.Tensor:
        """``forward`` + the loss block of ssl_train.py:448-466 fused on the device (no .item() sync)."""
        context_f1, context_f2 = self.context_encoder(x1[0]), self.context_encoder(x2[0])
        target_f1, target_f2 = self.target_encoder(x1[1]), self.target_encoder(x2[1])
        return self.heads_loss(context_f1, context_f2, target_f1, target_f2, jigsaw_idx, fuser_weights, mode, tau, group)
