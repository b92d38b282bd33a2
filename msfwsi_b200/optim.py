"""Multi-tensor Adam of the reference's training loop on one CUDA kernel.

`tools/ssl_train.py:303-309` builds ``torch.optim.Adam`` over three learning-rate groups (``context_`` / ``target_`` /
``inter_`` parameters) and `:472-474` drives it through ``GradScaler`` (unscale, non-finite check, skipped step).
``FusedAdam`` is a ``torch.optim.Adam`` subclass -- same constructor, ``param_groups`` and ``state_dict()`` layout
(``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so reference checkpoints load into it and vice versa -- whose
``step()`` is ONE launch of ``msf_adam_multi`` over a device-resident tensor table, with the gradient unscale, the
skip-on-overflow test and an optional EMA teacher update folded into the same pass.  No host synchronisation.
There is no CPU fallback: parameters must be CUDA fp32 tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L


def _dense(t: torch.Tensor) -> bool:
    """Element-wise kernels only need a dense layout shared by parameter, gradient and moments (channels-last conv
    weights are dense but not `is_contiguous()`)."""
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _like_layout(g: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    return g if g.stride() == p.stride() else torch.empty_like(p, dtype=g.dtype).copy_(g)


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._table_key = None
        self._keep = None
        self._ema: Dict[int, torch.Tensor] = {}
        self._ema_momentum = None
        self._shadow: Dict[int, torch.Tensor] = {}
        self._steps_verified = False  # all `step` counters equal (one shared device counter drives the bias corrections)
        # GradScaler.step() protocol of torch's fused optimizers: the scaler leaves the gradients scaled, sets
        # `self.grad_scale` / `self.found_inf` (device tensors) and this step unscales / skips on the device
        self._step_supports_amp_scaling = True

    # ---- optional EMA teacher, updated in the same pass as the parameters --------------------------------------
    def attach_ema(self, student: Sequence[torch.Tensor], teacher: Sequence[torch.Tensor], momentum: float) -> None:
        """teacher[i] <- momentum * teacher[i] + (1 - momentum) * student[i] right after student[i] is stepped."""
        student, teacher = list(student), list(teacher)
        if len(student) != len(teacher):
            raise ValueError("attach_ema: student / teacher lists differ in length")
        for s, t in zip(student, teacher):
            if s.shape != t.shape or t.dtype != torch.float32 or not t.is_cuda or t.stride() != s.stride():
                raise ValueError("attach_ema: teachers must be CUDA fp32 tensors with the students' shapes and strides")
            self._ema[id(s)] = t
        self._ema_momentum = float(momentum)
        self._table_key = None

    # ---- 16-bit operand copies of parameters (the head Linears' GEMM operands), rewritten in the same pass -----
    def attach_shadows(self, pairs) -> None:
        """pairs: iterable of (parameter, shadow) with shadow a dense CUDA bf16 / fp16 tensor of the parameter's shape
        and strides.  After every step shadow == parameter.to(shadow.dtype), written by msf_adam_multi itself, so a
        consumer never sees a copy that went stale behind the raw-pointer update (and no cast pass runs per step)."""
        for p, sh in pairs:
            if sh.shape != p.shape or sh.stride() != p.stride() or not sh.is_cuda or sh.dtype not in (torch.bfloat16, torch.float16):
                raise ValueError("attach_shadows: shadows must be CUDA bf16 / fp16 tensors with the parameter's shape and strides")
            self._shadow[id(p)] = sh
        self._table_key = None

    def load_state_dict(self, state_dict) -> None:
        """torch.optim.Adam.load_state_dict + invalidation of the device pointer table (it holds the addresses of the
        moment tensors, which this call replaces) and of the equal-steps check."""
        super().load_state_dict(state_dict)
        self._table_key = None
        self._keep = None
        self._steps_verified = False

    def add_param_group(self, param_group) -> None:
        super().add_param_group(param_group)
        self._table_key = None
        self._steps_verified = False

    def _build(self, params, grads, groups_of):
        dev = params[0].device
        n = len(params)
        fresh = 0
        for p in params:
            st = self.state[p]
            if len(st) == 0:
                fresh += 1
                st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif not st["step"].is_cuda:
                st["step"] = st["step"].to(device=dev, dtype=torch.float32)  # checkpoints written by the unfused optimizer
        # One device step counter (the first parameter's) drives the bias corrections of every entry, so all counters
        # must agree: a parameter that joins later (first gradient after others have stepped) or a checkpoint with
        # heterogeneous steps is refused instead of silently getting the wrong correction.  Host check only when the
        # set of states changed (fresh states, load_state_dict), never in the steady state.
        if 0 < fresh < n or not self._steps_verified:
            steps = torch.stack([self.state[p]["step"].reshape(()).to(device=dev, dtype=torch.float32) for p in params])
            if not bool((steps == steps[0]).all().item()):
                raise RuntimeError("FusedAdam: parameters carry different `step` counts (a parameter received its first gradient "
                                   "later than the others, or the checkpoint has heterogeneous steps); one shared counter drives "
                                   "the bias corrections here -- use torch.optim.Adam for this state")
            self._steps_verified = True
        numels = (C.c_int64 * n)(*[p.numel() for p in params])
        prefix = (C.c_int32 * (n + 1))()
        L.check(L.lib().msf_adam_plan(numels, n, prefix), "msf_adam_plan")
        table = torch.zeros((n, 8), dtype=torch.int64).pin_memory()  # pinned: the upload below never drains the stream
        any_ema = False
        for i, (p, g) in enumerate(zip(params, grads)):
            st = self.state[p]
            t = self._ema.get(id(p))
            any_ema |= t is not None
            table[i, 0], table[i, 1], table[i, 2], table[i, 3] = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            table[i, 4] = 0 if t is None else t.data_ptr()
            table[i, 5] = p.numel()
            sh = self._shadow.get(id(p))
            # int32 group in the low half, int32 shadow dtype in the high half (little-endian struct layout)
            table[i, 6] = groups_of[i] | ((L.dtype_code(sh.dtype) if sh is not None else 0) << 32)
            table[i, 7] = 0 if sh is None else sh.data_ptr()
        prefix_h = torch.tensor(list(prefix), dtype=torch.int32).pin_memory()
        self._keep = (table.to(dev, non_blocking=True), prefix_h.to(dev, non_blocking=True), int(prefix[n]), n, any_ema)
        self._pinned = (table, prefix_h)  # must outlive the asynchronous copies
        if any_ema and any(self._ema.get(id(p)) is None for p in params):
            raise ValueError("attach_ema must cover every stepped parameter or none (the kernel updates teachers for all entries)")

    @torch.no_grad()
    def step(self, closure=None, inv_scale: Optional[torch.Tensor] = None, found_inf: Optional[torch.Tensor] = None,
             grads: Optional[Sequence[torch.Tensor]] = None):
        """One fused step.  ``inv_scale`` / ``found_inf`` are device fp32 scalars (``1/scale`` and the overflow flag, see
        :meth:`check_grads`); under ``GradScaler.step(optimizer)`` (tools/ssl_train.py:473) they are taken from the
        ``grad_scale`` / ``found_inf`` attributes the scaler sets.  ``grads`` (one tensor per parameter, in
        ``param_groups`` order) replaces ``p.grad`` -- e.g. bf16 / fp16 gradient buffers, which autograd cannot attach
        to fp32 parameters."""
        gs, fi = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        if inv_scale is None and gs is not None:
            inv_scale = gs.to(torch.float32).reciprocal()
        if found_inf is None and fi is not None:
            found_inf = fi.to(torch.float32)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        given = None if grads is None else iter(grads)
        params, grads, groups_of = [], [], []
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise RuntimeError("FusedAdam implements plain Adam (amsgrad=False, maximize=False)")
            for p in group["params"]:
                g = p.grad if given is None else next(given)
                if g is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not _dense(p):
                    raise RuntimeError("FusedAdam needs dense CUDA fp32 parameters (there is no CPU fallback)")
                if g.shape != p.shape or not g.is_cuda:
                    raise RuntimeError("FusedAdam: gradient shape / device mismatch")
                params.append(p); grads.append(_like_layout(g, p)); groups_of.append(gi)
        if not params:
            return loss
        gdt = grads[0].dtype
        if any(g.dtype != gdt for g in grads):
            raise RuntimeError("FusedAdam: all gradients must share one dtype")
        key = self._key(params, grads)
        if key != self._table_key:  # gradients are re-allocated by zero_grad(set_to_none=True): rebuild the pointer table
            self._build(params, grads, groups_of)
            self._table_key = self._key(params, grads)  # _build may have created the moment tensors
        table, prefix, total_chunks, n, any_ema = self._keep
        g0 = self.param_groups[0]
        for group in self.param_groups:
            if (group["betas"], group["eps"], group["weight_decay"]) != (g0["betas"], g0["eps"], g0["weight_decay"]):
                raise RuntimeError("FusedAdam: groups may differ in lr only (the reference's three groups do)")
        dev = params[0].device
        lrs = tuple(float(g["lr"]) for g in self.param_groups)
        if getattr(self, "_lr_key", None) != lrs:  # learning rates change once per epoch at most (ssl_train.py:392-403)
            self._lr_host = torch.tensor(lrs, dtype=torch.float32).pin_memory()
            self._lr_dev, self._lr_key = self._lr_host.to(dev, non_blocking=True), lrs
        lr = self._lr_dev
        step_t = self.state[params[0]]["step"]
        # one shared device step counter drives the bias corrections; every parameter's `step` follows it
        if found_inf is not None:
            torch._foreach_add_([self.state[p]["step"] for p in params], (1.0 - found_inf).reshape(()).to(torch.float32))
        else:
            torch._foreach_add_([self.state[p]["step"] for p in params], 1.0)
        m = self._ema_momentum if any_ema else 0.0
        L.check(L.lib().msf_adam_multi(L.ptr(table), L.ptr(prefix), n, total_chunks, L.dtype_code(gdt), L.ptr(lr), float(g0["betas"][0]),
                                       float(g0["betas"][1]), float(g0["eps"]), float(g0["weight_decay"]), L.ptr(step_t), L.ptr(inv_scale),
                                       L.ptr(found_inf), int(any_ema), float(m), float(1.0 - m), L.stream_ptr()), "msf_adam_multi")
        L.launch_count += 1
        L.bump_param_epoch()  # parameters changed behind autograd's back (no `_version` bump): derived caches must refresh
        self._last = (lr, grads)  # keep the device lr array and the gradient tensors alive until the kernel has run
        return loss

    def _key(self, params, grads):
        """Everything the device pointer table bakes in: parameter, gradient, moment, teacher and shadow addresses."""
        ptrs = []
        for p, g in zip(params, grads):
            st = self.state.get(p, {})
            m, v = st.get("exp_avg"), st.get("exp_avg_sq")
            t, sh = self._ema.get(id(p)), self._shadow.get(id(p))
            ptrs.append((p.data_ptr(), g.data_ptr(), 0 if m is None else m.data_ptr(), 0 if v is None else v.data_ptr(),
                         0 if t is None else t.data_ptr(), 0 if sh is None else sh.data_ptr()))
        return tuple(ptrs)

    @torch.no_grad()
    def check_grads(self, found_inf: torch.Tensor) -> None:
        """found_inf (device fp32 scalar, zeroed by the caller) <- 1 if any gradient holds NaN / Inf; one launch."""
        params = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not params:
            return
        grads = [_like_layout(p.grad, p) for p in params]
        key = self._key(params, grads)
        if key != self._table_key:
            groups_of = [gi for gi, g in enumerate(self.param_groups) for p in g["params"] if p.grad is not None]
            self._build(params, grads, groups_of)
            self._table_key = self._key(params, grads)
        table, prefix, total_chunks, n, _ = self._keep
        L.check(L.lib().msf_grad_check_multi(L.ptr(table), L.ptr(prefix), n, total_chunks, L.dtype_code(grads[0].dtype), L.ptr(found_inf),
                                             L.stream_ptr()), "msf_grad_check_multi")
        L.launch_count += 1
