"""Checkpoint-format compatibility with the reference driver.

`tools/ssl_train.py:375-387` saves ``{"epoch", "arch", "state_dict", "optimizer", "scaler"}`` with ``torch.save`` where
``state_dict`` comes from the DDP-wrapped model (every key prefixed ``module.``); `tools/ssl_finetune.py:146-172` then
keeps the keys ``module.context_encoder.*`` / ``module.target_encoder.*`` (minus ``.fc``) and loads them into two
torchvision-layout ResNet encoders.  The functions below write and read exactly that layout for this repo's
``MSFWSI`` (whose parameter / buffer names equal the reference's), so checkpoints move freely between the two.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Optional, Tuple

import torch

PREFIX = "module."


def _unwrap(model: torch.nn.Module) -> torch.nn.Module:
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def reference_state_dict(model: torch.nn.Module) -> "OrderedDict[str, torch.Tensor]":
    """``model.state_dict()`` with the ``module.`` prefix the reference's DDP-wrapped model produces."""
    return OrderedDict((PREFIX + k, v) for k, v in _unwrap(model).state_dict().items())


def save_checkpoint(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, scaler=None,
                    epoch: int = 0, arch: str = "resnet18") -> Dict[str, Any]:
    """Write the dict of ssl_train.py:377-383 (``epoch`` is stored as given; the reference passes ``epoch + 1``)."""
    state = {"epoch": int(epoch), "arch": arch, "state_dict": reference_state_dict(model),
             "optimizer": optimizer.state_dict() if optimizer is not None else {},
             "scaler": scaler.state_dict() if scaler is not None else {}}
    torch.save(state, path)
    return state


def load_checkpoint(path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, scaler=None,
                    map_location="cpu", strict: bool = True, weights_only: bool = True, resume_eps: Optional[float] = None) -> int:
    """Resume like ssl_train.py:312-330: accepts checkpoints written by the reference (``module.`` prefix) or by a
    non-DDP run (no prefix).  Returns the stored epoch.

    ``weights_only=True`` (default) restricts unpickling to tensors and plain containers -- everything this format
    holds -- so a checkpoint file cannot execute code; pass False only for trusted files with exotic payloads.
    ``resume_eps``: the reference force-sets Adam's ``eps`` on every param group after a resume (``eps = 0.1``,
    ssl_train.py:325-326); pass ``resume_eps=0.1`` to reproduce a resumed reference run exactly, leave None to keep
    the optimizer's own eps."""
    ckpt = torch.load(path, map_location=map_location, weights_only=weights_only)
    sd = ckpt["state_dict"]
    if all(k.startswith(PREFIX) for k in sd):
        sd = OrderedDict((k[len(PREFIX):], v) for k, v in sd.items())
    _unwrap(model).load_state_dict(sd, strict=strict)
    if optimizer is not None and ckpt.get("optimizer"):
        optimizer.load_state_dict(ckpt["optimizer"])
        if resume_eps is not None:
            for group in optimizer.param_groups:
                group["eps"] = float(resume_eps)
    if scaler is not None and ckpt.get("scaler"):
        scaler.load_state_dict(ckpt["scaler"])
    return int(ckpt.get("epoch", 0))


def split_encoders(state_dict: Dict[str, torch.Tensor]) -> Tuple["OrderedDict[str, torch.Tensor]", "OrderedDict[str, torch.Tensor]"]:
    """The key surgery of ssl_finetune.py:153-170: (context, target) encoder state dicts in torchvision ResNet layout,
    without the ``fc`` entries."""
    ctx, tgt = OrderedDict(), OrderedDict()
    for k, v in state_dict.items():
        for name, out in (("context_encoder", ctx), ("target_encoder", tgt)):
            head = f"{PREFIX}{name}."
            if k.startswith(head) and not k.startswith(head + "fc"):
                out[k[len(head):]] = v
    return ctx, tgt
