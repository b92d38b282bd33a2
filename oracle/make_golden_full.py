"""Generate tests/golden/full_module_B4.npz from the UNMODIFIED reference module end to end (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden_full.py [--ref /root/reference] [--out tests/golden]

The whole drop-in surface at once: ``src.models.backbone.MSFWSI`` with the reference's own ``resnet18`` encoders
(``pretrained=False`` through a wrapper: the hard-coded download of backbone.py:58-63 cannot run offline), closed-form
weights everywhere (encoders: ``make_golden_encoder.fill_closed_form``; heads: ``msf_oracle.closed_form_head_params``),
closed-form 64x64 views for B = 4 samples, train-mode forward, the loss block of tools/ssl_train.py:448-466 (restated:
that file needs albumentations), backward -- in fp64.  Stored: the loss, a few output tensors and, for all 264 parameter
tensors, the gradient norm and a probe projection.  tests/test_full_module_golden_gpu.py replays it on the GPU.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import make_golden_encoder as GE  # noqa: E402
from oracle import msf_oracle as O  # noqa: E402
from oracle.make_golden import FUSER_WEIGHTS, reference_loss_block  # noqa: E402

B, K, IMG = 4, 16, 64


def fill(model: nn.Module) -> None:
    """Closed-form weights for the whole module (shared with the GPU test)."""
    GE.fill_closed_form(model.context_encoder)
    GE.fill_closed_form(model.target_encoder)
    with torch.no_grad():  # the two encoders must differ: shift the target encoder's convolution weights
        for i, (name, p) in enumerate(sorted(model.target_encoder.named_parameters())):
            if p.dim() == 4:
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                p.copy_(O.closed_form_tensor(tuple(p.shape), 300.0 + i, (6.0 / fan_in) ** 0.5).to(p.dtype))
    sd = {k: v.to(next(model.parameters()).dtype) for k, v in O.closed_form_head_params().items()}
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys


def inputs(dtype=torch.float64):
    ctx = [(O.closed_form_tensor((B, 3, IMG, IMG), 1000 + v, 1.5) + 0.1).to(dtype) for v in range(2)]
    tgt = [(O.closed_form_tensor((B * K, 3, IMG, IMG), 2000 + v, 1.5) + 0.1).to(dtype) for v in range(2)]
    g = torch.Generator().manual_seed(3407)  # tools/ssl_train.py:572-574
    rev = [torch.stack([O.jigsaw_indices(g, K)[1] for _ in range(B)]) for _ in range(2)]
    return ctx, tgt, rev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from src.models import resnet as ref_resnet
    from src.models.backbone import MSFWSI  # the unmodified reference module
    model = MSFWSI(lambda **kw: ref_resnet.resnet18(**{**kw, "pretrained": False}), 4, 2048, 512, 0.5, False).double().train()
    fill(model)
    ctx, tgt, rev = inputs()
    outputs = model((ctx[0], tgt[0]), (ctx[1], tgt[1]), [rev[0], rev[1]])
    loss = reference_loss_block(outputs, nn.CosineSimilarity(dim=1), FUSER_WEIGHTS)  # tools/ssl_train.py:422, 448-466
    loss.backward()
    cos = nn.CosineSimilarity(dim=1)
    scale = sum(0.5 * FUSER_WEIGHTS[l] * (abs(cos(p1, z2).mean().item()) + abs(cos(p2, z1).mean().item()))
                for branch in outputs for l, (p1, p2, z1, z2) in enumerate(zip(*branch)))
    gold = {"loss": np.float64(loss.item()), "pair_scale": np.float64(scale)}  # the 24 signed terms cancel: errors are relative to their magnitudes
    for bname, branch in zip(("ctx", "tgt", "ms"), outputs):
        for tname, tup in zip(("p1", "z2"), (branch[0], branch[3])):
            gold[f"{bname}_{tname}_3"] = tup[3].detach().numpy().astype(np.float32)  # level 4 (dim 512 / 4608)
    n = 0
    for name, prm in model.named_parameters():
        g = prm.grad.flatten()
        gold["gnorm/" + name] = np.float64(g.norm().item())
        gold["gprobe/" + name] = np.float64((g * O.closed_form_tensor((g.numel(),), 7.0, 1.0).double()).sum().item())
        n += 1
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "full_module_B4.npz")
    np.savez_compressed(path, **gold)
    print(f"wrote {path}: loss {loss.item():+.9f}, {n} parameter tensors, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
