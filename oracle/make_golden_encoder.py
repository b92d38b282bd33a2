"""Generate tests/golden/encoder_resnet18.npz from the UNMODIFIED reference encoder (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    python oracle/make_golden_encoder.py [--ref /root/reference] [--out tests/golden]

Imports ``src.models.resnet.resnet18`` from the reference checkout (``pretrained=False, return_features=True,
zero_init_residual=True`` as src/models/backbone.py:58-65 builds it, ``fc = Identity``), fills every parameter with a closed
form (``oracle.msf_oracle.closed_form_tensor`` scaled per tensor: no RNG stream, nothing large to ship), runs a train-mode
forward + backward in fp64 on a closed-form input batch and stores the four pooled pyramid features, the batch-norm running
statistics of a few layers and, per parameter, the gradient norm and a probe projection.  tests/test_encoder_golden_gpu.py
replays the same recipe through this repo's encoder (S1 stem layout, N1 fused batch norms, cuDNN convolutions) on the GPU.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import msf_oracle as O  # noqa: E402

N, IMG = 6, 64


def fill_closed_form(model: torch.nn.Module) -> None:
    """Deterministic weights with trained-network-like scales; shared with the GPU test (which imports this module)."""
    with torch.no_grad():
        for i, (name, p) in enumerate(sorted(model.named_parameters())):
            if p.dim() == 4:      # convolution: He-like scale
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                p.copy_(O.closed_form_tensor(tuple(p.shape), 100.0 + i, (6.0 / fan_in) ** 0.5).to(p.dtype))
            elif name.endswith("weight"):  # batch-norm gamma in [0.5, 1.5], a few of them negative
                g = O.closed_form_tensor(tuple(p.shape), 100.0 + i, 0.5).to(p.dtype) + 1.0
                g[::7] = -g[::7]
                p.copy_(g)
            else:                 # batch-norm beta
                p.copy_(O.closed_form_tensor(tuple(p.shape), 100.0 + i, 0.3).to(p.dtype))


def closed_form_input(dtype=torch.float64) -> torch.Tensor:
    return (O.closed_form_tensor((N, 3, IMG, IMG), 9.0, 1.7) + 0.2).to(dtype)


def loss_of(features) -> torch.Tensor:
    return sum((f * O.closed_form_tensor(tuple(f.shape), 20.0 + i, 1.0).to(f.dtype).to(f.device)).sum() / f.shape[1] for i, f in enumerate(features))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from src.models import resnet as ref_resnet  # the unmodified reference
    enc = ref_resnet.resnet18(pretrained=False, return_features=True, zero_init_residual=True)
    enc.fc = torch.nn.Identity()
    enc = enc.double().train()
    fill_closed_form(enc)
    x = closed_form_input()
    feats = enc(x)
    loss_of(feats).backward()
    out = {f"feat{i}": f.detach().numpy() for i, f in enumerate(feats)}
    sd = enc.state_dict()
    for k in ("bn1.running_mean", "bn1.running_var", "layer1.0.bn2.running_var", "layer2.0.downsample.1.running_mean", "layer4.1.bn2.running_var"):
        out["rs/" + k] = sd[k].numpy()
    for name, p in enc.named_parameters():
        g = p.grad.flatten()
        out["gnorm/" + name] = np.array(float(g.norm()))
        out["gprobe/" + name] = np.array(float((g * O.closed_form_tensor((g.numel(),), 7.0, 1.0).double()).sum()))
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "encoder_resnet18.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes; features {[tuple(f.shape) for f in feats]}")


if __name__ == "__main__":
    main()
