"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

It imports ``src.models.backbone.MSFWSI`` and ``src.utils.data.bcss.blockshaped`` from the
reference checkout, drives them with closed-form inputs and parameters
(``oracle.msf_oracle.closed_form_*`` -- no dependence on torch's RNG stream, so the 99.8 M
head parameters need not be shipped), restates the 19-line loss block of
``tools/ssl_train.py:448-466`` (that file cannot be imported: albumentations is absent) and
stores the outputs.  /root/reference does not exist on the GPU box, so tests read only the
committed fixtures.

The reference's ``base_encoder`` argument is a callable (backbone.py:58-63); we pass a tiny
deterministic stub encoder so the fixture pins exactly the hot path (everything after the
encoder calls) and not 11 M ResNet weights.
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

from oracle import msf_oracle as O  # noqa: E402

B, K, IMG = 3, 16, 8
FUSER_WEIGHTS = (0.1, 0.4, 0.7, 1.0)  # tools/ssl_train.py:623-625


class StubEncoder(nn.Module):
    """Deterministic stand-in for ``resnet18(return_features=True)``: four non-negative
    pooled feature vectors of widths 64/128/256/512 (cf. src/models/resnet.py:244-254)."""

    def __init__(self, salt: float, **_ignored):
        super().__init__()
        self.fc = nn.Identity()
        self.mats = nn.ParameterList(
            [nn.Parameter(O.closed_form_tensor((d, 3 * IMG * IMG), salt + i, 0.3).double()) for i, d in enumerate(O.INTER_DIM)]
        )

    def forward(self, x):
        flat = x.flatten(1)
        return tuple(torch.relu(flat @ m.t()) for m in self.mats)


def reference_loss_block(outputs, contrast_loss, fuser_weights):
    # restated verbatim in structure from tools/ssl_train.py:448-466
    total = 0
    for branch in outputs:
        part = 0
        for i, (p1, p2, z1, z2) in enumerate(zip(*branch)):
            part += (-(contrast_loss(p1, z2).mean() + contrast_loss(p2, z1).mean()) * 0.5) * fuser_weights[i]
        total = total + part
    return total


def build_inputs():
    ctx = [O.closed_form_tensor((B, 3, IMG, IMG), 1000 + v, 1.0).double() for v in range(2)]
    tgt = [O.closed_form_tensor((B * K, 3, IMG, IMG), 2000 + v, 1.0).double() for v in range(2)]
    g = torch.Generator().manual_seed(3407)  # tools/ssl_train.py:572-574
    rev = []
    for _ in range(2):
        rows = [O.jigsaw_indices(g, K)[1] for _ in range(B)]
        rev.append(torch.stack(rows))
    return ctx, tgt, rev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from src.models.backbone import MSFWSI  # the unmodified reference module
    from src.utils.data.bcss import blockshaped as ref_blockshaped

    os.makedirs(args.out, exist_ok=True)
    torch.manual_seed(3407)

    salts = iter((11.0, 29.0))
    model = MSFWSI(lambda **kw: StubEncoder(next(salts), **kw), 4, 2048, 512, 0.5, False).double()
    sd_heads = O.closed_form_head_params()
    missing = model.load_state_dict({k: v.double() for k, v in sd_heads.items()}, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all(k.split(".")[0] in ("context_encoder", "target_encoder") or "running_" in k or "num_batches" in k
               for k in missing.missing_keys), missing.missing_keys
    model.train()

    ctx, tgt, rev = build_inputs()
    with torch.no_grad():
        cf = [model.context_encoder(c) for c in ctx]
        tf = [model.target_encoder(t) for t in tgt]

    # hot-path inputs as leaf tensors so feature gradients are recorded
    feats = {}
    for v in range(2):
        for l in range(4):
            feats[f"ctx_f{v+1}_{l}"] = cf[v][l].float().double().requires_grad_(True)  # fp32-representable
            feats[f"tgt_f{v+1}_{l}"] = tf[v][l].float().double().requires_grad_(True)

    class Feed(nn.Module):  # replays the stored features through the reference forward
        def __init__(self, per_call):
            super().__init__()
            self.fc = nn.Identity()
            self.per_call, self.i = per_call, 0

        def forward(self, x):
            out = self.per_call[self.i % len(self.per_call)]
            self.i += 1
            return out

    model.context_encoder = Feed([tuple(feats[f"ctx_f{v+1}_{l}"] for l in range(4)) for v in range(2)])
    model.target_encoder = Feed([tuple(feats[f"tgt_f{v+1}_{l}"] for l in range(4)) for v in range(2)])

    outputs = model((ctx[0], tgt[0]), (ctx[1], tgt[1]), [rev[0], rev[1]])
    contrast = nn.CosineSimilarity(dim=1)  # tools/ssl_train.py:422
    loss = reference_loss_block(outputs, contrast, FUSER_WEIGHTS)
    loss.backward()

    gold = {"loss": np.float64(loss.item()), "B": np.int64(B), "rev1": rev[0].numpy(), "rev2": rev[1].numpy()}
    for name, t in feats.items():
        gold[name] = t.detach().numpy().astype(np.float32)
        gold["grad_" + name] = t.grad.numpy().astype(np.float64)
    for bname, branch in zip(("ctx", "tgt", "ms"), outputs):
        for tname, tup in zip(("p1", "p2", "z1", "z2"), branch):
            for l, t in enumerate(tup):
                gold[f"{bname}_{tname}_{l}"] = t.detach().numpy().astype(np.float32)
    # per-parameter gradient summaries (norm + probe dot) instead of 99.8 M values
    for name, prm in model.named_parameters():
        if prm.grad is None or name.split(".")[0].endswith("encoder"):
            continue
        gflat = prm.grad.flatten()
        probe = O.closed_form_tensor((gflat.numel(),), 7.0, 1.0).double()
        gold["gnorm/" + name] = np.float64(gflat.norm().item())
        gold["gprobe/" + name] = np.float64((gflat * probe).sum().item())
    # running statistics after one train-mode forward (2 calls per projector: view 1 then 2)
    for name, buf in model.named_buffers():
        parts = name.split(".")
        if "running_" in name and parts[0].endswith(("projector", "predictor")) and parts[1] == "0":  # level 0 only
            gold["buf/" + name] = buf.numpy().astype(np.float64)

    # fp32 run of the same module (anchors the fp32 tolerance statement)
    model32 = model.float()
    f32 = {k: v.detach().float() for k, v in feats.items()}
    model32.context_encoder = Feed([tuple(f32[f"ctx_f{v+1}_{l}"] for l in range(4)) for v in range(2)])
    model32.target_encoder = Feed([tuple(f32[f"tgt_f{v+1}_{l}"] for l in range(4)) for v in range(2)])
    with torch.no_grad():
        out32 = model32((ctx[0].float(), tgt[0].float()), (ctx[1].float(), tgt[1].float()), [rev[0], rev[1]])
        gold["loss_fp32"] = np.float64(reference_loss_block(out32, contrast, FUSER_WEIGHTS).item())
    np.savez_compressed(os.path.join(args.out, "heads_loss_B3.npz"), **gold)

    # blockshaped: crop coordinates (bcss.py:203-216) on a labelled array
    arr = np.arange(8 * 12 * 3, dtype=np.int64).reshape(8, 12, 3)
    bs = {"arr": arr, "tiles_4x4": ref_blockshaped(arr, 4, 4), "tiles_2x6": ref_blockshaped(arr, 2, 6)}
    big = np.arange(1024 * 1024 * 3, dtype=np.int64).reshape(1024, 1024, 3) % 251
    tiles = ref_blockshaped(big.astype(np.uint8), 256, 256)
    assert tiles.shape == (16, 256, 256, 3)  # bcss.py:176
    bs["tiles_1024_corner_sums"] = tiles.reshape(16, -1).astype(np.int64).sum(axis=1)
    bs["tiles_1024_first_px"] = tiles[:, 0, 0, :].astype(np.int64)
    np.savez_compressed(os.path.join(args.out, "blockshaped.npz"), **bs)

    # hooknet centre crop (hooknet.py:29-32) -- the file itself needs segmentation_models_pytorch,
    # which is absent, so the slice expression is evaluated here exactly as written there.
    x = O.closed_form_tensor((2, 128, 32, 32), 5.0, 1.0)
    crop = x[:, :, 16 - 4: 16 + 4, 16 - 4: 16 + 4]
    np.savez_compressed(os.path.join(args.out, "hooknet_crop.npz"), x_salt=np.float64(5.0), crop=crop.numpy())
    # state-dict key/shape contract of the full reference module (real ResNet-18 encoders,
    # pretrained=False wrapper because the hard-coded pretrained=True cannot download offline)
    import json
    from src.models import resnet as ref_resnet
    full = MSFWSI(lambda **kw: ref_resnet.resnet18(**{**kw, "pretrained": False}), 4, 2048, 512, 0.5, False)
    contract = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(args.out, "state_dict_contract.json"), "w") as fh:
        json.dump({"n_params": sum(p.numel() for p in full.parameters()), "n_param_tensors": len(list(full.parameters())),
                   "state_dict": contract}, fh, indent=0)
    print("wrote", sorted(os.listdir(args.out)))
    print("loss fp64", loss.item(), "loss fp32", gold["loss_fp32"])


if __name__ == "__main__":
    main()
