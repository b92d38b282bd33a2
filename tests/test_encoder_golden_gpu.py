"""The encoder side (S1 stem layout, N1 fused batch norms / ReLU / residual / stem max-pool / folded average pool, cuDNN
convolutions) against the UNMODIFIED reference ResNet-18: tests/golden/encoder_resnet18.npz holds its fp64 pooled
features, running statistics and parameter-gradient norms / probes for closed-form weights and inputs
(oracle/make_golden_encoder.py, src/models/resnet.py:145-254 of the reference).  fp32 on the GPU (TF32 off): features
1e-4 relative, gradients 2e-3 by norm and probe (20 layers of batch statistics over 6 images amplify rounding)."""
import os

import numpy as np
import pytest
import torch

import msfwsi_b200 as M
from oracle import make_golden_encoder as G
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_encoder_matches_reference_resnet18_golden(golden_dir):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = np.load(os.path.join(golden_dir, "encoder_resnet18.npz"))
    enc = M.resnet18(pretrained=False, return_features=True, zero_init_residual=True)
    enc.fc = torch.nn.Identity()
    G.fill_closed_form(enc)
    enc = enc.to(DEV).to(memory_format=torch.channels_last).train()
    x = G.closed_form_input(torch.float32).to(DEV).contiguous(memory_format=torch.channels_last)
    feats = enc(x)
    for i, f in enumerate(feats):
        ref = torch.from_numpy(gold[f"feat{i}"])
        err = float((f.detach().double().cpu() - ref).norm() / ref.norm())
        assert err <= 1e-4, (i, err)
    G.loss_of(feats).backward()
    sd = enc.state_dict()
    for k in ("bn1.running_mean", "bn1.running_var", "layer1.0.bn2.running_var", "layer2.0.downsample.1.running_mean", "layer4.1.bn2.running_var"):
        ref = torch.from_numpy(gold["rs/" + k])
        assert torch.allclose(sd[k].double().cpu(), ref, rtol=1e-4, atol=1e-6), k
    n = 0
    for name, p in enc.named_parameters():
        g = p.grad.double().flatten().cpu()
        gn, gp = float(gold["gnorm/" + name]), float(gold["gprobe/" + name])
        probe = O.closed_form_tensor((g.numel(),), 7.0, 1.0).double()
        assert abs(float(g.norm()) - gn) <= 2e-3 * gn + 1e-9, (name, float(g.norm()), gn)
        assert abs(float((g * probe).sum()) - gp) <= 2e-3 * gn * float(probe.norm()) + 1e-9, name
        n += 1
    assert n == 60  # every parameter tensor of the ResNet-18 encoder (conv + bn; fc is Identity)
