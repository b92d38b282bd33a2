"""The whole drop-in module -- both ResNet-18 encoders (S1 + N1 + cuDNN), gather / concat (A1), heads (fp32 Linears + N1),
fused cosine loss (L1) and every backward -- against the UNMODIFIED reference `MSFWSI` + loss block run end to end in fp64
(tests/golden/full_module_B4.npz from oracle/make_golden_full.py; src/models/backbone.py:34-222, tools/ssl_train.py:448-466).
fp32 on the GPU with TF32 off.  B = 4 makes the context heads' batch norms ill-conditioned (4 rows), which amplifies fp32
rounding: loss within 1e-3 of the magnitude of its 24 terms, stored outputs 1e-2 relative, every one of the 264 parameter
gradients within 3e-2 by norm and probe (all of them finite and non-zero -- DDP's find_unused_parameters=False needs that)."""
import os

import numpy as np
import pytest
import torch

import msfwsi_b200 as M
from oracle import make_golden_full as G
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_full_module_matches_reference_end_to_end(golden_dir):
    import warnings
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = np.load(os.path.join(golden_dir, "full_module_B4.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = M.MSFWSI(M.resnet18, 4, 2048, 512, 0.5, False)
    G.fill(model)
    model = model.to(DEV).to(memory_format=torch.channels_last).train()
    ctx, tgt, rev = G.inputs(torch.float32)
    cl = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)
    out = model((cl(ctx[0]), cl(tgt[0])), (cl(ctx[1]), cl(tgt[1])), [rev[0], rev[1]])  # jigsaw_idx stays on the CPU like the dataloader's
    loss = M.ssl_loss(out, M.DEFAULT_FUSER_WEIGHTS, mode="cosine")
    assert abs(loss.item() - float(gold["loss"])) <= 1e-3 * float(gold["pair_scale"]), (loss.item(), float(gold["loss"]), float(gold["pair_scale"]))
    for bi, bname in enumerate(("ctx", "tgt", "ms")):
        for tname, ti in (("p1", 0), ("z2", 3)):
            ref = torch.from_numpy(gold[f"{bname}_{tname}_3"]).double()
            got = out[bi][ti][3].detach().double().cpu()
            assert float((got - ref).norm() / ref.norm()) <= 1e-2, (bname, tname)
    loss.backward()
    n = 0
    for name, prm in model.named_parameters():
        assert prm.grad is not None, name
        g = prm.grad.double().flatten().cpu()
        gn, gp = float(gold["gnorm/" + name]), float(gold["gprobe/" + name])
        probe = O.closed_form_tensor((g.numel(),), 7.0, 1.0).double()
        assert abs(float(g.norm()) - gn) <= 3e-2 * gn + 1e-12, (name, float(g.norm()), gn)
        assert abs(float((g * probe).sum()) - gp) <= 3e-2 * gn * float(probe.norm()) + 1e-12, name
        n += 1
    assert n == 264
