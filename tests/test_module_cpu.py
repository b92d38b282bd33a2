"""Drop-in contract of msfwsi_b200.MSFWSI against the reference's state-dict (fixture generated from the
unmodified reference module by oracle/make_golden.py) and the optimizer prefix filter of ssl_train.py:281-307."""
import json
import os
import warnings

import pytest
import torch

import msfwsi_b200 as M


@pytest.fixture(scope="module")
def model():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return M.MSFWSI(M.resnet18, 4, 2048, 512, 0.5, False)


def test_state_dict_contract_matches_reference(model, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "state_dict_contract.json")))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref["state_dict"]
    assert sum(p.numel() for p in model.parameters()) == ref["n_params"] == 123551584
    assert len(list(model.parameters())) == ref["n_param_tensors"] == 264


def test_optimizer_prefix_filter_covers_every_parameter(model):
    # tools/ssl_train.py:281-307 builds three Adam groups by name prefix and drops anything else
    names = [n for n, _ in model.named_parameters()]
    assert all(n.startswith(("context_", "target_", "inter_")) for n in names)


def test_constructor_semantics(model):
    assert model.K == 16 and model.n_keep == 8
    assert [int(d) for d in model.ms_inter_dim] == [576, 1152, 2304, 4608]
    assert isinstance(model.context_encoder.fc, torch.nn.Identity)
    m2 = M.MSFWSI(lambda **kw: M.resnet18(**{**kw, "pretrained": False}), 2, mask_ratio=0.25)
    assert m2.K == 4 and m2.n_keep == 3


def test_survives_sync_batchnorm_conversion(model):
    import copy
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(model))
    n_sync = sum(isinstance(m, torch.nn.SyncBatchNorm) for m in conv.modules())
    # nothing is left for torch to convert: the 48 head BNs are FusedBatchNorm1d and the 40 encoder BNs FusedBatchNorm2d,
    # which reduce their statistics across ranks themselves with SyncBatchNorm semantics (the reference converts all 88,
    # SURVEY 0); the call must still go through unchanged
    assert n_sync == 0
    from msfwsi_b200.module import FusedBatchNorm1d
    from msfwsi_b200.resnet import FusedBatchNorm2d
    assert sum(isinstance(m, FusedBatchNorm2d) for m in conv.modules()) == 40
    assert sum(isinstance(m, FusedBatchNorm1d) for m in conv.modules()) == 48
    assert set(conv.state_dict()) == set(model.state_dict())


def test_encoder_feature_contract():
    enc = M.resnet18(return_features=True, zero_init_residual=True)
    enc.fc = torch.nn.Identity()
    enc.eval()
    with torch.no_grad():
        feats = enc(torch.randn(2, 3, 64, 64))
    assert [tuple(f.shape) for f in feats] == [(2, 64), (2, 128), (2, 256), (2, 512)]
    assert all((f >= 0).all() for f in feats)  # pooled post-ReLU maps


def test_forward_requires_cuda_library_path(model):
    x = (torch.randn(2, 3, 64, 64), torch.randn(32, 3, 64, 64))
    rev = [torch.arange(16).repeat(2, 1), torch.arange(16).repeat(2, 1)]
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x, x, rev)  # the hot path has no CPU fallback


def test_bad_loss_mode():
    with pytest.raises(ValueError):
        M.ssl_loss((), mode="nope")


def test_checkpoint_layout_matches_reference_driver(model, golden_dir, tmp_path):
    """ssl_train.py:375-387 / ssl_finetune.py:146-172: dict fields, `module.` prefix, encoder key surgery, resume."""
    from msfwsi_b200 import checkpoint as CK
    ref_keys = set(json.load(open(os.path.join(golden_dir, "state_dict_contract.json")))["state_dict"])
    opt = torch.optim.Adam([p for p in model.parameters()], lr=1e-3)
    path = str(tmp_path / "checkpoint_0000.pth.tar")
    state = CK.save_checkpoint(path, model, opt, None, epoch=1, arch="resnet18")
    assert set(state) == {"epoch", "arch", "state_dict", "optimizer", "scaler"}
    assert {k[len("module."):] for k in state["state_dict"]} == ref_keys and all(k.startswith("module.") for k in state["state_dict"])
    ctx, tgt = CK.split_encoders(torch.load(path, weights_only=False)["state_dict"])
    plain = M.resnet18()  # torchvision-layout encoder, as the fine-tune stage builds
    want = {k for k in plain.state_dict() if not k.startswith("fc")}
    assert set(ctx) == want and set(tgt) == want
    plain.load_state_dict(ctx, strict=False)
    # resume into a fresh model: identical tensors, epoch restored, also from an un-prefixed checkpoint
    import copy
    fresh = copy.deepcopy(model)
    with torch.no_grad():
        for p in fresh.parameters():
            p.add_(1.0)
    assert CK.load_checkpoint(path, fresh) == 1
    assert all(torch.equal(a, b) for a, b in zip(fresh.state_dict().values(), model.state_dict().values()))
    torch.save({"epoch": 3, "state_dict": model.state_dict()}, path)
    assert CK.load_checkpoint(path, fresh) == 3


def test_stem_space_to_depth_weight_identity_on_cpu():
    """S1 host logic: conv7x7/2(x, w) == conv4x4/1(pad + pixel_unshuffle(x), stem_s2d_weight(w)) (resnet.py:155), and the
    rearrangement is differentiable back to the (O, 3, 7, 7) parameter."""
    import torch.nn.functional as F
    from msfwsi_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 16, 20, dtype=torch.float64, generator=g)
    w = torch.randn(5, 3, 7, 7, dtype=torch.float64, generator=g, requires_grad=True)
    xs = F.pad(F.pixel_unshuffle(F.pad(x, (3, 3, 3, 3)), 2), (0, 0, 0, 0, 0, 4))  # what msf_stem_s2d produces on the device
    y = F.conv2d(xs, ops.stem_s2d_weight(w))
    y_ref = F.conv2d(x, w, None, 2, 3)
    assert torch.allclose(y, y_ref, rtol=1e-12, atol=1e-12)
    gy = torch.randn(y.shape, dtype=torch.float64, generator=g)
    (gw,) = torch.autograd.grad(y, w, gy, retain_graph=True)
    (gw_ref,) = torch.autograd.grad(y_ref, w, gy)
    assert gw.shape == (5, 3, 7, 7) and torch.allclose(gw, gw_ref, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        ops.stem_s2d_weight(torch.zeros(4, 3, 3, 3))


def test_product_ops_refuse_cpu_tensors():
    """No CPU fallback anywhere in the product path: the operators and the optimizer raise on CPU tensors."""
    from msfwsi_b200 import FusedAdam, ops
    with pytest.raises(RuntimeError):
        ops.bn_act2d(torch.zeros(2, 8, 2, 2), None, None, None, None)
    with pytest.raises(RuntimeError):
        ops.stem_s2d(torch.zeros(1, 3, 4, 4))
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p], lr=1e-3).step()


def test_encoder_module_wiring_matches_reference_resnet18_golden(golden_dir):
    """Same recipe as the GPU test, on CPU in fp64 (plain ATen path of FusedBatchNorm2d, plain stem convolution): pins the
    module wiring -- parameter names, block structure, pooled features -- to the unmodified reference ResNet-18."""
    import numpy as np
    from oracle import make_golden_encoder as G
    gold = np.load(os.path.join(golden_dir, "encoder_resnet18.npz"))
    enc = M.resnet18(pretrained=False, return_features=True, zero_init_residual=True)
    enc.fc = torch.nn.Identity()
    enc = enc.double().train()
    G.fill_closed_form(enc)
    feats = enc(G.closed_form_input())
    for i, f in enumerate(feats):
        assert torch.allclose(f.detach(), torch.from_numpy(gold[f"feat{i}"]), rtol=1e-9, atol=1e-10), i
    G.loss_of(feats).backward()
    for name, p in enc.named_parameters():
        assert abs(float(p.grad.norm()) - float(gold["gnorm/" + name])) <= 1e-8 * float(gold["gnorm/" + name]) + 1e-12, name
