"""GPU parity of the drop-in module's hot path (heads + loss) against the golden vectors generated from the
unmodified reference module, and against the CPU oracle at the config-1 batch (B=8)."""
import os
import warnings

import numpy as np
import pytest
import torch

import msfwsi_b200 as M
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
W = (0.1, 0.4, 0.7, 1.0)


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__()
        self.fc = torch.nn.Identity()


@pytest.fixture(scope="module")
def model():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = M.MSFWSI(lambda **kw: _Null(**kw), 4, 2048, 512, 0.5, False)
    missing = m.load_state_dict(O.closed_form_head_params(), strict=False)
    assert not missing.unexpected_keys
    return m.to(DEV).train()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "heads_loss_B3.npz")))


def _leaf_feats(gold, dtype=torch.float32):
    mk = lambda n: torch.from_numpy(gold[n]).to(dtype).to(DEV).requires_grad_(True)
    cf = [tuple(mk(f"ctx_f{v}_{l}") for l in range(4)) for v in (1, 2)]
    tf = [tuple(mk(f"tgt_f{v}_{l}") for l in range(4)) for v in (1, 2)]
    rev = [torch.from_numpy(gold["rev1"]), torch.from_numpy(gold["rev2"])]  # CPU int64, as the dataloader yields them
    return cf, tf, rev


def _pair_scale(gold):
    s = 0.0
    for b in ("ctx", "tgt", "ms"):
        for l in range(4):
            for pn, zn in (("p1", "z2"), ("p2", "z1")):
                p, z = torch.from_numpy(gold[f"{b}_{pn}_{l}"]).double(), torch.from_numpy(gold[f"{b}_{zn}_{l}"]).double()
                s += 0.5 * W[l] * abs(float(O.cosine_rows(p, z).mean()))
    return s


def test_heads_and_loss_match_reference_golden_fp32(model, gold):
    model.zero_grad(set_to_none=True)
    cf, tf, rev = _leaf_feats(gold)
    out = model.heads(cf[0], cf[1], tf[0], tf[1], rev)
    for bname, branch in zip(("ctx", "tgt", "ms"), out):
        for tname, tup in zip(("p1", "p2", "z1", "z2"), branch):
            for l, t in enumerate(tup):
                ref = torch.from_numpy(gold[f"{bname}_{tname}_{l}"])
                # B=3 batch-norm is ill-conditioned (fp32 GEMM/BN rounding is amplified), so the bar is on the
                # relative Frobenius error per tensor
                err = (t.detach().cpu().double() - ref.double()).norm() / ref.double().norm()
                assert err <= 2e-3, f"{bname}_{tname}_{l}: {err:.2e}"
                assert t.requires_grad == tname.startswith("p")  # z detached (backbone.py:188-191)
    loss = M.ssl_loss(out, W, mode="cosine")
    ref_loss = float(gold["loss"])
    assert abs(loss.item() - ref_loss) <= 1e-5 * max(abs(ref_loss), _pair_scale(gold))
    (loss * 1024.0).backward()  # GradScaler-style non-unit upstream gradient
    for v in range(2):
        for l in range(4):
            assert _cos(cf[v][l].grad, torch.from_numpy(gold[f"grad_ctx_f{v+1}_{l}"])) >= 0.9999
            assert _cos(tf[v][l].grad, torch.from_numpy(gold[f"grad_tgt_f{v+1}_{l}"])) >= 0.9999
    n_checked = 0
    for name, prm in model.named_parameters():
        key = "gnorm/" + name
        if key not in gold:
            continue
        assert prm.grad is not None, f"{name} got no gradient (DDP find_unused_parameters=False needs all)"
        g = prm.grad.double().flatten().cpu() / 1024.0
        probe = O.closed_form_tensor((g.numel(),), 7.0, 1.0).double()
        gn, gp = float(gold[key]), float(gold["gprobe/" + name])
        assert abs(g.norm().item() - gn) <= 2e-3 * gn + 1e-9, name
        assert abs((g * probe).sum().item() - gp) <= 2e-3 * gn * probe.norm().item() + 1e-9, name
        n_checked += 1
    assert n_checked == len(O.head_param_shapes())


def test_running_stats_follow_reference(model, gold):
    fresh = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(DEV).train()
    fresh.load_state_dict(O.closed_form_head_params(), strict=False)
    cf, tf, rev = _leaf_feats(gold)
    with torch.no_grad():
        fresh.heads(cf[0], cf[1], tf[0], tf[1], rev)
    sd = fresh.state_dict()
    keys = [k for k in gold if k.startswith("buf/")]
    assert keys
    for k in keys:
        assert torch.allclose(sd[k[4:]].cpu().double(), torch.from_numpy(gold[k]), rtol=1e-3, atol=1e-5), k


def test_config1_batch8_against_oracle(model):
    B, K = 8, 16
    g = torch.Generator().manual_seed(3407)
    rev = [torch.stack([O.jigsaw_indices(g, K)[1] for _ in range(B)]) for _ in range(2)]
    cf = [tuple(O.closed_form_tensor((B, d), 40 + 10 * v + l, 1.0).abs() for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
    tf = [tuple(O.closed_form_tensor((B * K, d), 80 + 10 * v + l, 1.0).abs() for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
    sd = {k: v.double() for k, v in O.closed_form_head_params().items()}
    ref_out = O.heads_forward(tuple(t.double() for t in cf[0]), tuple(t.double() for t in cf[1]), tuple(t.double() for t in tf[0]),
                              tuple(t.double() for t in tf[1]), rev[0], rev[1], sd)
    ref_loss = O.ssl_loss_block(ref_out, W)
    dcf = [tuple(t.to(DEV) for t in v) for v in cf]
    dtf = [tuple(t.to(DEV) for t in v) for v in tf]
    out = model.heads(dcf[0], dcf[1], dtf[0], dtf[1], [r.to(DEV) for r in rev])
    loss = M.ssl_loss(out, W)
    scale = sum(0.5 * W[l] * abs(float(O.cosine_rows(p, z).mean())) for br in ref_out for l, (p1, p2, z1, z2) in enumerate(zip(*br))
                for p, z in ((p1, z2), (p2, z1)))
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * max(abs(ref_loss.item()), scale)
    for br, rbr in zip(out, ref_out):
        for tup, rtup in zip(br, rbr):
            for t, r in zip(tup, rtup):
                assert (t.detach().cpu().double() - r).norm() / r.norm() <= 1e-3


def test_bf16_autocast_step_matches_torch_expression(model):
    """Under autocast(bf16) the fused loss must agree (<= 2e-3) with the reference's own expression evaluated by
    PyTorch on the same device and the same head outputs; gradients cosine >= 0.9999."""
    B, K = 16, 16
    g = torch.Generator().manual_seed(1)
    rev = [torch.stack([O.jigsaw_indices(g, K)[1] for _ in range(B)]) for _ in range(2)]
    mk = lambda shape, s: O.closed_form_tensor(shape, s, 1.0).abs().to(DEV).to(torch.bfloat16)
    cf = [tuple(mk((B, d), 300 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
    tf = [tuple(mk((B * K, d), 400 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
    cos = torch.nn.CosineSimilarity(dim=1)
    results = []
    for fused in (True, False):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model.heads(cf[0], cf[1], tf[0], tf[1], rev)
            if fused:
                loss = M.ssl_loss(out, W)
            else:  # tools/ssl_train.py:448-466 verbatim in structure
                loss = 0
                for branch in out:
                    for i, (p1, p2, z1, z2) in enumerate(zip(*branch)):
                        loss = loss + (-(cos(p1, z2).mean() + cos(p2, z1).mean()) * 0.5) * W[i]
        loss.backward()
        results.append((loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (lf, gf), (lt, gt) = results
    assert abs(lf - lt) <= 2e-3 * max(abs(lt), 0.1)
    assert set(gf) == set(gt)
    for n in gf:
        assert _cos(gf[n], gt[n]) >= 0.9999, n


def test_full_forward_with_resnet_encoders_smoke():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = M.MSFWSI(M.resnet18, 4).to(DEV).train()
    B = 2
    x1 = (torch.randn(B, 3, 64, 64, device=DEV), torch.randn(16 * B, 3, 64, 64, device=DEV))
    x2 = (torch.randn(B, 3, 64, 64, device=DEV), torch.randn(16 * B, 3, 64, 64, device=DEV))
    rev = [torch.stack([torch.randperm(16).argsort() for _ in range(B)]) for _ in range(2)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = m.forward_loss(x1, x2, rev)
    loss.backward()
    assert torch.isfinite(loss) and -6.7 <= loss.item() <= 6.7
    assert all(p.grad is not None for p in m.parameters())


@pytest.mark.parametrize("rows,fin,fout,bias", [(256, 64, 64, False), (4096, 512, 128, False), (4096, 128, 512, True), (48, 4608, 1152, False),
                                                 (3, 576, 144, True), (1000, 16, 64, True)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, None])
def test_tc_linear_matches_torch_linear(rows, fin, fout, bias, dtype):
    """A stand-alone head Linear on this repo's GEMM kernels vs nn.Linear with the same weights under the same autocast
    (dtype None = no autocast: exact fp32 on the SIMT kernel, 1e-5)."""
    torch.manual_seed(0)
    lin = M.TCLinear(fin, fout, bias=bias).to(DEV)
    ref = torch.nn.Linear(fin, fout, bias=bias).to(DEV)
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(rows, fin, device=DEV)
    x = x if dtype is None else x.to(dtype)
    w = torch.randn(rows, fout, device=DEV)
    outs = []
    for mod in (lin, ref):
        xi = x.clone().requires_grad_(True)
        if dtype is None:
            y = mod(xi)
        else:
            with torch.autocast("cuda", dtype=dtype):
                y = mod(xi)
        (y.float() * w).sum().backward()
        outs.append((y.detach().float(), xi.grad.float(), mod.weight.grad.clone(), None if not bias else mod.bias.grad.clone()))
    (y1, gx1, gw1, gb1), (y0, gx0, gw0, gb0) = outs
    tol = 1e-5 if dtype is None else 4e-3
    assert (y1 - y0).norm() / y0.norm() <= tol
    assert _cos(gx1, gx0) >= 0.9999 and _cos(gw1, gw0) >= 0.9999
    if dtype is None:
        assert (gw1 - gw0).norm() / gw0.norm() <= 1e-5 and (gx1 - gx0).norm() / gx0.norm() <= 1e-5
    if bias:
        assert _cos(gb1, gb0) >= 0.9999


def test_activation_checkpointing_variant_matches_plain_model():
    """use_checkpoint=True (backbone.py:103-127: checkpoint_wrapper over every Conv2d / Linear except the stem) gives the
    same loss and gradients as the plain model with identical weights."""
    import copy
    import warnings
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        plain = M.MSFWSI(M.resnet18, 2, 2048, 512, 0.5, False).to(DEV).to(memory_format=torch.channels_last).train()
        ckpt = M.MSFWSI(M.resnet18, 2, 2048, 512, 0.5, True).to(DEV).to(memory_format=torch.channels_last).train()
    sd = {k.replace("_checkpoint_wrapped_module.", ""): v for k, v in ckpt.state_dict().items()}
    assert set(sd) == set(plain.state_dict())
    ckpt.load_state_dict({k: plain.state_dict()[k.replace("_checkpoint_wrapped_module.", "")] for k in ckpt.state_dict()})
    B, K = 4, 4
    g = torch.Generator().manual_seed(4)
    mk = lambda n: torch.randn(n, 3, 64, 64, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    x1, x2 = (mk(B), mk(B * K)), (mk(B), mk(B * K))
    rev = [torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).to(DEV) for _ in range(2)]
    losses = []
    for m in (plain, ckpt):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = m.forward_loss(x1, x2, rev, W, mode="cosine")
        loss.backward()
        losses.append(float(loss))
    assert abs(losses[0] - losses[1]) <= 2e-3 * max(1.0, abs(losses[0]))
    gp = dict(plain.named_parameters())
    for n, p in ckpt.named_parameters():
        q = gp[n.replace("_checkpoint_wrapped_module.", "")]
        assert p.grad is not None and q.grad is not None, n
    assert _cos(ckpt.context_encoder.conv1.weight.grad, plain.context_encoder.conv1.weight.grad) >= 0.99
