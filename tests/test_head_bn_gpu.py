"""The grouped head batch-norm kernels (csrc/head_bn.cu) one by one against torch in fp64 on the same inputs:
statistics -> finalize (scale / shift / mean / invstd / running statistics), apply (+ L2-normalised rows), backward
reduce -> finalize -> element-wise.  torch's BatchNorm1d / ReLU / autograd are the oracle at the reference's call sites
(src/models/backbone.py:15-16, 18-19, 21, 28-29).  Tolerances: fp32 1e-5 (outputs) / 2e-5 (gradients), bf16 one output ulp."""
import ctypes as C

import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
P2 = C.c_void_p * 2


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def _forward(y, gamma, beta, relu, rm, rv):
    """stats -> finalize -> apply for the two views stacked in y (2, rows, C); returns a, (sc, sh, mu, istd)."""
    _, rows, Cc = y.shape
    code = L.dtype_code(y.dtype)
    groups = (rows + 31) // 32
    cs = torch.empty(2, groups, 2, Cc, device=DEV)
    mats = (L.HeadMat * 2)(*[L.HeadMat(y[v].data_ptr(), cs[v].data_ptr(), rows, Cc) for v in range(2)])
    L.check(L.lib().msf_head_bn_stats(mats, 2, code, L.stream_ptr()), "stats")
    sc, sh, mu, istd = (torch.empty(2, Cc, device=DEV) for _ in range(4))
    it = (L.HeadBnItem * 1)(L.HeadBnItem(P2(cs[0].data_ptr(), cs[1].data_ptr()), P2(sc[0].data_ptr(), sc[1].data_ptr()), P2(sh[0].data_ptr(), sh[1].data_ptr()),
                                         P2(mu[0].data_ptr(), mu[1].data_ptr()), P2(istd[0].data_ptr(), istd[1].data_ptr()), L.ptr(gamma), L.ptr(beta),
                                         L.ptr(rm), L.ptr(rv), rows, Cc, 2, 1))
    L.check(L.lib().msf_head_bn_finalize(it, 1, 1e-5, 0.1, 1, 0, 1, 0, 0, 0, 1, L.stream_ptr()), "finalize")
    a = torch.empty_like(y)
    ap = (L.HeadApplyItem * 2)(*[L.HeadApplyItem(y[v].data_ptr(), a[v].data_ptr(), 0, 0, sc[v].data_ptr(), sh[v].data_ptr(), mu[v].data_ptr(), rows, Cc, int(relu), 0) for v in range(2)])
    L.check(L.lib().msf_head_bn_apply(ap, 2, code, 1e-8, L.stream_ptr()), "apply")
    return a, (sc, sh, mu, istd)


def _backward(g, y, stats, relu, affine):
    sc, sh, mu, istd = stats
    _, rows, Cc = y.shape
    code = L.dtype_code(y.dtype)
    rbs = (rows + 255) // 256
    part = torch.empty(2, rbs, 2, Cc, device=DEV)
    c1, c2 = torch.empty(2, Cc, device=DEV), torch.empty(2, Cc, device=DEV)
    dgam, dbet = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV)
    dy = torch.empty_like(y)
    items = (L.HeadBwdItem * 2)(*[L.HeadBwdItem(g[v].data_ptr(), y[v].data_ptr(), dy[v].data_ptr(), part[v].data_ptr(), sc[v].data_ptr(), sh[v].data_ptr(),
                                                mu[v].data_ptr(), istd[v].data_ptr(), c1[v].data_ptr(), c2[v].data_ptr(), rows, Cc, int(relu), 1) for v in range(2)])
    L.check(L.lib().msf_head_bn_bwd_reduce(items, 2, code, L.stream_ptr()), "reduce")
    fin = (L.HeadBwdFinItem * 1)(L.HeadBwdFinItem(P2(part[0].data_ptr(), part[1].data_ptr()), P2(c1[0].data_ptr(), c1[1].data_ptr()), P2(c2[0].data_ptr(), c2[1].data_ptr()),
                                                  dgam.data_ptr() if affine else 0, dbet.data_ptr() if affine else 0, rows, Cc, 2, 0))
    L.check(L.lib().msf_head_bn_bwd_finalize(fin, 1, 1, 0, 1, 0, 0, 0, 1, L.stream_ptr()), "bwd finalize")
    L.check(L.lib().msf_head_bn_bwd_elemt(items, 2, code, L.stream_ptr()), "elemt")
    return dy, dgam, dbet


@pytest.mark.parametrize("rows,Cc", [(256, 256), (1024, 256), (512, 64), (3, 576), (300, 128), (45, 16), (2000, 512), (64, 4608)])
@pytest.mark.parametrize("relu,affine", [(True, True), (False, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_bn_chain_matches_torch(rows, Cc, relu, affine, dtype):
    if dtype == torch.bfloat16 and Cc % 8:
        pytest.skip("16-bit rows are moved in 8-element chunks")
    g_ = torch.Generator(device=DEV).manual_seed(rows * 7 + Cc)
    y = (torch.randn(2, rows, Cc, device=DEV, generator=g_) * 1.5 + 0.7).to(dtype)
    gr = torch.randn(2, rows, Cc, device=DEV, generator=g_).to(dtype)
    gamma = (torch.rand(Cc, device=DEV, generator=g_) + 0.5) if affine else None
    beta = (torch.rand(Cc, device=DEV, generator=g_) * 0.6 - 0.3) if affine else None
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    a, stats = _forward(y, gamma, beta, relu, rm, rv)
    dy, dgam, dbet = _backward(gr, y, stats, relu, affine)
    # oracle: torch in fp64, view 0 then view 1 through the same module state
    bn = torch.nn.BatchNorm1d(Cc, affine=affine).to(DEV).double().train()
    if affine:
        with torch.no_grad():
            bn.weight.copy_(gamma)
            bn.bias.copy_(beta)
    y64 = y.double().requires_grad_(True)
    outs = []
    for v in range(2):
        o = bn(y64[v])
        outs.append(torch.relu(o) if relu else o)
    ref = torch.stack(outs)
    (ref * gr.double()).sum().backward()
    out_tol, grad_tol = (1e-5, 2e-5) if dtype == torch.float32 else (6e-3, 1.5e-2)
    assert _rel(a, ref) <= out_tol, _rel(a, ref)
    assert torch.allclose(rm.double(), bn.running_mean, rtol=1e-4, atol=1e-5) and torch.allclose(rv.double(), bn.running_var, rtol=1e-4, atol=1e-5)
    if rows > 8:  # with a handful of rows the input gradient is the small difference of large terms: checked through the head-stage tests
        for v in range(2):
            assert _rel(dy[v], y64.grad[v]) <= grad_tol, (v, _rel(dy[v], y64.grad[v]))
    if affine:
        assert _rel(dgam, bn.weight.grad) <= grad_tol and _rel(dbet, bn.bias.grad) <= grad_tol, (_rel(dgam, bn.weight.grad), _rel(dbet, bn.bias.grad))


@pytest.mark.parametrize("rows,Cc", [(256, 256), (100, 64), (7, 1152)])
def test_apply_normalised_rows(rows, Cc):
    g_ = torch.Generator(device=DEV).manual_seed(3)
    y = torch.randn(rows, Cc, device=DEV, generator=g_).to(torch.bfloat16)
    sc, sh = torch.rand(Cc, device=DEV, generator=g_) + 0.5, torch.randn(Cc, device=DEV, generator=g_)
    z, zh, inv = torch.empty_like(y), torch.empty_like(y), torch.empty(rows, device=DEV)
    ap = (L.HeadApplyItem * 1)(L.HeadApplyItem(y.data_ptr(), z.data_ptr(), zh.data_ptr(), inv.data_ptr(), sc.data_ptr(), sh.data_ptr(), 0, rows, Cc, 0, 0))
    L.check(L.lib().msf_head_bn_apply(ap, 1, L.MSF_BF16, 1e-8, L.stream_ptr()), "apply")
    zr = (y.double() * sc.double() + sh.double()).float().to(torch.bfloat16)
    assert torch.equal(z, zr)
    n = zr.double().norm(dim=1)
    assert torch.allclose(inv.double(), 1.0 / n, rtol=1e-5)
    assert _rel(zh, zr.double() / n[:, None]) <= 4e-3


@pytest.mark.parametrize("rows,Cc", [(512, 64), (300, 256), (8, 4608)])
def test_column_sums(rows, Cc):
    g = torch.randn(rows, Cc, device=DEV).to(torch.bfloat16)
    assert _rel(ops.column_sums(g), g.double().sum(0)) <= 1e-5
    g32 = torch.randn(rows, Cc, device=DEV)
    assert _rel(ops.column_sums(g32), g32.double().sum(0)) <= 1e-5
