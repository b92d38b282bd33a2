"""GPU parity of the channels-last BatchNorm2d kernels (N1: bn [+residual] [+relu] [+stem max-pool], forward and
backward, called through the C ABI) against the plain PyTorch sequence the reference's ResNet runs
(src/models/resnet.py:59-82, 244-247): F.batch_norm -> (+identity) -> relu -> max_pool2d.  This is a floating-point
kernel: the oracle is torch fp32 / fp64 on the same inputs.  Tolerances: fp32 1e-5 relative (outputs, running stats,
gradients); bf16 one output ulp (2^-8 relative) and gradient cosine >= 0.9999."""
import pytest
import torch
import torch.nn.functional as F

from msfwsi_b200 import ops
from msfwsi_b200 import resnet as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _ref(x, w, b, res, relu, pool, eps=1e-5, dtype=torch.float64):
    """The unfused sequence in fp64 on the (already rounded) inputs."""
    x = x.detach().to(dtype).requires_grad_(True)
    w = w.detach().to(dtype).requires_grad_(True)
    b = b.detach().to(dtype).requires_grad_(True)
    r = None if res is None else res.detach().to(dtype).requires_grad_(True)
    rm, rv = torch.zeros(x.shape[1], dtype=dtype, device=x.device), torch.ones(x.shape[1], dtype=dtype, device=x.device)
    y = F.batch_norm(x, rm, rv, w, b, True, 0.1, eps)
    if r is not None:
        y = y + r
    if relu:
        y = F.relu(y)
    if pool:
        y = F.max_pool2d(y, 3, 2, 1)
    return x, w, b, r, y, rm, rv


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,relu,res", [((6, 64, 14, 14), True, False), ((5, 128, 7, 9), False, False), ((4, 256, 6, 6), True, True),
                                            ((3, 512, 4, 4), False, True), ((7, 24, 5, 3), True, True), ((130, 64, 28, 28), True, False)])
def test_bn_act_forward_backward_vs_torch(dtype, shape, relu, res):
    if dtype != torch.float32 and shape[1] % 8:
        pytest.skip("16-bit rows need C % 8 == 0")
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = (torch.randn(shape, device=DEV, generator=g) * 1.7 + 0.4).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    r = (torch.randn(shape, device=DEV, generator=g)).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True) if res else None
    w = torch.empty(shape[1], device=DEV).uniform_(-1.5, 1.5, generator=g).requires_grad_(True)  # negative gammas too
    b = torch.empty(shape[1], device=DEV).uniform_(-0.5, 0.5, generator=g).requires_grad_(True)
    rm, rv = torch.zeros(shape[1], device=DEV), torch.ones(shape[1], device=DEV)
    y = ops.bn_act2d(x, w, b, rm, rv, 1e-5, 0.1, relu=relu, residual=r)
    assert y.dtype == dtype and y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr, wr, br, rr, yr, rmr, rvr = _ref(x, w, b, r, relu, False)
    yr.backward(gy.double())
    out_tol = {torch.float32: 2e-5, torch.bfloat16: 2 ** -7, torch.float16: 2 ** -10}[dtype]
    assert torch.allclose(y.double(), yr, rtol=out_tol, atol=out_tol), float((y.double() - yr).abs().max())
    assert torch.allclose(rm.double(), rmr, rtol=1e-5, atol=1e-6) and torch.allclose(rv.double(), rvr, rtol=1e-5, atol=1e-6)
    if dtype == torch.float32:
        # the ReLU mask of an element within rounding distance of 0 may differ from fp64: compare where it is clear-cut
        assert (x.grad.double() - xr.grad).norm() / xr.grad.norm() <= 1e-4
        assert (w.grad.double() - wr.grad).norm() / wr.grad.norm() <= 1e-4
    assert _cos(x.grad, xr.grad) >= 0.9999 and _cos(w.grad, wr.grad) >= 0.9999 and _cos(b.grad, br.grad) >= 0.9999
    if res:
        assert _cos(r.grad, rr.grad) >= 0.9999


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 64, 16, 16), (2, 64, 15, 13), (5, 8, 9, 10), (2, 128, 2, 2), (33, 64, 56, 56)])
def test_stem_bn_relu_pool_vs_torch(dtype, shape):
    g = torch.Generator(device=DEV).manual_seed(sum(shape) + 1)
    x = (torch.randn(shape, device=DEV, generator=g) * 2 - 0.3).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = torch.empty(shape[1], device=DEV).uniform_(-1.5, 1.5, generator=g).requires_grad_(True)
    b = torch.empty(shape[1], device=DEV).uniform_(-0.5, 0.5, generator=g).requires_grad_(True)
    rm, rv = torch.zeros(shape[1], device=DEV), torch.ones(shape[1], device=DEV)
    y = ops.bn_act2d(x, w, b, rm, rv, 1e-5, 0.1, relu=True, pool=True)
    N, C, H, W = shape
    assert y.shape == (N, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1) and y.is_contiguous(memory_format=torch.channels_last)
    gy = torch.randn(y.shape, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    xr, wr, br, _, yr, rmr, rvr = _ref(x, w, b, None, True, True)
    yr.backward(gy.double())
    out_tol = 2e-5 if dtype == torch.float32 else 2 ** -7
    assert torch.allclose(y.double(), yr, rtol=out_tol, atol=out_tol)
    assert torch.allclose(rv.double(), rvr, rtol=1e-5, atol=1e-6)
    if dtype == torch.float32:
        assert (x.grad.double() - xr.grad).norm() / xr.grad.norm() <= 1e-4
    # bf16: ties between equal rounded values inside a window go to the first tap (ATen's rule); fp64 has no ties, so
    # a few gradients land on a different-but-equal-valued position
    assert _cos(x.grad, xr.grad) >= (0.9999 if dtype == torch.float32 else 0.995)
    assert _cos(w.grad, wr.grad) >= 0.9999 and _cos(b.grad, br.grad) >= 0.9999


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(5, 64, 7, 7), (3, 128, 14, 14), (40, 64, 56, 56)])
def test_block_output_with_folded_average_pool_gradient(dtype, shape):
    """relu(bn(x) + identity) with its global average pool as a second output: the pooled branch's gradient is folded
    into the batch-norm backward kernels (src/models/resnet.py:250-254 pools every layer output)."""
    g = torch.Generator(device=DEV).manual_seed(sum(shape) + 7)
    mk = lambda: torch.randn(shape, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    x, r = mk(), mk()
    w = torch.empty(shape[1], device=DEV).uniform_(0.5, 1.5, generator=g).requires_grad_(True)
    b = torch.empty(shape[1], device=DEV).uniform_(-0.5, 0.5, generator=g).requires_grad_(True)
    y, m = ops.bn_act2d(x, w, b, None, None, 1e-5, 0.1, relu=True, residual=r, want_mean=True)
    assert m.shape == shape[:2]
    gy = torch.randn(shape, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    gm = (torch.randn(shape[:2], device=DEV, generator=g) * shape[2] * shape[3] ** 0.5).to(dtype)  # comparable magnitude per pixel
    torch.autograd.backward([y, m], [gy, gm])
    xr, wr, br, rr, yr, _, _ = _ref(x, w, b, r, True, False)
    mr = yr.mean(dim=(2, 3))
    torch.autograd.backward([yr, mr], [gy.double(), gm.double()])
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7
    assert torch.allclose(m.double(), mr, rtol=tol, atol=tol)
    if dtype == torch.float32:
        assert (x.grad.double() - xr.grad).norm() / xr.grad.norm() <= 1e-4
        assert (r.grad.double() - rr.grad).norm() / rr.grad.norm() <= 1e-4
    assert _cos(x.grad, xr.grad) >= 0.9999 and _cos(r.grad, rr.grad) >= 0.9999
    assert _cos(w.grad, wr.grad) >= 0.9999 and _cos(b.grad, br.grad) >= 0.9999
    # pooled branch alone (no gradient on the feature map)
    x2, r2 = x.detach().clone().requires_grad_(True), r.detach().clone().requires_grad_(True)
    _, m2 = ops.bn_act2d(x2, w, b, None, None, 1e-5, 0.1, relu=True, residual=r2, want_mean=True)
    m2.backward(gm)
    xr2, _, _, rr2, yr2, _, _ = _ref(x, w, b, r, True, False)
    yr2.mean(dim=(2, 3)).backward(gm.double())
    assert _cos(x2.grad, xr2.grad) >= 0.9999 and _cos(r2.grad, rr2.grad) >= 0.9999


def test_stem_pool_tie_rule_matches_aten_bf16():
    """Same dtype, same unfused op order on ATen: with bf16 storage the arg-max ties must resolve like max_pool2d
    (first maximum in scan order), so the input gradients agree to bf16 rounding."""
    g = torch.Generator(device=DEV).manual_seed(11)
    shape = (4, 64, 24, 24)
    x = (torch.randn(shape, device=DEV, generator=g) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x = (x * 8).round() / 8  # coarse grid: many exact ties inside every window
    x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w = torch.ones(64, device=DEV)
    b = torch.zeros(64, device=DEV)
    y1 = ops.bn_act2d(x1, w, b, None, None, 1e-5, 0.1, relu=True, pool=True)
    y2 = F.max_pool2d(F.relu(F.batch_norm(x2, None, None, w, b, True, 0.1, 1e-5)), 3, 2, 1)
    gy = torch.randn(y2.shape, device=DEV, generator=g).to(torch.bfloat16)
    y1.backward(gy)
    y2.backward(gy)
    assert (y1.float() - y2.float()).abs().max() <= 2 ** -7 * y2.float().abs().max()
    assert _cos(x1.grad, x2.grad) >= 0.999


def test_fused_module_matches_batchnorm2d_and_state_dict_keys():
    torch.manual_seed(0)
    x = torch.randn(12, 64, 20, 20, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wgt = torch.randn_like(x, dtype=torch.float32)
    ref = torch.nn.BatchNorm2d(64).to(DEV).train()
    mine = R.FusedBatchNorm2d(64).to(DEV).train()
    assert set(mine.state_dict()) == set(ref.state_dict())
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.uniform_(-0.5, 0.5)
    mine.load_state_dict(ref.state_dict())
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, ym = ref(xr), mine(xm)
    (yr.float() * wgt).sum().backward()
    (ym.float() * wgt).sum().backward()
    assert torch.allclose(ym.float(), yr.float(), rtol=2e-2, atol=2e-2)
    assert torch.allclose(mine.running_mean, ref.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(mine.running_var, ref.running_var, rtol=1e-4, atol=1e-5)
    assert int(mine.num_batches_tracked) == 1
    assert _cos(xm.grad, xr.grad) >= 0.9999
    assert _cos(mine.weight.grad, ref.weight.grad) >= 0.9999 and _cos(mine.bias.grad, ref.bias.grad) >= 0.9999
    # eval mode: running statistics, plain ATen path
    mine.eval(); ref.eval()
    assert torch.allclose(mine(x).float(), ref(x).float(), rtol=2e-2, atol=2e-2)


def test_resnet18_encoder_matches_plain_torch_resnet():
    """Whole encoder (convs on cuDNN + the fused BN kernels) against the same weights run through plain
    nn.BatchNorm2d / relu / max_pool2d modules in fp32."""
    import copy
    torch.manual_seed(1)
    torch.backends.cudnn.allow_tf32 = False  # fp32 convolutions on both sides
    enc = R.resnet18(return_features=True, zero_init_residual=False).to(DEV).to(memory_format=torch.channels_last).train()
    enc.fc = torch.nn.Identity()

    class PlainBN(torch.nn.Module):  # the unfused sequence with the same parameters
        def __init__(self, f):
            super().__init__()
            self.bn = torch.nn.BatchNorm2d(f.num_features, f.eps, f.momentum).to(DEV)
            self.bn.load_state_dict(f.state_dict())
            self.act = f.act

        def forward(self, x, residual=None, want_mean=False):
            y = self.bn(x)
            if residual is not None:
                y = y + residual
            if self.act != "none":
                y = F.relu(y)
            y = F.max_pool2d(y, 3, 2, 1) if self.act == "relu_pool" else y
            return (y, y.mean(dim=(2, 3))) if want_mean else y

    ref = copy.deepcopy(enc)
    for name, m in list(ref.named_modules()):
        for cname, c in list(m.named_children()):
            if isinstance(c, R.FusedBatchNorm2d):
                setattr(m, cname, PlainBN(c))
    x = torch.randn(6, 3, 64, 64, device=DEV).contiguous(memory_format=torch.channels_last)
    fa, fb = enc(x), ref(x)
    for a, b_ in zip(fa, fb):
        assert float((a - b_).norm() / b_.norm()) <= 2e-3, float((a - b_).norm() / b_.norm())
    sum(f.square().mean() for f in fa).backward()
    sum(f.square().mean() for f in fb).backward()
    ga = {n: p.grad for n, p in enc.named_parameters()}
    for n, p in ref.named_parameters():
        key = n.replace(".bn.", ".")
        assert _cos(ga[key], p.grad) >= 0.999, (key, _cos(ga[key], p.grad))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,dim,affine,act", [(48, 64, True, "relu"), (256, 576, True, "relu"), (4096, 128, False, "none"),
                                                 (37, 4608, False, "none"), (512, 16, True, "relu")])
def test_fused_batchnorm1d_matches_torch(dtype, rows, dim, affine, act):
    """Heads' BatchNorm1d (+ReLU) on the [rows][C] kernels vs nn.BatchNorm1d (+F.relu) in fp64 (backbone.py:15-21,28)."""
    from msfwsi_b200.module import FusedBatchNorm1d
    g = torch.Generator(device=DEV).manual_seed(rows + dim)
    x = (torch.randn(rows, dim, device=DEV, generator=g) * 1.3 + 0.2).to(dtype).requires_grad_(True)
    mine = FusedBatchNorm1d(dim, affine=affine, act=act).to(DEV).train()
    ref = torch.nn.BatchNorm1d(dim, affine=affine).to(DEV).double().train()
    assert set(mine.state_dict()) == set(ref.state_dict())
    if affine:
        with torch.no_grad():
            mine.weight.uniform_(-1.5, 1.5, generator=g); mine.bias.uniform_(-0.5, 0.5, generator=g)
            ref.weight.copy_(mine.weight); ref.bias.copy_(mine.bias)
    y = mine(x)
    assert y.shape == (rows, dim) and y.dtype == dtype
    xr = x.detach().double().requires_grad_(True)
    yr = ref(xr)
    if act == "relu":
        yr = F.relu(yr)
    gy = torch.randn(rows, dim, device=DEV, generator=g).to(dtype)
    y.backward(gy)
    yr.backward(gy.double())
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7
    assert torch.allclose(y.double(), yr, rtol=tol, atol=tol)
    assert torch.allclose(mine.running_mean.double(), ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(mine.running_var.double(), ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(mine.num_batches_tracked) == 1
    if dtype == torch.float32:
        assert (x.grad.double() - xr.grad).norm() / xr.grad.norm() <= 1e-4
    assert _cos(x.grad, xr.grad) >= 0.9999
    if affine:
        assert _cos(mine.weight.grad, ref.weight.grad) >= 0.9999 and _cos(mine.bias.grad, ref.bias.grad) >= 0.9999
