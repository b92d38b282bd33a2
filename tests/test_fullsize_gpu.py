"""Parity at BASELINE.json's FULL sizes (configs 3-5: B=1024/GPU, N up to 131 072 keys), where the CPU oracle is too
slow to run directly: size-independent properties of the domain (round trips, linearity over key splits, orthogonality
of the normalise-Jacobian, idempotence, scale invariance, checksums) plus oracle spot checks on row subsets."""
import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def _nce_raw(q, k, tau=0.07, off=0):
    """forward + backward through the C ABI; returns (loss_sum, lse, grad_q fp32)."""
    nq, d = q.shape
    n = k.shape[0]
    qh, qi = ops.rownorm(q, torch.bfloat16)
    kh, _ = ops.rownorm(k, torch.bfloat16)
    wsb = L.lib().msf_infonce_workspace_bytes(nq, n, d, L.MSF_BF16)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    loss = torch.empty((), device=DEV)
    lse = torch.empty(nq, device=DEV)
    L.check(L.lib().msf_infonce_fwd(qh.data_ptr(), kh.data_ptr(), nq, n, d, off, tau, L.MSF_BF16, loss.data_ptr(), lse.data_ptr(),
                                    ws.data_ptr(), wsb, L.stream_ptr()), "fwd")
    g = torch.ones((), device=DEV)
    gq = torch.empty((nq, d), dtype=torch.float32, device=DEV)
    L.check(L.lib().msf_infonce_bwd(qh.data_ptr(), kh.data_ptr(), qi.data_ptr(), nq, n, d, off, tau, L.MSF_BF16, g.data_ptr(), 1.0 / nq,
                                    ws.data_ptr(), wsb, gq.data_ptr(), L.MSF_F32, L.stream_ptr()), "bwd")
    return loss, lse, gq, qh, kh


@pytest.mark.parametrize("nq,n,d,off", [(16384, 131072, 256, 32768), (16384, 131072, 128, 0), (4096, 32768, 64, 4096), (16384, 131072, 512, 16384),
                                         (65536, 65536, 128, 0)])
def test_infonce_full_size_properties(nq, n, d, off):
    """config 3/4 target-branch shapes (Nq = 16*B rows local, N = 8 ranks x Nq keys) and the c5 N = 65 536 case."""
    tau = 0.07
    k = torch.randn(n, d, device=DEV, generator=_gen(1))
    q = (0.3 * k[off:off + nq] + torch.randn(nq, d, device=DEV, generator=_gen(2))).to(torch.bfloat16)
    k = k.to(torch.bfloat16)
    loss, lse, gq, qh, kh = _nce_raw(q, k, tau, off)
    # (1) every row: logsumexp over all keys >= its positive logit, so row losses and their sum are >= 0
    pos = (qh.float() * kh[off:off + nq].float()).sum(1) / tau
    assert bool((lse >= pos - 1e-3).all()) and loss.item() > 0
    assert abs(loss.item() - float((lse - pos).sum())) <= 2e-3 * loss.item()
    # (2) linearity over key splits: exp-sums of disjoint key ranges add (this is what makes split-K over the keys and
    #     the rank-major all-gather exact).  lse does not depend on where the positives sit, so each half is run alone.
    half = n // 2
    rows = min(nq, half)
    lse_a = _nce_raw(q[:rows], k[:half], tau, 0)[1].double()
    lse_b = _nce_raw(q[:rows], k[half:], tau, 0)[1].double()
    assert torch.allclose(torch.logaddexp(lse_a, lse_b).float(), lse[:rows], rtol=0, atol=5e-3)
    # (3) the normalise-Jacobian makes every gradient row orthogonal to its query
    dots = (gq * q.float()).sum(1).abs()
    scale = gq.norm(dim=1) * q.float().norm(dim=1) + 1e-30
    assert float((dots / scale).max()) <= 2e-2  # bf16-rounded q_hat; exact orthogonality holds for the rounded vector
    assert float(((gq * qh.float()).sum(1).abs() / (gq.norm(dim=1) * qh.float().norm(dim=1) + 1e-30)).max()) <= 5e-3
    # (4) oracle spot check on 64 rows spread over the tiles (fp64 on the CPU against ALL keys)
    idx = torch.linspace(0, nq - 1, 64).long()
    qs, ks = q[idx.to(DEV)].double().cpu(), k.double().cpu()
    ph = qs / qs.norm(dim=1, keepdim=True)
    zh = ks / ks.norm(dim=1, keepdim=True)
    ref_lse = torch.logsumexp(ph @ zh.t() / tau, dim=1)
    assert torch.allclose(lse[idx.to(DEV)].double().cpu(), ref_lse, rtol=0, atol=2e-2)
    sm = torch.softmax(ph @ zh.t() / tau, dim=1)
    gh = (sm @ zh - zh[idx + off]) / (tau * nq)
    gref = (gh - ph * (ph * gh).sum(1, keepdim=True)) / qs.norm(dim=1, keepdim=True)
    got = gq[idx.to(DEV)].double().cpu()
    cos = float((got.flatten() @ gref.flatten()) / (got.norm() * gref.norm()))
    assert cos >= 0.9999, cos
    # (5) deterministic
    loss2 = _nce_raw(q, k, tau, off)[0]
    assert loss.item() == loss2.item()


def test_gather_concat_full_size_round_trip_and_checksums():
    """config 4: B = 1024 tiles per GPU, 4 levels x 2 views in one launch."""
    B, K, dims = 1024, 16, (64, 128, 256, 512)
    gen = torch.Generator().manual_seed(5)
    perm = torch.stack([torch.randperm(K, generator=gen) for _ in range(B)])
    rev = torch.argsort(perm, dim=1)
    ctx = [torch.randn(B, d, device=DEV, generator=_gen(10 + i)).to(torch.bfloat16) for i, d in enumerate(dims)] * 2
    raster = [torch.randn(B * K, d, device=DEV, generator=_gen(20 + i)).to(torch.bfloat16) for i, d in enumerate(dims)] * 2
    # shuffle like the dataset does (grid[perm]), then un-shuffle with the kernel: must give the raster order back
    flat_perm = (perm + torch.arange(B)[:, None] * K).flatten().to(DEV)
    shuffled = [r[flat_perm] for r in raster]
    revd = [rev.to(DEV)] * 8
    s, m = ops.gather_concat(ctx, shuffled, revd, K, 8)
    for i in range(8):
        assert torch.equal(s[i], raster[i])                                  # encode -> decode round trip, bit exact
        d = ctx[i].shape[1]
        assert torch.equal(m[i][:, :d], ctx[i])                              # concat head
        assert torch.equal(m[i][:, d:].reshape(B, 8, d), shuffled[i].reshape(B, K, d)[:, :8])  # first 8 SHUFFLED vectors
        assert torch.equal(s[i].float().sum(0), raster[i].float().sum(0))    # checksum of the multiset of rows


def test_cosine_full_size_identities():
    rows = 131072  # config 4 target branch, global
    for d, dt in ((64, torch.bfloat16), (512, torch.bfloat16), (256, torch.float32)):
        p = torch.randn(rows, d, device=DEV, generator=_gen(3)).to(dt)
        # cos(p, p) = 1 for every row -> loss = coef exactly (up to fp32 rounding of the mean)
        l_same = ops.cosine_loss([p], [p.clone()], [-0.5])
        assert abs(l_same.item() + 0.5) <= 2e-4
        # scale invariance: positive per-row rescaling of either argument leaves the loss unchanged
        z = torch.randn(rows, d, device=DEV, generator=_gen(4)).to(dt)
        l1 = ops.cosine_loss([p], [z], [-0.5])
        l2 = ops.cosine_loss([(p.float() * 4).to(dt)], [(z.float() * 0.5).to(dt)], [-0.5])  # powers of two: exact in bf16
        assert abs(l1.item() - l2.item()) <= 1e-6
        # antisymmetry
        l3 = ops.cosine_loss([p], [(-z.float()).to(dt)], [-0.5])
        assert abs(l1.item() + l3.item()) <= 1e-6


def test_ema_full_size_idempotence():
    shapes = [(4608, 4608)] * 3 + [(2304, 2304)] * 3 + [(512, 512, 3, 3)] * 4 + [(64,), (1,)]  # ~100 M parameters
    teacher = [torch.randn(*s, device=DEV, generator=_gen(i)) for i, s in enumerate(shapes)]
    student = [torch.randn(*s, device=DEV, generator=_gen(100 + i)) for i, s in enumerate(shapes)]
    t0 = [t.clone() for t in teacher]
    up = ops.EmaUpdater(teacher, student)
    up.step(1.0)  # m = 1: teacher unchanged
    assert all(torch.equal(a, b) for a, b in zip(teacher, t0))
    up.step(0.0)  # m = 0: teacher becomes the student
    assert all(torch.equal(a, b) for a, b in zip(teacher, student))
    up.step(0.5)  # fixed point: teacher == student stays put
    assert all(torch.equal(a, b) for a, b in zip(teacher, student))


def test_crop_resample_full_size_partition_of_unity():
    """Bilinear weights sum to one: a constant map stays constant; 16 integer tiles of the map reassemble it exactly."""
    B, Cc, H, W = 16, 128, 128, 128
    const = torch.full((B, Cc, H, W), 3.25, device=DEV, dtype=torch.bfloat16)
    boxes = ops.footprint_boxes(B, 4, H, W, DEV)
    out = ops.crop_resample(const, boxes, (128, 128))
    assert bool((out == 3.25).all())
    x = torch.randn(B, Cc, H, W, device=DEV, generator=_gen(9)).to(torch.bfloat16)
    tiles = ops.crop_resample(x, boxes, (32, 32))  # integer copy of the 4x4 grid (blockshaped order)
    rebuilt = tiles.reshape(B, 4, 4, Cc, 32, 32).permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, H, W)
    assert torch.equal(rebuilt, x)


@pytest.mark.timeout(600)
def test_stem_batchnorm_relu_pool_beyond_2_31_elements():
    """The stem activation of a 4096-image target batch (configs[1]: 16 tiles x batch 256) has 4096*64*112*112 =
    3.29e9 elements -- beyond 32-bit indexing.  Size-independent checks of the fused bn -> relu -> maxpool (N1) there:
    statistics against chunked fp64 sums, forward / backward of the first and last images against the unfused ATen
    sequence evaluated with the same global statistics, and sum identities of the gradient."""
    import torch.nn.functional as F
    N, C, H, W = 4096, 64, 112, 112
    g = _gen(11)
    x = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=DEV)
    for i in range(0, N, 512):  # filled chunk-wise: randn itself is limited to 2^31 elements per call
        x[i:i + 512] = (torch.randn((512, H, W, C), device=DEV, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    assert x.numel() > 2 ** 31
    xv = x.permute(0, 3, 1, 2).requires_grad_(True)  # NCHW view of the NHWC buffer
    w = torch.empty(C, device=DEV).uniform_(0.5, 1.5, generator=g).requires_grad_(True)
    b = torch.empty(C, device=DEV).uniform_(-0.5, 0.5, generator=g).requires_grad_(True)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    y = ops.bn_act2d(xv, w, b, rm, rv, 1e-5, 1.0, relu=True, pool=True)  # momentum 1: running stats = batch stats
    assert y.shape == (N, C, 56, 56)
    # (1) statistics: chunked fp64 reference
    s = torch.zeros(C, dtype=torch.float64, device=DEV)
    q = torch.zeros(C, dtype=torch.float64, device=DEV)
    for i in range(0, N, 256):
        c = x[i:i + 256].double()
        s += c.sum(dim=(0, 1, 2)); q += (c * c).sum(dim=(0, 1, 2))
    n = float(N * H * W)
    mean, var = s / n, q / n - (s / n) ** 2
    assert torch.allclose(rm.double(), mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.double(), var * n / (n - 1), rtol=1e-5, atol=1e-6)
    # (2) first / last images against ATen with the same statistics (eval-mode batch_norm = the affine map)
    gy = torch.empty_like(y)
    for i in range(0, N, 512):
        gy[i:i + 512] = torch.randn((512, C, 56, 56), device=DEV, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    for sl in (slice(0, 4), slice(N - 4, N)):
        xr = xv[sl].detach().float().requires_grad_(True)
        yr = F.max_pool2d(F.relu(F.batch_norm(xr, mean.float(), var.float(), w.detach(), b.detach(), False, 0.0, 1e-5)), 3, 2, 1)
        assert torch.allclose(y[sl].float(), yr, rtol=2 ** -7, atol=2 ** -7)
    # (3) gradient identities of batch norm: sum_rows dx = 0 and sum_rows dx * xhat = 0 per channel (up to bf16 rounding
    #     of dx), and grad_bias = sum of the pooled gradients that are alive
    dx = xv.grad
    assert dx.shape == xv.shape
    tot = torch.zeros(C, dtype=torch.float64, device=DEV)
    absum = torch.zeros(C, dtype=torch.float64, device=DEV)
    for i in range(0, N, 256):
        c = dx[i:i + 256].double()
        tot += c.sum(dim=(0, 2, 3)); absum += c.abs().sum(dim=(0, 2, 3))
    assert bool((tot.abs() <= 2e-3 * absum + 1e-6).all()), float((tot.abs() / absum).max())
    alive = torch.zeros(C, dtype=torch.float64, device=DEV)
    for i in range(0, N, 512):
        alive += (gy[i:i + 512].double() * (y[i:i + 512] > 0)).sum(dim=(0, 2, 3))
    assert torch.allclose(b.grad.double(), alive, rtol=1e-4, atol=1e-2)
