"""Grouped InfoNCE (msf_nce_grouped_fwd / _bwd: one flash launch per width class, grouped two-pass GEMMs above 256, one
finalize and one backward launch for all pairs) against the fp64 oracle (oracle/msf_oracle.py::infonce_loss / infonce_grad;
parity unpinned by the reference, which has no such loss -- anchored by logits.diag() * tau == cosine) evaluated on the
SAME bf16 inputs.  Tolerances (north_star, bf16): loss 2e-3 relative, gradient cosine >= 0.9999.
Covers: raw queries + row-norm partials (the predictor-tail epilogue form), pre-normalised queries, ragged row counts,
all width classes in one call, and rank-major key blocks of several "ranks" emulated on one GPU (3-D TMA key map,
positives in the middle block)."""
import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TAU = 0.07


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _problem(rows, dim, seed, scale=3.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    z = torch.randn(2, rows, dim, device=DEV, generator=g)
    p = (scale * (0.5 * z.flip(0) + torch.randn(2, rows, dim, device=DEV, generator=g))).to(torch.bfloat16)  # p[v] correlates with z[1-v]
    khat = torch.nn.functional.normalize(z, dim=2, eps=1e-8).to(torch.bfloat16)
    return p, khat


def _rowsq(p):
    _, rows, dim = p.shape
    blocks = (dim + 63) // 64
    pad = torch.zeros(2, rows, blocks * 64, device=DEV)
    pad[:, :, :dim] = p.float()
    return (pad * pad).view(2, rows, blocks, 64).sum(3).permute(0, 2, 1).contiguous()


def _oracle(p, khat, coef):
    """sum over the two directions of coef * InfoNCE(p[v], khat[1-v]) in fp64 on the bf16 values, with gradients."""
    loss, grads = 0.0, []
    for v in range(2):
        q = p[v].double().cpu()
        k = khat[1 - v].double().cpu()
        k = k / k.norm(dim=1, keepdim=True)  # the kernel treats the stored keys as unit vectors: compare like with like
        l, _, _ = O.infonce_loss(q, k, TAU)
        loss += coef * float(l)
        grads.append(coef * O.infonce_grad(q, k, TAU))
    return loss, torch.stack(grads)


@pytest.mark.parametrize("shapes", [[(256, 64)], [(4096, 128)], [(1000, 256)], [(300, 512)], [(96, 576)], [(256, 64), (256, 128), (4096, 128), (512, 256), (64, 512), (32, 1152)]])
@pytest.mark.parametrize("raw", [True, False])
def test_grouped_matches_oracle(shapes, raw):
    ps, khats, rowsqs, coefs = [], [], [], []
    for i, (rows, dim) in enumerate(shapes):
        p, khat = _problem(rows, dim, 10 + i)
        if not raw:
            p = torch.nn.functional.normalize(p.float(), dim=2).to(torch.bfloat16)
        ps.append(p.requires_grad_(True))
        khats.append(khat)
        rowsqs.append(_rowsq(p.detach()))
        coefs.append(0.5 * (0.1 + 0.3 * i))
    extras = {"khat": khats, "rowsq": rowsqs if raw else None, "kflat": None, "kgathered": None, "kready": None, "world": 1, "rank": 0}
    before = L.launch_count
    loss = ops.infonce_grouped(ps, extras, coefs, TAU)
    (loss * 8.0).backward()
    want, wgrads = 0.0, []
    for p, khat, c in zip(ps, khats, coefs):
        l, g = _oracle(p.detach(), khat, c)
        want += l
        wgrads.append(g)
    assert abs(float(loss) - want) <= 2e-3 * abs(want), (float(loss), want)
    for p, g in zip(ps, wgrads):
        assert p.grad is not None and p.grad.dtype == torch.bfloat16
        assert _cos(p.grad.cpu() / 8.0, g) >= 0.9999
        assert abs(float(p.grad.double().norm().cpu()) / 8.0 - float(g.norm())) <= 1e-2 * float(g.norm())


def test_rank_major_key_blocks_emulated():
    """world = 3 key blocks laid out rank-major in one buffer (what the single all-gather produces), positives in block 1."""
    rows, world, rank = 200, 3, 1
    for dim in (128, 576):
        g = torch.Generator(device=DEV).manual_seed(7 + dim)
        zall = torch.randn(world, 2, rows, dim, device=DEV, generator=g)
        kall = torch.nn.functional.normalize(zall, dim=3).to(torch.bfloat16).contiguous()  # [rank][view][rows][dim]: per-rank flat key buffers
        p = (2.0 * (0.5 * zall[rank].flip(0) + torch.randn(2, rows, dim, device=DEV, generator=g))).to(torch.bfloat16).contiguous()
        rsq = _rowsq(p)
        stride = 2 * rows * dim  # elements per rank block of the gathered buffer
        grad = torch.empty_like(p)
        n = 2
        pairs = (L.NcePair * n)()
        for v in range(2):
            kptr = kall[0, 1 - v].data_ptr()  # rank 0's copy of the keys of view 1-v
            pairs[v] = L.NcePair(p[v].data_ptr(), rsq[v].data_ptr(), kptr, grad[v].data_ptr(), stride, rows, rows, world, dim, rank, 0.5)
        wsb = L.lib().msf_nce_grouped_workspace_bytes(pairs, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
        loss, gout = torch.empty((), device=DEV), torch.ones((), device=DEV)
        L.check(L.lib().msf_nce_grouped_fwd(pairs, n, L.MSF_BF16, TAU, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr()), "fwd")
        L.check(L.lib().msf_nce_grouped_bwd(pairs, n, L.MSF_BF16, TAU, 1e-8, gout.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr()), "bwd")
        want = 0.0
        for v in range(2):
            q = p[v].double().cpu()
            keys = kall[:, 1 - v].reshape(world * rows, dim).double().cpu()
            keys = keys / keys.norm(dim=1, keepdim=True)
            l, _, _ = O.infonce_loss(q, keys, TAU, pos_offset=rank * rows)
            want += 0.5 * float(l)
            gr = 0.5 * O.infonce_grad(q, keys, TAU, rank * rows)
            assert _cos(grad[v].cpu(), gr) >= 0.9999, (dim, v)
        assert abs(float(loss) - want) <= 2e-3 * abs(want), (dim, float(loss), want)


def test_argument_validation():
    p, khat = _problem(64, 64, 1)
    ex = {"khat": [khat], "rowsq": None, "kflat": None, "kgathered": None, "kready": None, "world": 1, "rank": 0}
    with pytest.raises(ValueError):
        ops.infonce_grouped([p.float()], ex, [1.0], TAU)
    with pytest.raises(RuntimeError, match="tau"):
        ops.infonce_grouped([p], ex, [1.0], 0.001)
