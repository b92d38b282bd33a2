"""GPU parity tests of the InfoNCE KEY gradient (BASELINE.json:north_star (4): "an NCCL all-gather ... supplies global
negatives before the fused loss, and its gradient is reduce-scattered").  The reference detaches every key
(src/models/backbone.py:188-191), so this is the non-detached variant of the extension: parity unpinned, oracle =
oracle/msf_oracle.py:infonce_key_grad (pinned to torch autograd in tests/test_oracle_golden.py).
Through the C ABI: msf_infonce_dk (the flash pass with the roles swapped) + msf_infonce_dk_finish.
Bars: gradient cosine >= 0.9999; fp32 relative error <= 1e-4."""
import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _relerr(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _inputs(nq, n, dim, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(nq, dim, generator=g)
    k = torch.randn(n, dim, generator=g)
    k[:nq] = q + 2.0 * torch.randn(nq, dim, generator=g)  # positives correlate (cos ~ 0.45): the softmax is far from one-hot
    return q.to(dtype), k.to(dtype)


@pytest.mark.parametrize("nq,dim", [(200, 128), (70, 64), (64, 40), (333, 256)])
def test_key_gradient_fp32_autograd_entry(nq, dim):
    tau = 0.07
    q, k = _inputs(nq, nq, dim, 51)
    qd, kd = q.to(DEV).requires_grad_(True), k.to(DEV).requires_grad_(True)
    loss = ops.infonce_loss(qd, kd, tau=tau, precision=torch.float32, detach_keys=False)
    (2.0 * loss).backward()
    ref, _, _ = O.infonce_loss(q.double(), k.double(), tau)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(abs(ref.item()), 0.5)  # logits are O(1/tau): fp32 rounding floor for tiny losses
    gq = 2.0 * O.infonce_grad(q.double(), k.double(), tau)
    gk = 2.0 * O.infonce_key_grad(q.double(), k.double(), tau)
    assert _cos(qd.grad, gq) >= 0.9999 and _relerr(qd.grad, gq) <= 1e-4
    assert kd.grad is not None and kd.grad.dtype == torch.float32
    assert _cos(kd.grad, gk) >= 0.9999, _cos(kd.grad, gk)
    assert _relerr(kd.grad, gk) <= 1e-4, _relerr(kd.grad, gk)
    # torch expression on the GPU (normalize -> matmul -> cross_entropy), gradient through both operands
    q2, k2 = q.to(DEV).requires_grad_(True), k.to(DEV).requires_grad_(True)
    logits = torch.nn.functional.normalize(q2, dim=1, eps=1e-8) @ torch.nn.functional.normalize(k2, dim=1, eps=1e-8).t() / tau
    (2.0 * torch.nn.functional.cross_entropy(logits, torch.arange(nq, device=DEV))).backward()
    assert _cos(kd.grad, k2.grad) >= 0.9999


def test_keys_stay_detached_by_default():
    q, k = _inputs(64, 64, 64, 52)
    qd, kd = q.to(DEV).requires_grad_(True), k.to(DEV).requires_grad_(True)
    ops.infonce_loss(qd, kd, tau=0.07, precision=torch.float32).backward()  # reference semantics, backbone.py:188-191
    assert qd.grad is not None and kd.grad is None


@pytest.mark.parametrize("nq,dim", [(128, 64), (128, 128), (128, 256), (300, 128), (1000, 256), (4096, 128), (2500, 64), (5, 128)])
def test_key_gradient_bf16_flash(nq, dim):
    """bf16 operands -> the transposed TMA / tcgen05 / TMEM pass (infonce_grouped_kernel<D, COLB = true>)."""
    tau = 0.07
    q, k = _inputs(nq, nq, dim, 53, torch.bfloat16)
    qd, kd = q.to(DEV).requires_grad_(True), k.to(DEV).requires_grad_(True)
    loss = ops.infonce_loss(qd, kd, tau=tau, detach_keys=False)
    loss.backward()
    torch.cuda.synchronize()
    ref, _, _ = O.infonce_loss(q.double(), k.double(), tau)
    assert abs(loss.item() - ref.item()) <= 2e-3 * max(abs(ref.item()), 0.5)
    assert kd.grad.dtype == torch.bfloat16 and torch.isfinite(kd.grad).all()
    # 5 keys: the softmax terms nearly cancel and the bf16 rounding of the stored gradient itself is visible
    bar = 0.9999 if nq >= 64 else 0.999
    assert _cos(qd.grad, O.infonce_grad(q.double(), k.double(), tau)) >= bar
    c = _cos(kd.grad, O.infonce_key_grad(q.double(), k.double(), tau))
    assert c >= bar, c


@pytest.mark.parametrize("nq,dim", [(256, 512), (200, 576)])
def test_key_gradient_bf16_two_pass_widths(nq, dim):
    """Widths above 256: dK = P^T (Q_hat / sum) with the forward's 16-bit P still in its workspace (one tcgen05 GEMM)."""
    tau = 0.07
    q, k = _inputs(nq, nq, dim, 54, torch.bfloat16)
    qd, kd = q.to(DEV).requires_grad_(True), k.to(DEV).requires_grad_(True)
    ops.infonce_loss(qd, kd, tau=tau, detach_keys=False).backward()
    c = _cos(kd.grad, O.infonce_key_grad(q.double(), k.double(), tau))
    assert c >= 0.9999, c


@pytest.mark.parametrize("precision,dim", [(torch.float32, 64), (torch.bfloat16, 128), (torch.bfloat16, 256)])
def test_key_gradient_sharded_queries_emulated_reduce_scatter(precision, dim):
    """Two 'ranks' on one GPU through the bare C ABI: each owns half of the queries and keys, sees ALL keys (rank-major), and
    computes its partial over every key; the sum of the partials restricted to a rank's own rows (= what
    dist.reduce_scatter_tensor delivers) + msf_infonce_dk_finish must equal the single-process gradient x world."""
    tau, rows, world = 0.07, 384, 2
    n = rows * world
    q, k = _inputs(n, n, dim, 55, precision)
    prec = L.dtype_code(precision)
    k_hat, k_inv = ops.rownorm(k.to(DEV), precision)
    g = torch.full((), 1.0, device=DEV)
    parts, q_hats = [], []
    for r in range(world):
        q_hat, _ = ops.rownorm(q[r * rows:(r + 1) * rows].to(DEV), precision)
        off = r * rows
        ws_bytes = L.lib().msf_infonce_workspace_bytes(rows, n, dim, prec)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
        loss_sum = torch.empty((), dtype=torch.float32, device=DEV)
        L.check(L.lib().msf_infonce_fwd(q_hat.data_ptr(), k_hat.data_ptr(), rows, n, dim, off, tau, prec, loss_sum.data_ptr(), 0,
                                        ws.data_ptr(), ws_bytes, L.stream_ptr()), "fwd")
        dws_bytes = L.lib().msf_infonce_dk_workspace_bytes(rows, n, dim, prec)
        dws = torch.empty(dws_bytes, dtype=torch.uint8, device=DEV)
        dk = torch.empty((n, dim), dtype=torch.float32, device=DEV)
        L.check(L.lib().msf_infonce_dk(q_hat.data_ptr(), k_hat.data_ptr(), rows, n, dim, off, tau, prec, g.data_ptr(), 1.0 / rows,
                                       ws.data_ptr(), ws_bytes, dk.data_ptr(), dws.data_ptr(), dws_bytes, L.stream_ptr()), "dk")
        parts.append(dk)
        q_hats.append(q_hat)
    total = parts[0] + parts[1]
    single = O.infonce_key_grad(q.double(), k.double(), tau)  # mean over all n queries
    for r in range(world):
        mine = total[r * rows:(r + 1) * rows].contiguous()
        gz = torch.empty((rows, dim), dtype=torch.float32, device=DEV)
        L.check(L.lib().msf_infonce_dk_finish(mine.data_ptr(), q_hats[r].data_ptr(), k_hat[r * rows:(r + 1) * rows].data_ptr(),
                                              k_inv[r * rows:(r + 1) * rows].data_ptr(), rows, rows, dim, tau, prec, g.data_ptr(), 1.0 / rows,
                                              gz.data_ptr(), L.MSF_F32, L.stream_ptr()), "finish")
        ref = world * single[r * rows:(r + 1) * rows]
        assert _cos(gz, ref) >= 0.9999, (r, _cos(gz, ref))
        if precision == torch.float32:
            assert _relerr(gz, ref) <= 1e-4


def test_key_gradient_rejects_bad_workspace():
    q_hat, _ = ops.rownorm(torch.randn(64, 64, device=DEV), torch.float32)
    g = torch.ones((), device=DEV)
    dk = torch.empty((64, 64), device=DEV)
    ws = torch.empty(16, dtype=torch.uint8, device=DEV)
    rc = L.lib().msf_infonce_dk(q_hat.data_ptr(), q_hat.data_ptr(), 64, 64, 64, 0, 0.07, L.MSF_F32, g.data_ptr(), 1.0, ws.data_ptr(), 16,
                                dk.data_ptr(), ws.data_ptr(), 16, L.stream_ptr())
    assert rc != 0 and b"workspace" in L.lib().msf_last_error()
