"""Diagnose the tcgen05 InfoNCE main loop: compare the row-sum and O partials in the workspace with a torch
computation on the same bf16 operands.  Development tool (GPU box only)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L  # noqa: E402

dev = "cuda:0"


def run(qh, kh, tau=0.07):
    nq, d = qh.shape
    n = kh.shape[0]
    info = (C.c_int64 * 8)()
    L.check(L.lib().msf_infonce_plan_info(nq, n, d, L.MSF_BF16, info), "plan")
    splits, nq_pad, off_rs, off_o = info[0], info[1], info[2], info[3]
    wsb = L.lib().msf_infonce_workspace_bytes(nq, n, d, L.MSF_BF16)
    ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
    loss = torch.zeros((), device=dev)
    L.check(L.lib().msf_infonce_fwd(qh.data_ptr(), kh.data_ptr(), nq, n, d, 0, tau, L.MSF_BF16, loss.data_ptr(), 0, ws.data_ptr(), wsb,
                                    L.stream_ptr()), "fwd")
    torch.cuda.synchronize()
    rs = ws[off_rs:off_rs + splits * nq_pad * 4].view(torch.float32).view(splits, nq_pad).sum(0)[:nq]
    o = ws[off_o:off_o + splits * nq_pad * d * 4].view(torch.float32).view(splits, nq_pad, d).sum(0)[:nq]
    return rs, o, splits


def ref(qh, kh, tau=0.07):
    a = 1.4426950408889634 / tau
    s = qh.double() @ kh.double().t()
    p = torch.exp2(a * s - a)
    pb = p.float().to(torch.bfloat16).double()
    return p.sum(1), pb @ kh.double()


def report(name, qh, kh):
    rs, o, splits = run(qh, kh)
    rrs, ro = ref(qh, kh)
    e_rs = ((rs.double() - rrs).abs() / rrs.abs()).max().item()
    e_o = ((o.double() - ro).norm() / ro.norm()).item()
    print(f"{name}: splits={splits} rowsum max rel err {e_rs:.3e} | O rel err {e_o:.3e}")
    if e_o > 1e-2:
        d = qh.shape[1]
        err = (o.double() - ro).abs()
        print("   per-64-col block err:", [f"{err[:, c:c + 64].mean().item():.2e}" for c in range(0, d, 64)],
              " ref mag:", f"{ro.abs().mean().item():.2e}")
        print("   row 0 first 8 got:", o[0, :8].tolist())
        print("   row 0 first 8 ref:", ro[0, :8].tolist())
        # does a column permutation explain it?
        best = []
        for c in range(min(d, 16)):
            corr = [(torch.dot(o[:, c].double(), ro[:, j]) / (o[:, c].double().norm() * ro[:, j].norm() + 1e-30)).item() for j in range(d)]
            j = max(range(d), key=lambda t: corr[t])
            best.append((c, j, round(corr[j], 3)))
        print("   best matching ref column for got columns 0..15:", best)


def main():
    torch.manual_seed(0)
    for d in (64, 128, 256):
        for nq, n in ((128, 128), (128, 512), (256, 2048)):
            # 1) Q = 0 -> P constant; K column pattern only: isolates the GEMM2 B operand (MN-major) d-mapping
            kh = (torch.arange(d, device=dev).float() / d).repeat(n, 1).to(torch.bfloat16)
            qh = torch.zeros(nq, d, device=dev, dtype=torch.bfloat16)
            report(f"[d={d} nq={nq} n={n}] Q=0, K=f(col)", qh, kh)
            # 2) Q = 0, K key-dependent: isolates key (K-dim) mapping in GEMM2
            kh = ((torch.arange(n, device=dev).float() % 7) / 7).unsqueeze(1).repeat(1, d).to(torch.bfloat16) * \
                 (1 + (torch.arange(d, device=dev).float() % 3)).unsqueeze(0).to(torch.bfloat16)
            report(f"[d={d} nq={nq} n={n}] Q=0, K=f(key)*g(col)", qh, kh)
            # 3) random normalised
            k = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=1)
            q = torch.nn.functional.normalize(0.3 * k[:nq] + torch.nn.functional.normalize(torch.randn(nq, d, device=dev), dim=1), dim=1)
            report(f"[d={d} nq={nq} n={n}] random", q.to(torch.bfloat16), k.to(torch.bfloat16))


if __name__ == "__main__":
    main()
