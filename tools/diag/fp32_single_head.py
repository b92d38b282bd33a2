"""Diagnostic: single-head fp32 head_stage vs torch fp64 for several (rows, d); per-view dx error."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import msfwsi_b200 as M
from msfwsi_b200 import heads as H
from oracle import torch_ref as R

DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def one(rows, d, seed=0, nheads=1):
    torch.manual_seed(seed)
    refs, tmods = [], []
    for _ in range(nheads):
        pj, pd = M.make_projector(d, d).to(DEV), M.make_predictor(d, d // 4).to(DEV)
        with torch.no_grad():
            for mod in (pj, pd):
                for n, p in mod.named_parameters():
                    if p.dim() == 1 and n.endswith("weight"):
                        p.uniform_(0.5, 1.5)
                    elif p.dim() == 1:
                        p.uniform_(-0.3, 0.3)
        tpj, tpd = R.make_projector(d, d).to(DEV).double(), R.make_predictor(d, d // 4).to(DEV).double()
        tpj.load_state_dict(pj.state_dict()); tpd.load_state_dict(pd.state_dict())
        refs.append(H.HeadRefs(pj, pd)); tmods.append((tpj, tpd, pj, pd))
    xs = [torch.randn(2, rows, d, device=DEV).abs().requires_grad_(True) for _ in range(nheads)]
    p, z, _ = H.head_stage(xs, refs, True, None, dtype=torch.float32)
    w = [torch.randn(2, rows, d, device=DEV) for _ in range(nheads)]
    sum((pp * ww).sum() for pp, ww in zip(p, w)).backward()
    for hi, (tpj, tpd, pj, pd) in enumerate(tmods):
        x64 = xs[hi].detach().double().requires_grad_(True)
        outs = []
        for v in range(2):
            zz = tpj(x64[v]); outs.append(tpd(zz))
        (torch.stack(outs) * w[hi].double()).sum().backward()
        line = f"rows {rows} d {d} head {hi}: p {rel(p[hi], torch.stack(outs)):.1e} dx v0 {rel(xs[hi].grad[0], x64.grad[0]):.1e} v1 {rel(xs[hi].grad[1], x64.grad[1]):.1e}"
        worst = max((rel(a.grad, b.grad), n) for (n, a), (_, b) in zip(list(pj.named_parameters()) + list(pd.named_parameters()),
                                                                          list(tpj.named_parameters()) + list(tpd.named_parameters())))
        print(line, "worst param", worst[1], f"{worst[0]:.1e}", flush=True)
        if worst[0] > 1e-4:
            for (n, a), (_, b) in zip([("proj." + k, v) for k, v in pj.named_parameters()] + [("pred." + k, v) for k, v in pd.named_parameters()],
                                      list(tpj.named_parameters()) + list(tpd.named_parameters())):
                print(f"    {n}: {rel(a.grad, b.grad):.2e}")


for rows, d in ((256, 256), (1024, 256)):
    one(rows, d)
