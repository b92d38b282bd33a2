import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import msfwsi_b200 as M
from msfwsi_b200 import heads as H
DEV = "cuda:0"
rows, d = int(sys.argv[1]), int(sys.argv[2])
torch.manual_seed(0)
pj, pd = M.make_projector(d, d).to(DEV), M.make_predictor(d, d // 4).to(DEV)
with torch.no_grad():
    for mod in (pj, pd):
        for n, p in mod.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif p.dim() == 1:
                p.uniform_(-0.3, 0.3)
refs = [H.HeadRefs(pj, pd)]
x = torch.randn(2, rows, d, device=DEV).abs()
w = torch.randn(2, rows, d, device=DEV)
res = []
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    xs = x.clone().requires_grad_(True)
    for m in (pj, pd):
        m.zero_grad(set_to_none=True)
    p, z, _ = H.head_stage([xs], refs, True, None, dtype=torch.float32)
    (p[0] * w).sum().backward()
    torch.cuda.synchronize()
    res.append([xs.grad.clone()] + [q.grad.clone() for q in list(pj.parameters()) + list(pd.parameters())])
names = ["dx"] + [n for n, _ in list(pj.named_parameters())] + ["pred." + n for n, _ in list(pd.named_parameters())]
for it in range(1, len(res)):
    diffs = [(float((a - b).abs().max()), n) for a, b, n in zip(res[0], res[it], names) if not torch.equal(a, b)]
    print("run", it, "differs from run 0 in", diffs)
