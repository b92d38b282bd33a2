import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import msfwsi_b200 as M
from msfwsi_b200 import heads as H
from oracle import torch_ref as R
DEV = "cuda:0"
rows, d = 256, 256
torch.manual_seed(0)
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
xs = torch.randn(2, rows, d, device=DEV).abs().requires_grad_(True)
p, z, _ = H.head_stage([xs], [H.HeadRefs(pj, pd)], True, None, dtype=torch.float32)
w = torch.randn(2, rows, d, device=DEV)
(p[0] * w).sum().backward()
x64 = xs.detach().double().requires_grad_(True)
import torch.nn.functional as F
W1, g1, b1 = tpj[0].weight, tpj[1].weight, tpj[1].bias
bns, avs, outs = [], [], []
for v in range(2):
    y1 = x64[v] @ W1.t()
    bn = F.batch_norm(y1, None, None, g1, b1, True, 0.1, 1e-5)
    a = torch.relu(bn)
    a.retain_grad()
    bns.append(bn.detach()); avs.append(a)
    outs.append(tpd(tpj[3:](a)))
(torch.stack(outs) * w.double()).sum().backward()
db_m, db_t = pj[1].bias.grad.double(), b1.grad
diff = (db_m - db_t)
top = diff.abs().topk(4)
print("dbeta1 rel err", float(diff.norm() / db_t.norm()), "top cols", top.indices.tolist(), [f"{v:.3e}" for v in diff[top.indices].tolist()])
for c in top.indices.tolist()[:2]:
    for v in range(2):
        bn, g = bns[v][:, c], avs[v].grad[:, c]
        d_ = float(diff[c])
        cand = sorted(set(((g - d_).abs() < 1e-4 * max(1.0, abs(d_))).nonzero().flatten().tolist() + ((g + d_).abs() < 1e-4 * max(1.0, abs(d_))).nonzero().flatten().tolist()))
        print(f" col {c} view {v}: diff {d_:+.6e}; rows whose upstream grad matches |diff|: {cand}; their bn1 values: {[float(bn[i]) for i in cand]}; grads {[float(g[i]) for i in cand]}; min |bn1| in column: {float(bn.abs().min()):.3e}")
