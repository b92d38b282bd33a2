"""Diagnostic: per-parameter relative error of the fp32 head stage's gradients (and torch fp32's) against an fp64 run."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import msfwsi_b200 as M
from oracle import msf_oracle as O
from oracle import torch_ref as R

DEV = "cuda:0"
W = (0.1, 0.4, 0.7, 1.0)


class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__()
        self.fc = torch.nn.Identity()


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(11)
mine = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(DEV).train()
if len(sys.argv) > 2 and sys.argv[2] == "affine":
    with torch.no_grad():
        for n, p in mine.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.uniform_(0.5, 1.5)
            elif p.dim() == 1:
                p.uniform_(-0.3, 0.3)
ref = R.RefMSFWSI(lambda **kw: _Null(**kw), 4).to(DEV).train()
ref.load_state_dict(mine.state_dict())
ref64 = R.RefMSFWSI(lambda **kw: _Null(**kw), 4).to(DEV).double().train()
ref64.load_state_dict(mine.state_dict())
mk = lambda shape, s: O.closed_form_tensor(shape, s, 1.0).abs().to(DEV)
cf = [tuple(mk((B, d), 300 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
tf = [tuple(mk((B * 16, d), 400 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
g = torch.Generator().manual_seed(1)
rev = [torch.stack([O.jigsaw_indices(g, 16)[1] for _ in range(B)]).to(DEV) for _ in range(2)]


def run(model, lossfn, dt):
    c = [tuple(t.to(dt).clone().requires_grad_(True) for t in v) for v in cf]
    t = [tuple(x.to(dt).clone().requires_grad_(True) for x in v) for v in tf]
    out = model.heads(c[0], c[1], t[0], t[1], rev)
    loss = lossfn(out)
    loss.backward()
    return out, loss, c, t


om, lm, cm, tm = run(mine, lambda o: M.ssl_loss(o, W), torch.float32)
ot, lt, ct, tt = run(ref, lambda o: R.ref_ssl_loss(o, W), torch.float32)
o6, l6, c6, t6 = run(ref64, lambda o: R.ref_ssl_loss(o, W), torch.float64)
print("loss mine %.9f torch32 %.9f fp64 %.9f" % (float(lm), float(lt), float(l6)))
names = ("ctx", "tgt", "ms")
for bi, (bm, bt, b6) in enumerate(zip(om, ot, o6)):
    for ti, tn in enumerate(("p1", "p2", "z1", "z2")):
        for l in range(4):
            print(f"out {names[bi]}_{tn}_{l}: mine {rel(bm[ti][l], b6[ti][l]):.2e} torch {rel(bt[ti][l], b6[ti][l]):.2e}")
p6 = dict(ref64.named_parameters())
pt = dict(ref.named_parameters())
for n, p in mine.named_parameters():
    em, et = rel(p.grad, p6[n].grad), rel(pt[n].grad, p6[n].grad)
    flag = " <<<<" if em > 10 * et + 1e-6 else ""
    print(f"grad {n} {tuple(p.shape)}: mine {em:.2e} torch {et:.2e}{flag}")
for v in range(2):
    for l in range(4):
        print(f"dctx v{v} l{l}: mine {rel(cm[v][l].grad, c6[v][l].grad):.2e} torch {rel(ct[v][l].grad, c6[v][l].grad):.2e}")
        print(f"dtgt v{v} l{l}: mine {rel(tm[v][l].grad, t6[v][l].grad):.2e} torch {rel(tt[v][l].grad, t6[v][l].grad):.2e}")
