"""Diag: per-launch device time of every kernel of one heads-only step (torch.profiler), in launch order."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import msfwsi_b200 as M
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
B, K = int(os.environ.get("B", "256")), 16
class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__(); self.fc = torch.nn.Identity()
torch.manual_seed(1)
model = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(dev).train()
groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
opt = M.FusedAdam([g for g in groups if g["params"]], lr=1e-3); M.bind_optimizer(model, opt)
g = torch.Generator().manual_seed(3407)
feats = lambda n: [torch.randn(n, d, generator=g).abs().to(torch.bfloat16).to(dev) for d in (64, 128, 256, 512)]
c1, c2, t1, t2 = feats(B), feats(B), feats(B * K), feats(B * K)
rev = [torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).to(dev) for _ in range(2)]
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model.heads_loss(c1, c2, t1, t2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=os.environ.get("LOSS", "cosine"), tau=0.07)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(4): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = 0.0
for e in evs:
    name = e.name.replace("void msf::(anonymous namespace)::", "")[:60]
    print(f"{e.device_time:9.1f} us  {name}")
    tot += e.device_time
print("total device time %.1f us in %d kernels" % (tot, len(evs)))
