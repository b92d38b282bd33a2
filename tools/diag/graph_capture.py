"""Experiment: capture one whole heads-only training step (A1 gather/concat -> grouped head stage -> fused loss -> backward
-> FusedAdam) in a CUDA graph and replay it; compares the replayed trajectory with eager steps from the same state."""
import copy, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import msfwsi_b200 as M

dev = torch.device("cuda", 0)
B, K = int(os.environ.get("B", "256")), 16
mode = os.environ.get("LOSS", "cosine")

class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__(); self.fc = torch.nn.Identity()

def build(seed):
    torch.manual_seed(seed)
    model = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(dev).train()
    groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
    groups = [g for g in groups if g["params"]]
    opt = M.FusedAdam(groups, lr=1e-3)
    M.bind_optimizer(model, opt)
    return model, opt

g = torch.Generator().manual_seed(3407)
feats = lambda n: [torch.randn(n, d, generator=g).abs().to(torch.bfloat16).to(dev) for d in (64, 128, 256, 512)]
c1, c2, t1, t2 = feats(B), feats(B), feats(B * K), feats(B * K)
rev = [torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).to(dev) for _ in range(2)]

def make_step(model, opt):
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.heads_loss(c1, c2, t1, t2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=mode, tau=0.07)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    return step

model_a, opt_a = build(1)
model_b, opt_b = build(1)
model_b.load_state_dict(model_a.state_dict())
step_a, step_b = make_step(model_a, opt_a), make_step(model_b, opt_b)

# eager reference: 3 warm-up + 4 more steps
for _ in range(7):
    la = step_a()
torch.cuda.synchronize()

# graph: 3 warm-up steps on a side stream, capture the 4th, replay 3 more
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step_b()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
opt_b.zero_grad(set_to_none=True)
with torch.cuda.graph(graph, capture_error_mode=os.environ.get("CAPTURE_MODE", "global")):
    static_loss = step_b()
for _ in range(4):
    graph.replay()
torch.cuda.synchronize()
print("eager loss after 7 steps", float(la), "graph loss (step 7)", float(static_loss))
worst = 0.0
for (n, pa), (_, pb) in zip(model_a.named_parameters(), model_b.named_parameters()):
    worst = max(worst, float((pa - pb).abs().max()))
print("max |param_eager - param_graph| after 7 steps:", worst)
bufw = max(float((a.float() - b.float()).abs().max()) for (_, a), (_, b) in zip(model_a.named_buffers(), model_b.named_buffers()))
print("max buffer difference:", bufw)

def bench(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
print("eager  ms/step (device, wall):", bench(step_a))
print("graph  ms/step (device, wall):", bench(graph.replay))
