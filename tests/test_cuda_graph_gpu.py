"""The whole hot-path training step -- A1 gather/concat, grouped head stage (tcgen05 GEMMs + batch norms), fused loss, their
backward and the FusedAdam step (tools/ssl_train.py:441-474 around src/models/backbone.py:147-222) -- captured in ONE CUDA
graph: nothing in it syncs with the host, so a replay must reproduce the eager trajectory bit for bit (same kernels, same
order, deterministic reductions)."""
import pytest
import torch

import msfwsi_b200 as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__()
        self.fc = torch.nn.Identity()


def _build(seed):
    torch.manual_seed(seed)
    model = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(DEV).train()
    groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
    opt = M.FusedAdam([g for g in groups if g["params"]], lr=1e-3)
    M.bind_optimizer(model, opt)
    return model, opt


@pytest.mark.parametrize("mode", ["cosine", "infonce"])
def test_whole_step_replays_bit_identically_from_a_cuda_graph(mode):
    B, K = 32, 16
    g = torch.Generator().manual_seed(3407)
    feats = lambda n: [torch.randn(n, d, generator=g).abs().to(torch.bfloat16).to(DEV) for d in (64, 128, 256, 512)]
    c1, c2, t1, t2 = feats(B), feats(B), feats(B * K), feats(B * K)
    rev = [torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).to(DEV) for _ in range(2)]
    (ma, oa), (mb, ob) = _build(1), _build(1)
    mb.load_state_dict(ma.state_dict())

    def make(model, opt):
        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = model.heads_loss(c1, c2, t1, t2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=mode, tau=0.07)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss.detach()
        return step

    step_a, step_b = make(ma, oa), make(mb, ob)
    for _ in range(6):  # eager: 3 + 1 (warm-up inside GraphedStep) + 2 steps
        la = step_a()
    for _ in range(3):  # same history as the eager twin before the capture
        step_b()
    graphed = M.GraphedStep(lambda _inputs: step_b(), {}, ob, warmup=1)  # one more warm-up step, then the capture
    lb = graphed.loss
    for _ in range(2):
        graphed()
    torch.cuda.synchronize()
    assert torch.isfinite(la) and la.item() == lb.item()
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert torch.equal(pa, pb), n
    for (n, ba), (_, bb) in zip(ma.named_buffers(), mb.named_buffers()):
        assert torch.equal(ba, bb), n  # running statistics and num_batches_tracked advance inside the graph
    # the 16-bit GEMM operand copies follow the replayed optimizer steps too: one more eager forward agrees
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ea = ma.heads_loss(c1, c2, t1, t2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=mode, tau=0.07)
        eb = mb.heads_loss(c1, c2, t1, t2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=mode, tau=0.07)
    assert ea.item() == eb.item()
