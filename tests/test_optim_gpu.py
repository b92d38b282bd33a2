"""GPU parity of the multi-tensor Adam step (O1, msf_adam_multi through FusedAdam) against torch.optim.Adam -- the
optimizer of tools/ssl_train.py:303-309 -- on the same seeded parameters / gradients, including the three learning-rate
groups, the GradScaler protocol of :472-474 (scaled gradients, overflow -> skipped step) and the folded EMA update.
Floating point: fp32 arithmetic in a different association than ATen's kernels, so the bar is 1e-6 relative per step
sequence (not bit-exact), stated here."""
import pytest
import torch

from msfwsi_b200 import FusedAdam, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(64, 3, 7, 7), (64,), (4608, 512), (1152,), (5,), (4099,), (2304, 576), (1,), (128, 64, 3, 3)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.randn(*s, generator=g).to(DEV) for s in SHAPES]
    ps[-1] = ps[-1].contiguous(memory_format=torch.channels_last)  # conv weight of a channels_last model
    return [p.requires_grad_(True) for p in ps]


def _close(a, b, tol=2e-6):
    """relative to the tensor's scale: the two implementations associate the fp32 operations differently"""
    return float((a.double() - b.double()).norm()) <= tol * float(a.double().norm()) + 1e-30


def _groups(ps):
    return [{"params": ps[:3], "lr": 1e-3}, {"params": ps[3:6], "lr": 3e-3}, {"params": ps[6:], "lr": 5e-4}]


@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(gdtype, wd):
    pa, pb = _params(1), _params(1)
    ref = torch.optim.Adam(_groups(pa), lr=1e-3, weight_decay=wd)
    mine = FusedAdam(_groups(pb), lr=1e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for it in range(6):
        low = []
        for a, b in zip(pa, pb):
            gr = (torch.randn(a.shape, generator=g) * (10.0 ** ((it % 3) - 1))).to(DEV).to(gdtype)
            gr = torch.empty_like(a, dtype=gdtype).copy_(gr)  # gradient in the parameter's memory layout
            a.grad = gr.float()  # torch.optim.Adam wants grads in the parameter dtype
            b.grad = gr.clone() if gdtype == torch.float32 else None
            low.append(gr)
        ref.step()
        mine.step(grads=None if gdtype == torch.float32 else low)  # 16-bit gradient buffers go in explicitly
    for a, b in zip(pa, pb):
        assert _close(a, b), float((a - b).abs().max())
    sa, sb = ref.state_dict(), mine.state_dict()
    assert sa["param_groups"][1]["lr"] == sb["param_groups"][1]["lr"] and set(sa["state"][0]) == set(sb["state"][0])
    for k in sa["state"]:
        assert float(sb["state"][k]["step"]) == float(sa["state"][k]["step"]) == 6.0
        assert _close(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"])
        assert _close(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"])
    # the state dict of the unfused optimizer resumes the fused one (reference checkpoints, ssl_train.py:321-323)
    fresh = FusedAdam(_groups(_params(1)), lr=1e-3, weight_decay=wd)
    fresh.load_state_dict(sa)
    assert float(fresh.state_dict()["state"][0]["step"]) == 6.0


def test_fused_adam_under_gradscaler_skips_on_overflow_and_unscales():
    pa, pb = _params(3), _params(3)
    ref = torch.optim.Adam(_groups(pa), lr=1e-3)
    mine = FusedAdam(_groups(pb), lr=1e-3)
    sa = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=1000)
    sb = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=1000)
    g = torch.Generator().manual_seed(4)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.empty_like(a).copy_(torch.randn(a.shape, generator=g).to(DEV) * 1024.0)  # "scaled" gradients
            if it == 2:
                gr[(0,) * gr.dim()] = float("inf")  # overflow step: both must skip and halve the scale
            a.grad, b.grad = gr.clone(), gr.clone()
        # what scaler.scale(loss).backward() leaves behind; scaler.step / update as in ssl_train.py:473-474
        for s_, o_ in ((sa, ref), (sb, mine)):
            s_._lazy_init_scale_growth_tracker(torch.device(DEV)) if s_._scale is None else None
            s_.step(o_)
            s_.update()
    assert sa.get_scale() == sb.get_scale() == 512.0
    for a, b in zip(pa, pb):
        assert torch.isfinite(b).all() and _close(a, b)
    assert float(mine.state_dict()["state"][0]["step"]) == 4.0  # the overflow step did not count


def test_check_grads_and_folded_ema():
    ps = _params(5)
    teachers = [p.detach().clone() + 0.5 for p in ps]
    t_ref = [t.clone() for t in teachers]
    mine = FusedAdam(_groups(ps), lr=1e-2)
    mine.attach_ema(ps, teachers, momentum=0.99)
    found = torch.zeros((), device=DEV)
    g = torch.Generator().manual_seed(6)
    for p in ps:
        p.grad = torch.empty_like(p).copy_(torch.randn(p.shape, generator=g).to(DEV))
    mine.check_grads(found)
    assert float(found) == 0.0
    mine.step(found_inf=found)
    up = ops.EmaUpdater(t_ref, [p.detach() for p in ps])  # E1 on the stepped parameters = the folded update
    up.step(0.99)
    for a, b in zip(teachers, t_ref):
        assert torch.equal(a, b)
    ps[3].grad.view(-1)[1] = float("nan")
    before = [p.detach().clone() for p in ps]
    mine.check_grads(found)
    assert float(found) == 1.0
    mine.step(found_inf=found)
    assert all(torch.equal(a, b.detach()) for a, b in zip(before, ps))  # skipped
