"""2-GPU NCCL parity of the sharded InfoNCE path (keys all-gathered over NVLink, queries local) against the CPU
oracle on the concatenated batch.  Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from msfwsi_b200 import ops
        from oracle import msf_oracle as O
        res = {}
        for dim, rows, dtype, tol in ((128, 1024, torch.bfloat16, 2e-3), (512, 256, torch.bfloat16, 2e-3), (64, 200, torch.float32, 1e-5)):
            g = torch.Generator().manual_seed(7)
            k_all = torch.randn(world * rows, dim, generator=g)
            p_all = (0.3 * k_all + torch.randn(world * rows, dim, generator=g)).to(dtype)
            k_all = k_all.to(dtype)
            p = p_all[rank * rows:(rank + 1) * rows].cuda().requires_grad_(True)
            z = k_all[rank * rows:(rank + 1) * rows].cuda()
            loss = ops.infonce_loss(p, z, tau=0.07)
            loss.backward()
            t = loss.detach().clone()
            dist.all_reduce(t)
            ref, _, _ = O.infonce_loss(p_all.double(), k_all.double(), 0.07)
            gref = O.infonce_grad(p_all.double(), k_all.double(), 0.07)[rank * rows:(rank + 1) * rows]
            got = p.grad.double().cpu() / world  # DDP averages gradients over ranks
            cos = float((got.flatten() @ gref.flatten()) / (got.norm() * gref.norm()))
            assert abs(t.item() / world - ref.item()) <= tol * abs(ref.item()), (dim, t.item() / world, ref.item())
            assert cos >= 0.9999, (dim, cos)
            res[dim] = (t.item() / world, ref.item(), cos)
            # keys NOT detached (north_star (4)): per-rank msf_infonce_dk partials, NCCL reduce-scatter, finish on the owner
            p2 = p_all[rank * rows:(rank + 1) * rows].cuda().requires_grad_(True)
            z2 = k_all[rank * rows:(rank + 1) * rows].cuda().requires_grad_(True)
            ops.infonce_loss(p2, z2, tau=0.07, detach_keys=False).backward()
            kref = O.infonce_key_grad(p_all.double(), k_all.double(), 0.07)[rank * rows:(rank + 1) * rows]
            gotk = z2.grad.double().cpu() / world
            cosk = float((gotk.flatten() @ kref.flatten()) / (gotk.norm() * kref.norm()))
            assert cosk >= 0.9999, (dim, "key gradient", cosk)
            assert torch.equal(p2.grad, p.grad)  # the query gradient does not depend on the flag
        # encoder BatchNorm (fused kernels): self-synchronising statistics == BatchNorm2d over the concatenated batch
        import msfwsi_b200.resnet as R
        g = torch.Generator().manual_seed(3)
        x_all = torch.randn(world * 6, 64, 12, 12, generator=g)
        bn_ref = torch.nn.BatchNorm2d(64).train()
        x_ref = x_all.clone().requires_grad_(True)
        y_ref = bn_ref(x_ref)
        w_all = torch.randn(x_all.shape, generator=g)
        (y_ref * w_all).sum().backward()
        bn = R.FusedBatchNorm2d(64).cuda().train()
        xl = x_all[rank * 6:(rank + 1) * 6].cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = bn(xl)
        (y * w_all[rank * 6:(rank + 1) * 6].cuda()).sum().backward()
        assert torch.allclose(y.detach().cpu(), y_ref[rank * 6:(rank + 1) * 6].detach(), rtol=1e-4, atol=1e-4)
        assert torch.allclose(bn.running_var.cpu(), bn_ref.running_var, rtol=1e-4, atol=1e-5)
        assert torch.allclose(xl.grad.cpu(), x_ref.grad[rank * 6:(rank + 1) * 6], rtol=1e-3, atol=1e-4)  # cross-rank backward sums
        gw = bn.weight.grad.clone(); dist.all_reduce(gw)  # local parameter gradients sum to the full-batch gradient
        assert torch.allclose(gw.cpu(), bn_ref.weight.grad, rtol=1e-3, atol=1e-3)
        # C1: the NVLink peer-memory all-reduce the batch norms use == a plain sum over ranks, bit-identical on all ranks
        from msfwsi_b200 import ops as _ops
        red = _ops.PeerReducer.get(dist.group.WORLD, torch.device("cuda", torch.cuda.current_device()))
        assert red is not None, "symmetric memory unavailable on this box"
        for n in (1, 129, 1025, 9217):
            v = (torch.arange(n, dtype=torch.float64, device="cuda") + 1.0) * (rank + 1) * 1e-3
            want = (torch.arange(n, dtype=torch.float64, device="cuda") + 1.0) * 1e-3 * sum(range(1, world + 1))
            red.all_reduce_(v)
            assert torch.allclose(v, want, rtol=1e-15, atol=0)
            gathered = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(gathered, v)
            assert all(torch.equal(gathered[0], t) for t in gathered)
        # heads' BatchNorm1d + ReLU with cross-rank statistics == BatchNorm1d over the concatenated rows
        from msfwsi_b200.module import FusedBatchNorm1d
        z_all = torch.randn(world * 40, 128, generator=g)
        bn1_ref = torch.nn.BatchNorm1d(128).train()
        z_ref = z_all.clone().requires_grad_(True)
        o_ref = torch.relu(bn1_ref(z_ref))
        wz = torch.randn(z_all.shape, generator=g)
        (o_ref * wz).sum().backward()
        bn1 = FusedBatchNorm1d(128, act="relu").cuda().train()
        zl = z_all[rank * 40:(rank + 1) * 40].cuda().requires_grad_(True)
        o = bn1(zl)
        (o * wz[rank * 40:(rank + 1) * 40].cuda()).sum().backward()
        assert torch.allclose(o.detach().cpu(), o_ref[rank * 40:(rank + 1) * 40].detach(), rtol=1e-4, atol=1e-4)
        assert torch.allclose(bn1.running_mean.cpu(), bn1_ref.running_mean, rtol=1e-4, atol=1e-5)
        assert torch.allclose(zl.grad.cpu(), z_ref.grad[rank * 40:(rank + 1) * 40], rtol=1e-3, atol=1e-4)
        q.put((rank, "ok", res))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-800:], None))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(300)
def test_sharded_infonce_nccl_matches_single_process_oracle():
    world, port = 2, 29600 + (os.getpid() % 300)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def _worker_heads(rank, world, port, q):
    """The data-parallel promise "G ranks x B == one process x G*B" (what SyncBatchNorm + DDP give the reference,
    tools/ssl_train.py:160-170) for the grouped head stage: statistics exchanged inside msf_head_bn_finalize /
    msf_head_bn_bwd_finalize over NVLink peer memory, keys of all 24 InfoNCE pairs in ONE all-gather."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import msfwsi_b200 as M
        from oracle import msf_oracle as O

        class _Null(torch.nn.Module):
            def __init__(self, **_):
                super().__init__()
                self.fc = torch.nn.Identity()

        W = (0.1, 0.4, 0.7, 1.0)
        B, K = 24, 16
        torch.manual_seed(5)  # same weights on every rank
        sharded = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(dev).train()
        single = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(dev).train()
        single.load_state_dict(sharded.state_dict())
        mk = lambda shape, s: O.closed_form_tensor(shape, s, 1.0).abs().to(dev)
        g = torch.Generator().manual_seed(9)
        cf_all = [tuple(mk((world * B, d), 300 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
        tf_all = [tuple(mk((world * B * K, d), 400 + 10 * v + l) for l, d in enumerate(O.INTER_DIM)) for v in range(2)]
        rev_all = [torch.stack([O.jigsaw_indices(g, K)[1] for _ in range(world * B)]).to(dev) for _ in range(2)]
        sl = lambda ts, n: tuple(t[rank * n:(rank + 1) * n].contiguous() for t in ts)
        res = {}
        for mode, dtype in (("cosine", None), ("cosine", torch.bfloat16), ("infonce", torch.bfloat16)):
            cast = (lambda t: t) if dtype is None else (lambda t: t.to(dtype))
            cfa = [tuple(cast(t) for t in v) for v in cf_all]
            tfa = [tuple(cast(t) for t in v) for v in tf_all]
            outs = []
            for model, cf, tf, rev, whole in ((sharded, [sl(v, B) for v in cfa], [sl(v, B * K) for v in tfa], [r[rank * B:(rank + 1) * B] for r in rev_all], False),
                                              (single, cfa, tfa, rev_all, True)):
                model.zero_grad(set_to_none=True)
                if whole:  # the single-process run must not synchronise: hide the process group from the module
                    real = dist.is_initialized
                    dist.is_initialized = lambda: False
                try:
                    if dtype is None:
                        loss = model.heads_loss(cf[0], cf[1], tf[0], tf[1], rev, W, mode=mode)
                    else:
                        with torch.autocast("cuda", dtype=dtype):
                            loss = model.heads_loss(cf[0], cf[1], tf[0], tf[1], rev, W, mode=mode)
                    loss.backward()
                finally:
                    if whole:
                        dist.is_initialized = real
                outs.append((loss.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}))
            (ls, gs), (lw, gw) = outs
            t = ls.clone()
            dist.all_reduce(t)  # mean of the per-rank losses == the loss of the concatenated batch (equal shard sizes)
            tol = 1e-5 if dtype is None else 3e-3
            assert abs(float(t) / world - float(lw)) <= tol * max(1.0, abs(float(lw))), (mode, dtype, float(t) / world, float(lw))
            a, b = [], []
            for n in gs:
                gsum = gs[n].clone()
                dist.all_reduce(gsum)  # DDP averages the per-rank gradients
                a.append((gsum / world).flatten().double())
                b.append(gw[n].flatten().double())
            a, b = torch.cat(a), torch.cat(b)
            cos = float((a @ b) / (a.norm() * b.norm()))
            assert cos >= (0.99999 if dtype is None else 0.999), (mode, dtype, cos)
            res[f"{mode}-{dtype}"] = (float(t) / world, float(lw), cos)
        # running statistics saw the global batch on every rank
        sd_s, sd_w = sharded.state_dict(), single.state_dict()
        for k_ in sd_s:
            if k_.endswith("running_var"):
                assert torch.allclose(sd_s[k_], sd_w[k_], rtol=2e-2, atol=2e-3), k_
        q.put((rank, "ok", res))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()[-1500:], None))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(300)
def test_grouped_head_stage_two_ranks_equal_concatenated_batch():
    world, port = 2, 29900 + (os.getpid() % 90)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_heads, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res
    print(res[0][2])
