"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: rank-major key all-gather, positive-index
offsets, and the data-parallel promise "G ranks x B == one process x G*B" for the InfoNCE extension."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from msfwsi_b200 import ops
        from oracle import msf_oracle as O
        rows, dim, tau = 12, 32, 0.07
        p_all = O.closed_form_tensor((world * rows, dim), 1.0, 1.0).double()
        z_all = O.closed_form_tensor((world * rows, dim), 2.0, 1.0).double()
        p, z = p_all[rank * rows:(rank + 1) * rows], z_all[rank * rows:(rank + 1) * rows]
        keys, off = ops.all_gather_keys(z.contiguous())
        assert off == rank * rows
        assert torch.equal(keys, z_all), "all-gather must be rank-major"
        loss_local, _, _ = O.infonce_loss(p, keys, tau, pos_offset=off)
        grad_local = O.infonce_grad(p, keys, tau, off, n_rows_global=rows)  # d(local mean)/dp
        # DDP averages gradients over ranks; the loss meter averages the per-rank losses
        t = torch.stack([loss_local.detach()])
        dist.all_reduce(t)
        loss_ddp = t / world
        loss_single, _, _ = O.infonce_loss(p_all, z_all, tau)
        grad_single = O.infonce_grad(p_all, z_all, tau)[rank * rows:(rank + 1) * rows]
        assert abs(loss_ddp.item() - loss_single.item()) < 1e-12
        assert torch.allclose(grad_local / world, grad_single, rtol=1e-10, atol=1e-14)
        # keys NOT detached (north_star (4)): every rank's queries touch every key; the reduce-scatter (backward of the
        # rank-major all-gather, as ops._infonce_key_grad issues it) hands each rank the summed rows of its own keys
        part = O.infonce_key_grad(p, keys, tau, off, n_rows_global=rows).contiguous()
        mine = torch.empty((rows, dim), dtype=part.dtype)
        dist.reduce_scatter_tensor(mine, part, op=dist.ReduceOp.SUM)
        single = O.infonce_key_grad(p_all, z_all, tau)[rank * rows:(rank + 1) * rows]
        assert torch.allclose(mine / world, single, rtol=1e-10, atol=1e-14)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_infonce_equals_single_process():
    world, port = 2, 29500 + (os.getpid() % 500)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_all_gather_keys_single_process_passthrough():
    from msfwsi_b200 import ops
    z = torch.randn(4, 8)
    keys, off = ops.all_gather_keys(z)
    assert keys is z and off == 0
