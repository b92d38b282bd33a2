"""Latency of the cross-rank exchanges at G GPUs, back to back (no compute between calls, so no rank skew):
msf_peer_allreduce_f64 (encoder batch-norm statistics, C1) at the sizes the encoders use, and an NCCL all-reduce of the
same vector for comparison.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_peer_exchange.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import ops  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    red = ops.PeerReducer.get(dist.group.WORLD, dev)
    assert red is not None, "no peer memory"
    rows = []
    for n in (129, 257, 1025):  # 2C+1 doubles for C = 64, 128, 512
        v = torch.randn(n, dtype=torch.float64, device=dev)
        for name, fn in (("peer", lambda: red.all_reduce_(v)), ("nccl", lambda: dist.all_reduce(v))):
            for _ in range(20):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 200
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1000 / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows.append({"gpus": world, "doubles": n, "kind": name, "us_per_call_back_to_back": round(float(t.item()), 2)})
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
