"""c5 microbench at G GPUs (BASELINE.json configs[4]): fused InfoNCE fwd+bwd with the queries sharded N/G per rank and
the keys global (NCCL all-gather of the normalised bf16 keys, rank-major, positives at rank*N/G + i -- SURVEY 8e).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_infonce_dist.py [--rows 65536] [--dims 128 256]

One timed iteration per rank = row-normalise local q and k, all-gather the keys, forward (loss + O partials), backward
(grad_q).  CUDA events on the launching stream, L2 flushed between iterations, MAX over ranks; whole-job TFLOP/s =
4*N*N*D / time (algorithmic FLOPs of all ranks together); peak = G x MEASURED_PEAKS.json bf16_tflops."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L  # noqa: E402
from msfwsi_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", dest="n", type=int, nargs="+", default=[16384, 65536])
    ap.add_argument("--dims", dest="d", type=int, nargs="+", default=[128, 256])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tau", type=float, default=0.07)
    ap.add_argument("--out", default=None)
    ap.add_argument("--key-grad", action="store_true",
                    help="also time the NON-detached variant (north_star (4)): forward + dq + msf_infonce_dk + NCCL reduce-scatter + finish")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak1 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        peak1 = 1590.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for d in args.d:
        for n in args.n:
            nq = n // world
            g = torch.Generator(device=dev).manual_seed(3407)  # same stream on every rank: rank r takes rows [r*nq, (r+1)*nq)
            k_all = torch.randn(n, d, device=dev, generator=g)
            q_all = 0.3 * k_all + torch.randn(n, d, device=dev, generator=g)
            q = q_all[rank * nq:(rank + 1) * nq].to(torch.bfloat16).contiguous()
            k = k_all[rank * nq:(rank + 1) * nq].to(torch.bfloat16).contiguous()
            del k_all, q_all
            # the grouped entry points of the training step: keys as rank-major blocks (exactly what one all-gather leaves behind),
            # read through the 3-D TMA map {dim, row, rank}; positives in block `rank`
            probe = (L.NcePair * 1)(L.NcePair(q.data_ptr(), 0, q.data_ptr(), q.data_ptr(), nq * d, nq, nq, world, d, rank, 1.0))
            ws_bytes = L.lib().msf_nce_grouped_workspace_bytes(probe, 1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            gout = torch.ones((), device=dev)
            gq = torch.empty_like(q)
            kgath = torch.empty(world * nq, d, dtype=torch.bfloat16, device=dev)
            st = L.stream_ptr()

            def step():
                qh, qi = ops.rownorm(q, torch.bfloat16)
                kh, _ = ops.rownorm(k, torch.bfloat16)
                if world > 1:
                    dist.all_gather_into_tensor(kgath, kh)
                    keys = kgath
                else:
                    keys = kh
                pr = (L.NcePair * 1)(L.NcePair(qh.data_ptr(), 0, keys.data_ptr(), gq.data_ptr(), nq * d, nq, nq, world, d, rank, 1.0))
                L.check(L.lib().msf_nce_grouped_fwd(pr, 1, L.MSF_BF16, args.tau, 1e-8, loss.data_ptr(), ws.data_ptr(), ws_bytes, st), "fwd")
                L.check(L.lib().msf_nce_grouped_bwd(pr, 1, L.MSF_BF16, args.tau, 1e-8, gout.data_ptr(), ws.data_ptr(), ws_bytes, st), "bwd")

            for _ in range(3):
                step()
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ts.append(float(t.item()))
            ts.sort()
            ms = ts[len(ts) // 2]
            total = loss.detach().clone()
            if world > 1:
                dist.all_reduce(total)
            flops = 4.0 * n * n * d
            row = {"gpus": world, "N": n, "Nq_per_gpu": nq, "D": d, "mean_loss": float(total.item()) / world, "ms_fwd_bwd": ms,
                   "tflops_whole_job": flops / ms / 1e9, "frac_of_peak": flops / ms / 1e9 / (peak1 * world), "peak_tflops_per_gpu": peak1}
            if args.key_grad:
                prec = L.MSF_BF16
                n_keys = world * nq
                wsb = L.lib().msf_infonce_workspace_bytes(nq, n_keys, d, prec)
                ws2 = torch.empty(wsb, dtype=torch.uint8, device=dev)
                dwb = L.lib().msf_infonce_dk_workspace_bytes(nq, n_keys, d, prec)
                dws = torch.empty(dwb, dtype=torch.uint8, device=dev)
                dk_all = torch.empty((n_keys, d), dtype=torch.float32, device=dev)
                dk_loc = torch.empty((nq, d), dtype=torch.float32, device=dev)
                gz = torch.empty_like(k)
                off = rank * nq

                def step_kg():
                    qh, qi = ops.rownorm(q, torch.bfloat16)
                    kh, ki = ops.rownorm(k, torch.bfloat16)
                    if world > 1:
                        dist.all_gather_into_tensor(kgath, kh)
                        keys = kgath
                    else:
                        keys = kh
                    lib = L.lib()
                    L.check(lib.msf_infonce_fwd(qh.data_ptr(), keys.data_ptr(), nq, n_keys, d, off, args.tau, prec, loss.data_ptr(), 0, ws2.data_ptr(), wsb, st), "fwd")
                    L.check(lib.msf_infonce_bwd(qh.data_ptr(), keys.data_ptr(), qi.data_ptr(), nq, n_keys, d, off, args.tau, prec, gout.data_ptr(), 1.0 / nq,
                                                ws2.data_ptr(), wsb, gq.data_ptr(), L.MSF_BF16, st), "bwd")
                    L.check(lib.msf_infonce_dk(qh.data_ptr(), keys.data_ptr(), nq, n_keys, d, off, args.tau, prec, gout.data_ptr(), 1.0 / nq, ws2.data_ptr(), wsb,
                                               dk_all.data_ptr(), dws.data_ptr(), dwb, st), "dk")
                    if world > 1:
                        dist.reduce_scatter_tensor(dk_loc, dk_all)
                        mine = dk_loc
                    else:
                        mine = dk_all
                    L.check(lib.msf_infonce_dk_finish(mine.data_ptr(), qh.data_ptr(), kh.data_ptr(), ki.data_ptr(), nq, nq, d, args.tau, prec, gout.data_ptr(),
                                                      1.0 / nq, gz.data_ptr(), L.MSF_BF16, st), "finish")

                for _ in range(3):
                    step_kg()
                ts = []
                for _ in range(args.iters):
                    flush.zero_()
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    step_kg()
                    e1.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ts.append(float(t.item()))
                ts.sort()
                ms_kg = ts[len(ts) // 2]
                row["key_grad"] = {"ms_fwd_dq_dk_reduce_scatter": ms_kg, "reduce_scatter_bytes_per_rank": n_keys * d * 4,
                                   "frac_of_peak_6NND_algorithmic": 6.0 * n * n * d / ms_kg / 1e9 / (peak1 * world),
                                   "frac_of_peak_8NND_executed": 8.0 * n * n * d / ms_kg / 1e9 / (peak1 * world)}
                del ws2, dws, dk_all, dk_loc
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0 and args.out:
        json.dump(rows, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
