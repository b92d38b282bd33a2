"""Development probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm.empty(4096, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; world", hdl.world_size, "rank", hdl.rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4],
          "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:4], "has dev ptr arrays:", hasattr(hdl, "buffer_ptrs_dev"), flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.float32)
    print(rank, "peer value", float(peer[0]), "(expected", float((rank + 1) % world + 1), ")", flush=True)
    hdl.barrier()
    # latency of a small NCCL all-reduce vs torch's one-shot symmetric-memory all-reduce
    import time
    x = torch.ones(1025, dtype=torch.float64, device=dev)
    for _ in range(20): dist.all_reduce(x)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): dist.all_reduce(x)
    b.record(); torch.cuda.synchronize()
    if rank == 0: print("NCCL all_reduce 1025 x f64: %.1f us per call (GPU timeline, back to back)" % (a.elapsed_time(b) * 1000 / 200), flush=True)
    try:
        y = symm.empty(2048, dtype=torch.float32, device=dev); symm.rendezvous(y, dist.group.WORLD); y.fill_(1.0)
        for _ in range(20): torch.ops.symm_mem.one_shot_all_reduce(y, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize(); a.record()
        for _ in range(200): torch.ops.symm_mem.one_shot_all_reduce(y, "sum", dist.group.WORLD.group_name)
        b.record(); torch.cuda.synchronize()
        if rank == 0: print("symm_mem one_shot_all_reduce 2048 x f32: %.1f us per call" % (a.elapsed_time(b) * 1000 / 200), flush=True)
    except Exception as e:
        print(rank, "one_shot_all_reduce failed:", repr(e)[:300], flush=True)
except Exception as e:
    import traceback; print(rank, "symmetric memory FAILED:", repr(e)[:500], flush=True); traceback.print_exc()
dist.destroy_process_group()
