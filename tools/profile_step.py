"""Development tool: torch.profiler breakdown of one bench.py step (same model, optimizer and bf16 NHWC inputs as bench.py's
resident-input step) on rank 0; run under torchrun for N > 1.  Prints the kernel table sorted by device time."""
import os, sys, warnings
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msfwsi_b200 as M

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local); torch.backends.cudnn.benchmark = True  # like bench.py
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B = int(os.environ.get("B", "256"))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = M.MSFWSI(M.resnet18, 4)
if world > 1 or os.environ.get("FORCE_SYNCBN"):
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
model = model.to(dev).to(memory_format=torch.channels_last).train()
class LossStep(torch.nn.Module):
    def __init__(s, m): super().__init__(); s.model = m
    def forward(s, x1, x2, rev): return s.model.forward_loss(x1, x2, rev, M.DEFAULT_FUSER_WEIGHTS, mode=os.environ.get("LOSS", "cosine"))
sm = LossStep(model)
if world > 1:
    sm = torch.nn.parallel.DistributedDataParallel(sm, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=128)
groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
opt = M.FusedAdam(groups, lr=1e-3)
M.bind_optimizer(model, opt)
mk = lambda n: torch.randn(n, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
c1, c2, t1, t2 = mk(B), mk(B), mk(16 * B), mk(16 * B)
rev = [torch.stack([torch.randperm(16) for _ in range(B)]).to(dev) for _ in range(2)]
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = sm((c1, t1), (c2, t2), rev)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter(); step(); t_cpu = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
if rank == 0:
    print(f"world={world} B={B}: CPU returned after {t_cpu*1e3:.0f} ms, GPU done after {t_all*1e3:.0f} ms  (CPU-bound if the two are close)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get("ROWS", "45")), max_name_column_width=90))
    if os.environ.get("TRACE"):
        prof.export_chrome_trace(os.environ["TRACE"])
if world > 1:
    dist.destroy_process_group()
