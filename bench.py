"""bench.py -- SSL pre-training tiles/s of the MSF-WSI hot path on N B200s + the rooflines of its kernels.

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port) on host cores

A "step" is one full pre-training step of BASELINE.json configs[1] (BCSS-shaped, per-GPU batch 256, bf16 autocast,
synthetic 1024^2-tile-shaped inputs: 2 context + 32 target 224^2 views and 2 jigsaw index rows per tile):
ResNet-18 encoders on PyTorch/cuDNN (not a CUDA target of this repo) -> hot path (gather/concat kernel, heads,
fused loss kernel) -> backward -> Adam.  The default objective is the REFERENCE's (SimSiam negative cosine,
tools/ssl_train.py:448-466) so that every arm -- this repo, the GPU-eager reference graph, the CPU port -- times the same
thing; the same step with the north-star InfoNCE objective is timed next to it (`infonce_step`).

JSON line (rank 0): `value` = K steps with inputs resident in HBM; `e2e` = the same step fed from pinned host memory
with the loss read back every step; `roofline` = the north-star kernel (fused InfoNCE forward+backward chain at the c5
size N = 65536) against the measured bf16 tensor peak; `roofline_kernels` = every kernel of this repo inside the timed
steps (library profiler); `roofline_hbm` = A1 / A2 / E1 at sizes beyond L2 against the measured copy bandwidth;
`gpu_eager_baseline` = the reference's own module graph (stock torch modules, torch.optim.Adam, oracle/torch_ref.py) on
the same GPU(s) in the same job; `cpu_baseline` = the same graph on the host cores (N = 1 only); `dist_parity` (N > 1) =
sharded-vs-gathered InfoNCE, peer all-reduce vs NCCL, synced batch norm vs the concatenated batch, checked before timing.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ssl_pretrain_tiles_per_sec"
UNIT = "tiles/s"
FUSER_WEIGHTS = (0.1, 0.4, 0.7, 1.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch in tiles (configs[1]: 256; configs[3] c4: 1024 with --heads-only)")
    ap.add_argument("--img", type=int, default=224)
    ap.add_argument("--loss", default="cosine", choices=["cosine", "infonce"],
                    help="cosine = the reference's SimSiam objective (default, like-for-like with every baseline arm); "
                         "infonce = the north-star extension.  The other one is timed as well (`*_step` keys)")
    ap.add_argument("--tau", type=float, default=0.07)
    ap.add_argument("--cpu-sample", type=int, default=8, help="tiles per CPU-baseline step (configs[0]: batch 8)")
    ap.add_argument("--heads-only", action="store_true",
                    help="time the hot path alone (gather/concat + heads + loss + backward + Adam on the head parameters) from "
                         "synthetic pooled features: the c4 regime (--batch 1024) does not fit one GPU with the encoders attached")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="N = 1: capture the whole step (forward, backward, optimizer) in ONE CUDA graph after the warm-up and time its replays; "
                         "`value` / `e2e` are then the replayed step, the eager step is reported beside it (`eager_step`)")
    ap.add_argument("--no-hot-path", action="store_true", help="skip the heads-only eager vs CUDA-graph row of the default N = 1 run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-hbm-rows", action="store_true")
    ap.add_argument("--no-second-loss", action="store_true")
    ap.add_argument("--e2e-source", default="uint8", choices=["uint8", "bf16"],
                    help="what the end-to-end step copies from the host: uint8 = the 1024^2 source tiles + crop boxes / flips / jigsaw "
                         "permutations, views built on the device (msf_view_crops_s2d); bf16 = the 34 normalised views per tile")
    return ap.parse_args()


def workload_name(args):
    if args.heads_only:
        return (f"MSF-WSI hot path only (gather/concat + 12 projectors + 12 predictors + loss, fwd+bwd+Adam), bf16, batch {args.batch}/GPU, "
                f"synthetic pooled pyramid features, loss={args.loss}")
    return (f"BCSS fold-0 SSL pretraining bf16, batch {args.batch}/GPU, synthetic L0_1024_s512-shaped tiles "
            f"(2x(3,{args.img},{args.img}) context + 2x(16,3,{args.img},{args.img}) target views), loss={args.loss}")


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's module graph (oracle port) on host cores
# ------------------------------------------------------------------------------------------------------------
def run_cpu_port(sample_tiles, steps, warmup, loss, tau, img):
    from oracle.cpu_step import ReferenceStep  # bench.py executes oracle/ only in its baseline legs
    port = ReferenceStep(sample_tiles, img=img, loss=loss, tau=tau, device="cpu")
    for _ in range(warmup):
        port.step()
    times = [port.step() for _ in range(steps)]
    total = sum(times)
    return {"value": sample_tiles * steps / total, "unit": UNIT, "cores": port.threads, "kind": "port",
            "sample": f"{steps} step(s) of {sample_tiles} tiles ({34 * sample_tiles} encoder images of {img}^2) after {warmup} warm-up, fp32, "
                      f"stock torch modules on the CPU (oracle/torch_ref.py: plain ResNet-18 encoders + nn.Linear/nn.BatchNorm1d heads + "
                      f"{loss} loss + torch.optim.Adam), full step incl. backward",
            "ms_per_step": 1e3 * total / steps}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    steps, warmup = min(max(1, args.steps), 3), max(0, min(args.warmup, 1))  # bounded: each CPU step costs seconds
    cb = run_cpu_port(args.cpu_sample, steps, warmup, args.loss, args.tau, args.img)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args), "cpu_sample_tiles": args.cpu_sample},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# ------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.f.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    import msfwsi_b200 as M
    from msfwsi_b200 import _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(3407 + rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dist_parity = None
    if world > 1:
        dist_parity = check_dist_parity(torch, dist, M, ops, dev, rank, world)  # raises when out of tolerance

    class LossStep(torch.nn.Module):  # forward() returns the loss so DDP hooks see the whole hot path
        def __init__(self, model):
            super().__init__()
            self.model, self.mode = model, args.loss

        def forward(self, x1, x2, rev):
            if args.heads_only:
                return self.model.heads_loss(x1[0], x2[0], x1[1], x2[1], rev, FUSER_WEIGHTS, mode=self.mode, tau=args.tau)
            return self.model.forward_loss(x1, x2, rev, FUSER_WEIGHTS, mode=self.mode, tau=args.tau)

    class _NoEncoder(torch.nn.Module):
        def __init__(self, **_):
            super().__init__()
            self.fc = torch.nn.Identity()

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc = (lambda **kw: _NoEncoder(**kw)) if args.heads_only else M.resnet18  # random init: no network for ImageNet weights
        model = M.MSFWSI(enc, 4, 2048, 512, 0.5, False)
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)  # tools/ssl_train.py:160
    model = model.to(dev).to(memory_format=torch.channels_last).train()
    loss_step = LossStep(model)
    step_mod = loss_step
    if world > 1:
        # ssl_train.py:170; SyncBN keeps the buffers identical on every rank, so the per-forward buffer broadcast is
        # redundant; large buckets suit NVSwitch (latency-, not link-bound)
        step_mod = torch.nn.parallel.DistributedDataParallel(step_mod, device_ids=[local], broadcast_buffers=False,
                                                             gradient_as_bucket_view=True, bucket_cap_mb=128)
    lr = 1e-3 * (args.batch * world) ** 0.5 / 32 ** 0.5  # ssl_train.py:155
    groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
    groups = [g for g in groups if g["params"]]
    opt = M.FusedAdam(groups, lr=lr)  # torch.optim.Adam semantics (ssl_train.py:309), one msf_adam_multi launch per step
    M.bind_optimizer(model, opt)      # the kernel also rewrites the bf16 GEMM operands of the head Linears (no cast pass per step)

    B, K, img = args.batch, 16, args.img
    g = torch.Generator().manual_seed(3407 + rank)
    # Host side of the input pipeline: normalised views in the layout and precision the first convolution consumes under
    # bf16 autocast (NHWC, bf16 -- autocast would round the fp32 tensor to exactly these values on the device), pinned.
    if args.heads_only:
        def feats(n):
            return [torch.randn(n, d, generator=g).abs().to(torch.bfloat16).pin_memory() for d in (64, 128, 256, 512)]
        host = {"c1": feats(B), "c2": feats(B), "t1": feats(B * K), "t2": feats(B * K)}
    else:
        def view(n):
            return torch.randn(n, 3, img, img, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).pin_memory()
        host = {"c1": view(B), "c2": view(B), "t1": view(B * K), "t2": view(B * K)}
    host["r1"] = torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).pin_memory()
    host["r2"] = torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).pin_memory()

    def nbytes(v):
        return sum(nbytes(t) for t in v) if isinstance(v, (list, tuple)) else v.numel() * v.element_size()

    h2d_bytes = sum(nbytes(v) for v in host.values())
    MEAN, STD = (0.6998, 0.4785, 0.6609), (0.2203, 0.2407, 0.1983)  # scripts/bcss.sh:13-14
    host_u8 = None
    if not args.heads_only and args.e2e_source == "uint8":
        # What a dataloader hands over once the augmentations' random numbers are drawn (tools/ssl_train.py:175-217,
        # src/utils/data/bcss.py:164-182): the uint8 L0 tile, per view a RandomResizedCrop(scale=(0.5, 1)) box + flip bit, per
        # target view the jigsaw permutation.  Everything else (tiling, shuffle, crop, resize, flip, normalise, layout) runs
        # on the device in one launch.
        S, T = 1024, 256

        def rrc(n, side):  # albumentations / torchvision RandomResizedCrop box sampling, integer boxes
            area = side * side * (0.5 + 0.5 * torch.rand(n, generator=g))
            ratio = torch.exp(torch.empty(n).uniform_(-0.2876820724517809, 0.2876820724517809, generator=g))  # log(3/4) .. log(4/3)
            w = torch.sqrt(area * ratio).round().clamp(8, side).long()
            h = torch.sqrt(area / ratio).round().clamp(8, side).long()
            y0 = (torch.rand(n, generator=g) * (side - h + 1).float()).long().clamp(max=side - 1)
            x0 = (torch.rand(n, generator=g) * (side - w + 1).float()).long().clamp(max=side - 1)
            return torch.stack((y0, x0, torch.minimum(y0 + h, torch.tensor(side)), torch.minimum(x0 + w, torch.tensor(side))), dim=1)

        perms = [torch.stack([torch.randperm(K, generator=g) for _ in range(B)]) for _ in range(2)]
        crop_rows = []
        for v in range(2):  # context views: crop of the whole tile
            bx = rrc(B, S)
            crop_rows.append(torch.cat((torch.arange(B).view(B, 1), bx, torch.randint(0, 2, (B, 1), generator=g)), dim=1))
        for v in range(2):  # target views: tile perm[b, j] of blockshaped(img, 256, 256), cropped inside the tile
            bx = rrc(B * K, T).view(B, K, 4)
            crop_rows.append(ops.jigsaw_view_crops(perms[v], bx, torch.randint(0, 2, (B, K), generator=g), S, S, 4).long())
        host_u8 = {"src": torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, generator=g).pin_memory(),
                   "crops": torch.cat(crop_rows).to(torch.int32).pin_memory(),
                   "r1": perms[0].argsort(dim=1).pin_memory(), "r2": perms[1].argsort(dim=1).pin_memory()}  # rev = argsort(perm), bcss.py:172
        h2d_bytes = sum(nbytes(v) for v in host_u8.values())

    def to_dev():
        mv = lambda v: [t.to(dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(dev, non_blocking=True)
        return {k: mv(v) for k, v in host.items()}

    def step(d):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = step_mod((d["c1"], d["t1"]), (d["c2"], d["t2"]), [d["r1"], d["r2"]])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    resident = to_dev()
    for _ in range(max(3, args.warmup)):
        step(resident)
    barrier()

    def timed(n_steps, profile):
        clocks = Clocks(local) if (rank == 0 and profile) else None
        if profile:
            _lib.prof_begin(1 << 16)  # CUDA event pairs around every main kernel of this repo, on its launching stream
        launches0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        # process-wide start/end range (push/pop ranges are per thread and would miss the autograd thread's launches):
        # ncu --nvtx --nvtx-include "msf_timed_steps" captures exactly the launches of the timed steps
        rng = torch.cuda.nvtx.range_start("msf_timed_steps") if profile else None
        e0.record()
        for _ in range(n_steps):
            loss = step(resident)
        e1.record()
        barrier()
        if profile:
            torch.cuda.nvtx.range_end(rng)
        ms = max_over_ranks(e0.elapsed_time(e1))
        out = {"ms": ms, "launches": _lib.launch_count - launches0, "loss": float(loss.item())}
        if profile:
            out["prof"], out["prof_dropped"] = _lib.prof_end()
            out["clocks"] = clocks.stop() if clocks else None
        return out

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------------
    main = timed(args.steps, True)
    ms, launches, prof, prof_dropped, clk, last_loss = main["ms"], main["launches"], main["prof"], main["prof_dropped"], main["clocks"], main["loss"]
    value = B * world * args.steps / (ms / 1e3)

    # ---- the same step as ONE CUDA graph (launch-bound regime: the heads-only hot path spends its time in Python / autograd /
    # launch overhead, not on the GPU).  Nothing in the step syncs with the host, every kernel argument is a device pointer
    # from the graph's private pool or a value fixed at capture, the optimizer's pointer table is uploaded from pinned memory:
    # the capture needs no special path.  Single GPU only: the cross-rank exchanges carry host-side sequence numbers.
    eager_step = None
    if args.cuda_graph:
        if world > 1 or not args.heads_only:
            # the full step keeps the GPU busy for ~193 ms against ~46 ms of host time: nothing to gain from a capture there
            raise SystemExit("--cuda-graph times the launch-bound hot path: use it with --heads-only at N = 1")
        eager_step = {"value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "gpu_launches": launches}
        graphed = M.GraphedStep(step, resident, opt, warmup=3)
        step = graphed  # noqa: F811 -- from here on a step is a replay (inputs are copied into the captured tensors first)
        for _ in range(3):
            step(resident)
        main = timed(args.steps, False)
        ms, launches, last_loss = main["ms"], main["launches"], main["loss"]
        value = B * world * args.steps / (ms / 1e3)
        eager_step["speedup_graph_over_eager"] = value / eager_step["value"]

    other = None
    if not args.no_second_loss and not args.cuda_graph:
        loss_step.mode = "infonce" if args.loss == "cosine" else "cosine"
        for _ in range(3):
            step(resident)
        o = timed(args.steps, False)
        other = {"loss_mode": loss_step.mode, "value": B * world * args.steps / (o["ms"] / 1e3), "unit": UNIT, "ms_per_step": o["ms"] / args.steps,
                 "gpu_launches": o["launches"], "loss": o["loss"]}
        loss_step.mode = args.loss

    # ---- timed region 2: end to end from pinned host memory, loss read back every step -------------------
    # Every step's inputs are copied host -> device inside the timed region (K copies for K steps); the copy of step
    # i+1 is issued on a side stream while step i computes (double buffering), the first copy is exposed.
    if not args.cuda_graph:
        del resident
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def fetch():
        with torch.cuda.stream(copy_stream):
            d = to_dev() if host_u8 is None else {k: v.to(dev, non_blocking=True) for k, v in host_u8.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for v in d.values():
            for t in (v if isinstance(v, list) else [v]):
                t.record_stream(main_stream)
        return d, ev

    def build_views(d):
        """uint8 tiles + crop rows -> the 34 views per tile in the stem's input layout, one launch (D1b)."""
        if host_u8 is None:
            return d
        views = ops.view_crops_s2d(d["src"], d["crops"], (img, img), MEAN, STD, torch.bfloat16)
        return {"c1": views[:B], "c2": views[B:2 * B], "t1": views[2 * B:2 * B + B * K], "t2": views[2 * B + B * K:], "r1": d["r1"], "r2": d["r2"]}

    def e2e_steps(n):
        nxt = fetch()
        for i in range(n):
            d, ev = nxt
            main_stream.wait_event(ev)
            nxt = fetch() if i + 1 < n else None
            step(build_views(d)).item()  # device -> host read of the step's result

    e2e_steps(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # ---- every kernel of this repo inside the timed steps -------------------------------------------------
    # per kernel family: algorithmic work (bytes or FLOP) summed over the launches of the timed
    # region / device time between the event pairs the library recorded around them
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    hbm_peak, tc_peak = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    kernels = []
    for name, r in prof.items():
        if r["ms"] <= 0:
            continue
        if r["bound"] == "l":  # latency-bound exchange (NVLink round trip): no throughput roofline, report the time
            kernels.append({"kernel": name, "bound": "latency", "launches": r["launches"], "ms_per_step": r["ms"] / args.steps,
                            "share_of_step": r["ms"] / ms, "avg_us": 1e3 * r["ms"] / r["launches"]})
            continue
        if r["bound"] == "h":
            ach, peak, unit, bound = r["work"] / r["ms"] / 1e6, hbm_peak, "GB/s", "hbm"
        else:
            ach, peak, unit, bound = r["work"] / r["ms"] / 1e9, tc_peak, "TFLOP/s", "tensor"
        kernels.append({"kernel": name, "bound": bound, "launches": r["launches"], "ms_per_step": r["ms"] / args.steps,
                        "share_of_step": r["ms"] / ms, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                        "work_per_launch": r["work"] / r["launches"], "avg_us": 1e3 * r["ms"] / r["launches"]})
    kernels.sort(key=lambda k: -k["ms_per_step"])

    # ---- the contract's roofline: the north-star kernel (fused InfoNCE fwd+bwd, c5 size), measured in this process ----
    micro = roofline = None
    if rank == 0 and not args.no_microbench:
        micro = infonce_microbench(torch, ops, _lib, dev, peaks, traffic)
        roofline = micro["roofline"]
    hbm_rows = None
    if rank == 0 and not args.no_hbm_rows:
        hbm_rows = hbm_microbench(torch, ops, _lib, dev, peaks)

    # ---- baselines: the reference's own module graph on the same GPU(s) (all ranks), and on the host cores (N = 1) ----
    del model, loss_step, step_mod, opt
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    hot_path = None
    if rank == 0 and world == 1 and not args.heads_only and not args.no_hot_path:
        hot_path = hot_path_graph_bench(torch, M, _lib, dev, args)
    gpu_eager = None
    if not args.no_gpu_eager and not args.heads_only:
        gpu_eager = gpu_eager_baseline(torch, dist, args, dev, world, rank, barrier, max_over_ranks, value, e2e)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.heads_only:
        cb = run_cpu_port(args.cpu_sample, 2, 1, args.loss, args.tau, args.img)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": workload_name(args), "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                           "encoder": "none (heads only)" if args.heads_only else "resnet18 random-init (PyTorch/cuDNN, channels_last)",
                           "optimizer": "Adam, 3 lr groups (msf_adam_multi, bf16 GEMM operands rewritten in the same pass)",
                           "l2_policy": "per-step inputs (2.6 GB bf16) and activations exceed the 126 MB L2; microbenchmarks flush L2 between iterations",
                           "host_inputs": ("e2e: uint8 1024^2 source tiles + crop boxes / flips / jigsaw permutations, pinned; the 34 views per tile are built on "
                                           "the device in one launch (msf_view_crops_s2d)" if host_u8 is not None else
                                           "bf16 NHWC pinned (what the first convolution consumes under bf16 autocast)")},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline, "loss": last_loss,
                "encoder_images_per_sec": None if args.heads_only else value * 34}
        if hot_path:
            line["hot_path_cuda_graph"] = hot_path
        if eager_step:
            line["eager_step"] = eager_step
            line["config"]["execution"] = "one CUDA graph per step (captured after warm-up, replayed)"
        if other:
            line[other["loss_mode"] + "_step"] = other
        if gpu_eager:
            line["gpu_eager_baseline"] = gpu_eager
        if dist_parity:
            line["dist_parity"] = dist_parity
        line["roofline_kernels"] = kernels
        line["roofline_kernels_note"] = ("cudaEventRecord by the library on the launching stream immediately before/after each kernel, summed over the "
                                         f"timed steps; share of the step covered: {sum(k['share_of_step'] for k in kernels):.3f}; events dropped: {prof_dropped}")
        if micro:
            line["roofline_microbench"] = micro["rows"]
        if hbm_rows:
            line["roofline_hbm"] = hbm_rows
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# multi-GPU parity, checked on the box before anything is timed (N > 1)
# ------------------------------------------------------------------------------------------------------------
def check_dist_parity(torch, dist, M, ops, dev, rank, world):
    """Max errors of the cross-rank pieces of the path against their single-process definitions:
    (a) sharded InfoNCE (local queries x all-gathered keys) vs the loss / gradient of the gathered problem on one rank, and
    the reduce-scattered key gradient of the non-detached variant,
    (b) the NVLink peer all-reduce vs NCCL's, (c) batch norm with synced statistics vs batch norm of the concatenated batch.
    Raises if a bound is exceeded; the numbers go into the JSON line."""
    out = {}
    g = torch.Generator(device=dev).manual_seed(99)  # same stream on every rank: everyone can build the global problem
    rows, tau = 512, 0.07
    for dim in (128, 576):
        z_all = torch.randn(world * rows, dim, device=dev, generator=g)
        p_all = (0.5 * z_all + torch.randn(world * rows, dim, device=dev, generator=g)).to(torch.bfloat16)
        z_all = z_all.to(torch.bfloat16)
        sl = slice(rank * rows, (rank + 1) * rows)
        p = p_all[sl].clone().requires_grad_(True)
        loss = ops.infonce_loss(p, z_all[sl].contiguous(), tau=tau, group=dist.group.WORLD)  # local mean over local queries
        loss.backward()
        pf = p_all.clone().requires_grad_(True)
        full = ops.infonce_loss(pf, z_all, tau=tau, group=False)  # whole problem on this rank, no collective
        full.backward()
        t = loss.detach().clone()
        dist.all_reduce(t)
        out[f"infonce_d{dim}_loss_rel"] = abs(float(t) / world - float(full)) / abs(float(full))
        # d(mean over all rows)/dp_local = d(local mean)/dp_local / world
        a, b = (p.grad.double() / world).flatten(), pf.grad[sl].double().flatten()
        out[f"infonce_d{dim}_grad_cos"] = float((a @ b) / (a.norm() * b.norm()))
        out[f"infonce_d{dim}_grad_rel"] = float((a - b).norm() / b.norm())
        assert out[f"infonce_d{dim}_loss_rel"] <= 1e-5 and out[f"infonce_d{dim}_grad_cos"] >= 0.9999, out
        # keys NOT detached (north_star (4)): msf_infonce_dk partials of every rank reduce-scattered over NCCL vs the key
        # gradient of the gathered problem on one rank (x world: the ranks' local-mean losses add up before DDP averages)
        pk = p_all[sl].clone().requires_grad_(True)
        zk = z_all[sl].clone().requires_grad_(True)
        ops.infonce_loss(pk, zk, tau=tau, group=dist.group.WORLD, detach_keys=False).backward()
        pf2, zf2 = p_all.clone().requires_grad_(True), z_all.clone().requires_grad_(True)
        ops.infonce_loss(pf2, zf2, tau=tau, group=False, detach_keys=False).backward()
        a, b = (zk.grad.double() / world).flatten(), zf2.grad[sl].double().flatten()
        out[f"infonce_d{dim}_keygrad_cos"] = float((a @ b) / (a.norm() * b.norm()))
        out[f"infonce_d{dim}_keygrad_rel"] = float((a - b).norm() / b.norm())
        assert out[f"infonce_d{dim}_keygrad_cos"] >= 0.9999, out
    red = ops.PeerReducer.get(dist.group.WORLD, dev) if ops.USE_PEER_ALLREDUCE else None
    v = torch.randn(9217, device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(rank))
    want = v.clone()
    dist.all_reduce(want)
    if red is not None:
        got = red.all_reduce_(v.clone())
        out["peer_allreduce_max_abs"] = float((got - want).abs().max())
        assert out["peer_allreduce_max_abs"] <= 1e-12 * world, out
    else:
        out["peer_allreduce_max_abs"] = None  # symmetric memory unavailable: statistics go through NCCL
    bn = M.module.FusedBatchNorm1d(256, act="relu").to(dev).train()
    x_all = torch.randn(world * 64, 256, device=dev, generator=g) * 2 + 0.5
    x = x_all[rank * 64:(rank + 1) * 64].clone().requires_grad_(True)
    y = bn(x)
    (y * y).sum().backward()
    xa = x_all.clone().double().requires_grad_(True)
    ya = torch.relu(torch.nn.functional.batch_norm(xa, None, None, None, None, True, 0.1, 1e-5))
    (ya * ya).sum().backward()
    out["syncbn_out_max_abs"] = float((y.detach().double() - ya.detach()[rank * 64:(rank + 1) * 64]).abs().max())
    out["syncbn_dx_rel"] = float((x.grad.double() - xa.grad[rank * 64:(rank + 1) * 64]).norm() / xa.grad[rank * 64:(rank + 1) * 64].norm())
    out["syncbn_running_var_rel"] = float(((bn.running_var.double() - (0.9 + 0.1 * x_all.double().var(0, unbiased=True))).abs() /
                                            (0.9 + 0.1 * x_all.double().var(0, unbiased=True))).max())
    assert out["syncbn_out_max_abs"] <= 2e-5 and out["syncbn_dx_rel"] <= 1e-4 and out["syncbn_running_var_rel"] <= 1e-5, out
    worst = torch.tensor([v if v is not None and "cos" not in k else 0.0 for k, v in out.items()], dtype=torch.float64, device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)  # worst rank
    keys = list(out)
    for i, k in enumerate(keys):
        if out[k] is not None and "cos" not in k:
            out[k] = float(worst[i])
    out["ranks"] = world
    return out


# ------------------------------------------------------------------------------------------------------------
# baseline: the reference's module graph on the same GPU(s), PyTorch eager (cuBLAS / cuDNN / ATen / NCCL)
# ------------------------------------------------------------------------------------------------------------
def gpu_eager_baseline(torch, dist, args, dev, world, rank, barrier, max_over_ranks, value, e2e):
    from oracle.cpu_step import ReferenceStep  # baseline leg only
    steps, warm = max(2, min(args.steps, 10)), 3
    batch = args.batch
    note = None
    while True:
        try:
            ref = ReferenceStep(batch, img=args.img, loss=args.loss, tau=args.tau, seed=3407 + rank, device=str(dev), autocast_dtype=torch.bfloat16,
                                lr=1e-3 * (batch * world) ** 0.5 / 32 ** 0.5)
            for _ in range(warm):
                ref.step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                ref.step()
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            break
        except torch.OutOfMemoryError:
            ref = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            if world > 1 or batch <= 16:  # ranks must agree on the batch: no per-rank fallback under DDP
                return {"unavailable": f"out of memory at batch {batch}/GPU"}
            batch //= 2
            note = f"out of memory at batch {args.batch}/GPU: measured at {batch}/GPU"
    v = batch * world * steps / (ms / 1e3)
    out = {"value": v, "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warm, "per_gpu_batch": batch, "loss_mode": args.loss,
           "loss": float(ref.last.item()),
           "what": "the reference's module graph with stock torch modules (oracle/torch_ref.py: nn.Conv2d/BatchNorm2d ResNet-18 encoders, "
                   "nn.Linear/nn.BatchNorm1d heads, advanced-index gather + cat, 24 nn.CosineSimilarity calls, torch.optim.Adam; "
                   "SyncBatchNorm + DDP at N > 1), bf16 autocast, inputs resident, as tools/ssl_train.py:441-474 runs it",
           "speedup_value_over_eager": value / v, "speedup_e2e_over_eager": e2e / v}
    if note:
        out["note"] = note
    return out


# ------------------------------------------------------------------------------------------------------------
# microbenchmarks measured in the same process
# ------------------------------------------------------------------------------------------------------------
def hot_path_graph_bench(torch, M, _lib, dev, args):
    """The hot path WITHOUT the encoders (pooled pyramid features in: A1 gather/concat, the grouped head stage, the loss, their
    backward and FusedAdam -- what SURVEY section 8 scopes) at the step's batch: eager against ONE CUDA graph per step
    (msfwsi_b200.GraphedStep).  CUDA events, 3 warm-ups; the same numbers as `bench.py --heads-only [--cuda-graph]`."""
    B, K = args.batch, 16

    class _NoEncoder(torch.nn.Module):
        def __init__(self, **_):
            super().__init__()
            self.fc = torch.nn.Identity()

    model = M.MSFWSI(lambda **kw: _NoEncoder(**kw), 4, 2048, 512, 0.5, False).to(dev).train()
    groups = [{"params": [p for n, p in model.named_parameters() if n.startswith(pre)]} for pre in ("context_", "target_", "inter_")]
    opt = M.FusedAdam([g for g in groups if g["params"]], lr=1e-3)
    M.bind_optimizer(model, opt)
    g = torch.Generator().manual_seed(3407)
    feats = lambda n: [torch.randn(n, d, generator=g).abs().to(torch.bfloat16).to(dev) for d in (64, 128, 256, 512)]
    inputs = {"c1": feats(B), "c2": feats(B), "t1": feats(B * K), "t2": feats(B * K),
              "r": [torch.stack([torch.randperm(K, generator=g).argsort() for _ in range(B)]).to(dev) for _ in range(2)]}

    def step(d):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.heads_loss(d["c1"], d["c2"], d["t1"], d["t2"], d["r"], FUSER_WEIGHTS, mode=args.loss, tau=args.tau)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def timeit(fn, n):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    l0 = _lib.launch_count
    ms_eager = timeit(lambda: step(inputs), 10)
    per_step = (_lib.launch_count - l0) // 13
    graphed = M.GraphedStep(step, inputs, opt, warmup=1)
    ms_graph = timeit(graphed, 30)
    return {"what": f"hot path only (no encoders), batch {B}, loss={args.loss}: gather/concat + 12 projectors + 12 predictors + loss, forward + backward + Adam",
            "eager": {"ms_per_step": ms_eager, "tiles_per_sec": B / ms_eager * 1e3},
            "cuda_graph": {"ms_per_step": ms_graph, "tiles_per_sec": B / ms_graph * 1e3, "launches_per_step": graphed.launches_per_replay},
            "speedup_graph_over_eager": ms_eager / ms_graph, "eager_launches_per_step": per_step,
            "note": "the eager step is Python / autograd / launch bound; the replayed graph is bit-identical (tests/test_cuda_graph_gpu.py)"}


def infonce_microbench(torch, ops, _lib, dev, peaks, traffic):
    """c5 rows through the GROUPED entry points the training step calls (msf_nce_grouped_fwd / _bwd, one pair): the fused
    InfoNCE forward (flash tcgen05 launch + finalize + sum) and the whole fwd+bwd chain (row-normalise x2, forward, backward)
    with CUDA events on the launching stream, L2 flushed between iterations."""
    L = _lib
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = peaks.get("bf16_tflops", 1590.0)
    rows = []
    for n, d in ((16384, 128), (65536, 128), (65536, 256)):
        g = torch.Generator(device=dev).manual_seed(3407)
        k = torch.randn(n, d, device=dev, generator=g)
        q = (0.3 * k + torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
        k = k.to(torch.bfloat16)
        qh, _ = ops.rownorm(q, torch.bfloat16)
        kh, _ = ops.rownorm(k, torch.bfloat16)
        gq = torch.empty_like(q)
        pair = (L.NcePair * 1)(L.NcePair(qh.data_ptr(), 0, kh.data_ptr(), gq.data_ptr(), 0, n, n, 1, d, 0, 1.0))
        wsb = L.lib().msf_nce_grouped_workspace_bytes(pair, 1)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        loss, gout = torch.empty((), device=dev), torch.ones((), device=dev)
        st = L.stream_ptr()

        def fwd():
            L.check(L.lib().msf_nce_grouped_fwd(pair, 1, L.MSF_BF16, 0.07, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, st), "fwd")

        def chain():
            a, _ = ops.rownorm(q, torch.bfloat16)
            b, _ = ops.rownorm(k, torch.bfloat16)
            pr = (L.NcePair * 1)(L.NcePair(a.data_ptr(), 0, b.data_ptr(), gq.data_ptr(), 0, n, n, 1, d, 0, 1.0))
            L.check(L.lib().msf_nce_grouped_fwd(pr, 1, L.MSF_BF16, 0.07, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
            L.check(L.lib().msf_nce_grouped_bwd(pr, 1, L.MSF_BF16, 0.07, 1e-8, gout.data_ptr(), ws.data_ptr(), wsb, st), "bwd")

        out = {"N": n, "D": d, "flops": 4.0 * n * n * d}
        # the two measurements alternate, so both see the same clock / power state (a dense-MMA kernel repeated for a second
        # drops from the burst to the sustained clock: timing one after the other would compare different GPUs)
        for _ in range(3):
            fwd()
            chain()
        ts = {"fwd": [], "fwd_bwd": []}
        for _ in range(7):
            for name, fn in (("fwd", fwd), ("fwd_bwd", chain)):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts[name].append(a.elapsed_time(b))
        for name in ts:
            out[f"ms_{name}"] = statistics.median(ts[name])
            out[f"tflops_{name}"] = out["flops"] / out[f"ms_{name}"] / 1e9
            out[f"frac_{name}"] = out[f"tflops_{name}"] / peak
        rows.append(out)
    top = rows[-1]  # N = 65536, D = 256: the configuration the ncu traffic capture was taken on
    t = traffic.get("infonce_flash_fwd")
    roof = {"kernel": f"infonce_grouped_kernel<{top['D']}> N=Nq={top['N']} tau=0.07: forward+backward chain (row-normalise x2, fused tcgen05 "
                      "main loop, finalize, backward)",
            "bound": "tensor", "achieved": top["tflops_fwd_bwd"], "peak": peak, "unit": "TFLOP/s", "frac": top["frac_fwd_bwd"],
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone)" if peaks else "fallback 1590 (B200_PROFILING.md)",
            "algorithmic_flops_per_launch": top["flops"], "ms": top["ms_fwd_bwd"],
            "main_kernel_only": {"achieved": top["tflops_fwd"], "frac": top["frac_fwd"], "ms": top["ms_fwd"]},
            "d128": {"achieved": rows[1]["tflops_fwd_bwd"], "frac": rows[1]["frac_fwd_bwd"], "ms": rows[1]["ms_fwd_bwd"],
                     "main_kernel_only_frac": rows[1]["frac_fwd"]},
            "traffic": None if t is None else t.get("dram_bytes_per_launch"),
            "traffic_note": None if t is None else (f"ncu capture at {t.get('shape')}: {t.get('dram_bytes_per_launch'):.4g} B of DRAM traffic for "
                                                    f"{t.get('algorithmic_bytes_per_launch'):.4g} algorithmic B (ratio {t.get('ratio_to_algorithmic'):.3f}); "
                                                    f"{t.get('source')}"),
            "timing": "torch.cuda.Event pair on the current stream = the stream the C-ABI call launches on; median of 7 after 3 warm-ups, "
                      "256 MB L2 flush between iterations; forward-only and chain iterations alternate (same clock state)"}
    return {"rows": rows, "roofline": roof}


def hbm_microbench(torch, ops, _lib, dev, peaks):
    """A1 (gather/concat at the c4 size), A2 (crop-resample: 4x bilinear zoom and the integer copy) and E1 (EMA over the
    reference's 123.6 M parameters) against the measured copy bandwidth; working sets beyond the 126 MB L2."""
    L = _lib
    peak = peaks.get("hbm_gbs", 6650.0)
    rows = []

    def timeit(fn, iters=7, reps=4):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        return statistics.median(ts)

    def rec(name, nbytes, ms, note):
        rows.append({"kernel": name, "bound": "hbm", "bytes": nbytes, "ms": ms, "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": nbytes / ms / 1e6 / peak, "note": note})

    B, K, dims, e = 1024, 16, (64, 128, 256, 512), 2  # c4: 1024 tiles per GPU
    g = torch.Generator(device=dev).manual_seed(1)
    ctx = [torch.randn(B, d, device=dev, generator=g).to(torch.bfloat16) for _ in range(2) for d in dims]
    tgt = [torch.randn(B * K, d, device=dev, generator=g).to(torch.bfloat16) for _ in range(2) for d in dims]
    rev = [torch.stack([torch.randperm(K) for _ in range(B)]).to(dev)] * 8
    outs = [(torch.empty_like(t), torch.empty((B, 9 * c.shape[1]), dtype=c.dtype, device=dev)) for c, t in zip(ctx, tgt)]
    items = (L.GatherItem * 8)()
    for i in range(8):
        items[i] = L.GatherItem(tgt[i].data_ptr(), ctx[i].data_ptr(), rev[i].data_ptr(), outs[i][0].data_ptr(), outs[i][1].data_ptr(), ctx[i].shape[1], 0)
    st = L.stream_ptr()
    nb = sum((2 * B * K * d + B * d + 9 * B * d) * e for d in dims) * 2 + 2 * B * K * 8
    rec("A1 gather_concat fwd bf16", nb, timeit(lambda: L.check(L.lib().msf_gather_concat_fwd(items, 8, B, K, 8, L.MSF_BF16, None, st), "gather")),
        f"c4 size: B={B}/GPU, 4 levels x 2 views in one launch ({nb / 1e6:.0f} MB; 4x that at B=4096 runs at the same rate)")
    del ctx, tgt, outs
    for (Bc, Cc, H, W, oh, ow, tag) in ((64, 128, 128, 128, 128, 128, "4x bilinear zoom"), (256, 128, 128, 128, 32, 32, "integer copy (hooknet.py:29-32 case)")):
        feat = torch.randn(Bc, Cc, H, W, device=dev, generator=g).to(torch.bfloat16)
        boxes = ops.footprint_boxes(Bc, 4, H, W, dev)
        outp = torch.empty((Bc, 16, Cc, oh, ow), dtype=torch.bfloat16, device=dev)
        nb = feat.numel() * e + outp.numel() * e + Bc * 16 * 16
        fn = lambda: L.check(L.lib().msf_crop_resample_fwd(feat.data_ptr(), Bc, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.MSF_BF16, outp.data_ptr(), st), "crop")
        rec(f"A2 crop_resample fwd bf16, {tag}", nb, timeit(fn), f"feat {tuple(feat.shape)} -> 16 footprints of {oh}x{ow}; {100 * outp.numel() * e // nb}% of the bytes are writes")
        if tag.startswith("4x"):  # DRAM traffic of the same launch from the ncu capture (profiles/ncu_traffic.json)
            try:
                t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("crop_resample_fwd")
                if t and t.get("algorithmic_bytes_per_launch") == nb:
                    rows[-1]["traffic"] = t["dram_bytes_per_launch"]
            except Exception:
                pass
        gfeat = torch.empty((Bc, Cc, H, W), dtype=torch.float32, device=dev)
        nbb = outp.numel() * e + gfeat.numel() * 4 + Bc * 16 * 16
        fnb = lambda: L.check(L.lib().msf_crop_resample_bwd(outp.data_ptr(), Bc, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.MSF_BF16, gfeat.data_ptr(), st), "crop bwd")
        rec(f"A2 crop_resample bwd bf16 (gather form, no atomics), {tag}", nbb, timeit(fnb), f"grad_out {tuple(outp.shape)} -> fp32 grad_feat {tuple(gfeat.shape)}")
        del feat, outp, gfeat
    # D1b: the 34 views of 64 source tiles (context crops of the whole tile, target crops inside the 256^2 tiles), stem layout out
    Bs, S = 64, 1024
    src = torch.randint(0, 256, (Bs, S, S, 3), dtype=torch.uint8, device=dev, generator=g)
    cr = []
    for v in range(2):
        side = torch.randint(724, 1025, (Bs,), device=dev, generator=g)
        y0 = (torch.rand(Bs, device=dev, generator=g) * (S - side + 1)).long()
        x0 = (torch.rand(Bs, device=dev, generator=g) * (S - side + 1)).long()
        cr.append(torch.stack((torch.arange(Bs, device=dev), y0, x0, y0 + side, x0 + side, torch.zeros_like(y0)), 1))
    for v in range(2):
        side = torch.randint(181, 257, (Bs, 16), device=dev, generator=g)
        y0 = (torch.rand(Bs, 16, device=dev, generator=g) * (256 - side + 1)).long()
        x0 = (torch.rand(Bs, 16, device=dev, generator=g) * (256 - side + 1)).long()
        perm = torch.stack([torch.randperm(16, device=dev) for _ in range(Bs)])
        cr.append(ops.jigsaw_view_crops(perm, torch.stack((y0, x0, y0 + side, x0 + side), 2), torch.zeros_like(y0), S, S, 4).long())
    crops = torch.cat(cr).to(torch.int32)
    nv = crops.shape[0]
    area = ((crops[:, 3] - crops[:, 1]).double() * (crops[:, 4] - crops[:, 2]).double()).sum().item()
    outv = torch.empty((nv, 16, 115, 115), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    m3, s3 = (L.C.c_float * 3)(0.6998, 0.4785, 0.6609), (L.C.c_float * 3)(0.2203, 0.2407, 0.1983)
    nb = 3.0 * area + outv.numel() * 2
    fn = lambda: L.check(L.lib().msf_view_crops_s2d(src.data_ptr(), Bs, S, S, crops.data_ptr(), nv, 224, 224, m3, s3, outv.data_ptr(), L.MSF_BF16, 0, st), "views")
    rec("D1b view_crops_s2d uint8 -> bf16 stem layout", nb, timeit(fn),
        f"{nv} views of {Bs} source tiles (RandomResizedCrop boxes, jigsaw permutation); bytes = cropped source regions once ({3.0 * area / 1e6:.0f} MB) + "
        f"views written ({outv.numel() * 2 / 1e6:.0f} MB)")
    del src, outv
    sizes = [64 * 3 * 49, 64, 64] + [64 * 64 * 9] * 4 + [128 * 64 * 9, 128 * 128 * 9] + [256 * 256 * 9] * 3 + [512 * 512 * 9] * 3 + \
            [4608 * 4608] * 3 + [2304 * 2304] * 3 + [1152 * 1152] * 3 + [576 * 576] * 3 + [512, 256, 128, 64] * 8
    teacher = [torch.randn(n, device=dev) for n in sizes]
    student = [torch.randn(n, device=dev) for n in sizes]
    up = ops.EmaUpdater(teacher, student)
    rec("E1 ema_multi fp32", 12 * up.numel, timeit(lambda: up.step(0.996)), f"{len(sizes)} tensors, {up.numel / 1e6:.1f} M parameters, one launch")
    return rows


def main():
    args = parse()
    # Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) get stderr instead.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
