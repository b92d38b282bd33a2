"""HBM-roofline microbench of the memory-bound kernels (A1 gather/concat, L1 cosine loss, A2 crop-resample,
E1 EMA, row-normalise) at sizes larger than the 126 MB L2.  CUDA events on the launching stream, 3 warm-ups,
median of 10; achieved = algorithmic bytes / time; peak = MEASURED_PEAKS.json hbm_gbs."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L  # noqa: E402
from msfwsi_b200 import ops  # noqa: E402

dev = "cuda:0"


def timeit(fn, iters=10, reps=5):
    """Median over `iters` of the per-launch time of `reps` back-to-back launches inside one event pair (so the
    host-side launch path, ~10 us of ctypes per call, is hidden behind the previous launch)."""
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


def abi_gather(ctx, tgt, rev, K, n_keep):
    n = len(ctx)
    B = ctx[0].shape[0]
    outs = [(torch.empty_like(t), torch.empty((B, (n_keep + 1) * c.shape[1]), dtype=c.dtype, device=dev)) for c, t in zip(ctx, tgt)]
    items = (L.GatherItem * n)()
    for i in range(n):
        items[i] = L.GatherItem(tgt[i].data_ptr(), ctx[i].data_ptr(), rev[i].data_ptr(), outs[i][0].data_ptr(), outs[i][1].data_ptr(), ctx[i].shape[1], 0)
    code, st = L.dtype_code(ctx[0].dtype), L.stream_ptr()
    return (lambda: L.check(L.lib().msf_gather_concat_fwd(items, n, B, K, n_keep, code, None, st), "gather")), outs


def abi_cosine(ps, zs, coefs):
    n = len(ps)
    stats = torch.empty((sum(p.shape[0] for p in ps), 4), device=dev)
    grads = [torch.empty_like(p) for p in ps]
    pairs = (L.CosPair * n)()
    off = 0
    for i in range(n):
        pairs[i] = L.CosPair(ps[i].data_ptr(), zs[i].data_ptr(), stats[off:].data_ptr(), grads[i].data_ptr(), ps[i].shape[0], ps[i].shape[1], coefs[i])
        off += ps[i].shape[0]
    wsb = L.lib().msf_cosine_loss_workspace_bytes(pairs, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    loss, gout = torch.empty((), device=dev), torch.ones((), device=dev)
    code, st = L.dtype_code(ps[0].dtype), L.stream_ptr()
    fwd = lambda: L.check(L.lib().msf_cosine_loss_fwd(pairs, n, code, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, st), "cos fwd")
    bwd = lambda: L.check(L.lib().msf_cosine_loss_bwd(pairs, n, code, gout.data_ptr(), st), "cos bwd")
    return fwd, bwd, (stats, grads, ws, loss, gout, pairs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=4096, help="tiles per GPU for the A1 / cosine rows (c4 is 1024)")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    rows = []

    def rec(name, nbytes, ms, note=""):
        r = {"kernel": name, "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "note": note}
        rows.append(r)
        print(json.dumps(r), flush=True)

    B, K, dims = args.batch, 16, (64, 128, 256, 512)
    for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
        ctx = [torch.randn(B, d, device=dev).to(dt) for _ in range(2) for d in dims]
        tgt = [torch.randn(B * K, d, device=dev).to(dt) for _ in range(2) for d in dims]
        rev = [torch.stack([torch.randperm(K) for _ in range(B)]).to(dev)] * 8
        nbytes = sum((2 * B * K * d + B * d + 9 * B * d) * e for d in dims) * 2 + 2 * B * K * 8
        fn, keep = abi_gather(ctx, tgt, rev, K, 8)
        rec(f"A1 gather_concat fwd {dt}", nbytes, timeit(fn), f"B={B} 4 levels x 2 views, one launch")
        del keep
        ps = [t.clone() for t in tgt]
        zs = [torch.randn_like(t) for t in tgt]
        coefs = [-0.5] * len(ps)
        fwd, bwd, keep = abi_cosine(ps, zs, coefs)
        nb_f = sum(2 * t.numel() * e + t.shape[0] * 16 for t in tgt)
        rec(f"L1 cosine fwd {dt}", nb_f, timeit(fwd), f"{len(ps)} pairs, rows={B * K}")
        nb_b = sum(3 * t.numel() * e + t.shape[0] * 16 for t in tgt)
        rec(f"L1 cosine bwd {dt}", nb_b, timeit(bwd), "")
        x = tgt[3]
        xh, inv = torch.empty(x.shape, dtype=torch.bfloat16, device=dev), torch.empty(x.shape[0], device=dev)
        rn = lambda: L.check(L.lib().msf_rownorm(x.data_ptr(), x.shape[0], x.shape[1], L.dtype_code(dt), 1e-8, xh.data_ptr(), L.MSF_BF16,
                                                 inv.data_ptr(), L.stream_ptr()), "rownorm")
        rec(f"rownorm {dt}->bf16", x.numel() * (e + 2) + x.shape[0] * 4, timeit(rn), f"{tuple(x.shape)}")
        del ctx, tgt, ps, zs, keep
    # E1: both ResNet-18 encoders + all heads of the reference model = 123.55 M fp32 parameters
    sizes = [64 * 3 * 49, 64, 64] + [64 * 64 * 9] * 4 + [128 * 64 * 9, 128 * 128 * 9 * 3][0:2] + [256 * 256 * 9] * 3 + [512 * 512 * 9] * 3 + \
            [4608 * 4608] * 3 + [2304 * 2304] * 3 + [1152 * 1152] * 3 + [576 * 576] * 3 + [512, 256, 128, 64] * 8
    teacher = [torch.randn(n, device=dev) for n in sizes]
    student = [torch.randn(n, device=dev) for n in sizes]
    up = ops.EmaUpdater(teacher, student)
    rec("E1 ema fp32", 12 * up.numel, timeit(lambda: up.step(0.996)), f"{len(sizes)} tensors, {up.numel / 1e6:.1f} M params")
    del teacher, student, up
    # A2: low-mag map -> 16 tile footprints, 4x bilinear zoom and the integer (copy) case
    for (Bc, Cc, H, W, oh, ow, tag) in ((64, 128, 128, 128, 128, 128, "4x zoom"), (256, 128, 128, 128, 32, 32, "integer copy")):
        for dt, e in ((torch.bfloat16, 2),):
            feat = torch.randn(Bc, Cc, H, W, device=dev).to(dt)
            boxes = ops.footprint_boxes(Bc, 4, H, W, dev)
            nbytes = feat.numel() * e + Bc * 16 * Cc * oh * ow * e + Bc * 16 * 16
            outp = torch.empty((Bc, 16, Cc, oh, ow), dtype=dt, device=dev)
            cr = lambda: L.check(L.lib().msf_crop_resample_fwd(feat.data_ptr(), Bc, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.dtype_code(dt),
                                                               outp.data_ptr(), L.stream_ptr()), "crop")
            wr = Bc * 16 * Cc * oh * ow * e
            rec(f"A2 crop_resample fwd {dt} {tag}", nbytes, timeit(cr),
                f"feat {tuple(feat.shape)} -> {oh}x{ow}; {100 * wr // nbytes}% of the bytes are writes (write-only probe on this GPU: ~3.96 TB/s, tools/probe_write_bw.py)")
            del feat, outp
    if args.out:
        json.dump({"peak_GBps": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
