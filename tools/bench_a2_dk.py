"""Microbench of (a) the A2 crop-resample forward strip kernel (4x zoom / 2x zoom / integer copy; fp32 and bf16) and
(b) the InfoNCE key-gradient pass msf_infonce_dk (north_star (4)) beside the forward chain.  CUDA events on the launching
stream, 3 warm-ups, median of 10; L2 flushed between the tensor-bound iterations.  Prints one JSON object."""
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


def timeit(fn, iters=10, reps=5, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("--skip-dk", action="store_true")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6465.2))
    tf_burst = float(peaks.get("bf16_tflops_burst", peaks.get("bf16_tflops", 1651.5)))
    res = {"hbm_peak_GBps": hbm, "bf16_peak_TFLOPs": tf_burst, "crop_minb": os.environ.get("MSF_CROP_STRIP_MINB", "4"), "a2": [], "dk": []}
    for (Bc, Cc, H, W, oh, ow, tag) in ((64, 128, 128, 128, 128, 128, "4x zoom"), (128, 128, 128, 128, 64, 64, "2x zoom"),
                                         (256, 128, 128, 128, 32, 32, "integer copy")):
        for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
            if dt == torch.float32 and tag != "4x zoom":
                continue
            Bq = Bc if e == 2 else Bc // 2
            feat = torch.randn(Bq, Cc, H, W, device=dev).to(dt)
            boxes = ops.footprint_boxes(Bq, 4, H, W, dev)
            nbytes = feat.numel() * e + Bq * 16 * Cc * oh * ow * e + Bq * 16 * 16
            outp = torch.empty((Bq, 16, Cc, oh, ow), dtype=dt, device=dev)
            cr = lambda: L.check(L.lib().msf_crop_resample_fwd(feat.data_ptr(), Bq, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.dtype_code(dt),
                                                               outp.data_ptr(), L.stream_ptr()), "crop")
            ms = timeit(cr)
            res["a2"].append({"case": f"fwd {tag} {str(dt).split('.')[-1]}", "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / hbm})
            gfeat = torch.empty((Bq, Cc, H, W), dtype=torch.float32, device=dev)
            bw = lambda: L.check(L.lib().msf_crop_resample_bwd(outp.data_ptr(), Bq, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.dtype_code(dt),
                                                               gfeat.data_ptr(), L.stream_ptr()), "crop bwd")
            nb = Bq * 16 * Cc * oh * ow * e + gfeat.numel() * 4 + Bq * 16 * 16
            ms = timeit(bw)
            res["a2"].append({"case": f"bwd {tag} {str(dt).split('.')[-1]}", "bytes": nb, "ms": ms, "GBps": nb / ms / 1e6, "frac": nb / ms / 1e6 / hbm,
                              "form": "gather"})
            del gfeat
            del feat, outp
    if not args.skip_dk:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        g = torch.ones((), device=dev)
        for (n, dim) in ((16384, 128), (65536, 128), (65536, 256), (16384, 512)):
            gen = torch.Generator(device=dev).manual_seed(3407)
            q = torch.randn(n, dim, device=dev, generator=gen)
            k = q + 0.5 * torch.randn(n, dim, device=dev, generator=gen)
            q_hat, q_inv = ops.rownorm(q.to(torch.bfloat16), torch.bfloat16)
            k_hat, k_inv = ops.rownorm(k.to(torch.bfloat16), torch.bfloat16)
            prec = L.MSF_BF16
            wsb = L.lib().msf_infonce_workspace_bytes(n, n, dim, prec)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            dwb = L.lib().msf_infonce_dk_workspace_bytes(n, n, dim, prec)
            dws = torch.empty(dwb, dtype=torch.uint8, device=dev)
            loss = torch.empty((), device=dev)
            dk = torch.empty((n, dim), device=dev)
            gz = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
            gq = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
            st = L.stream_ptr()
            fwd = lambda: L.check(L.lib().msf_infonce_fwd(q_hat.data_ptr(), k_hat.data_ptr(), n, n, dim, 0, 0.07, prec, loss.data_ptr(), 0,
                                                          ws.data_ptr(), wsb, st), "fwd")
            bwd = lambda: L.check(L.lib().msf_infonce_bwd(q_hat.data_ptr(), k_hat.data_ptr(), q_inv.data_ptr(), n, n, dim, 0, 0.07, prec, g.data_ptr(),
                                                          1.0 / n, ws.data_ptr(), wsb, gq.data_ptr(), L.MSF_BF16, st), "bwd")
            def dkf():
                L.check(L.lib().msf_infonce_dk(q_hat.data_ptr(), k_hat.data_ptr(), n, n, dim, 0, 0.07, prec, g.data_ptr(), 1.0 / n, ws.data_ptr(), wsb,
                                               dk.data_ptr(), dws.data_ptr(), dwb, st), "dk")
                L.check(L.lib().msf_infonce_dk_finish(dk.data_ptr(), q_hat.data_ptr(), k_hat.data_ptr(), k_inv.data_ptr(), n, n, dim, 0.07, prec,
                                                      g.data_ptr(), 1.0 / n, gz.data_ptr(), L.MSF_BF16, st), "finish")
            fwd()
            t_f = timeit(fwd, iters=5, reps=1, flush=flush)
            t_b = timeit(bwd, iters=5, reps=1, flush=flush)
            t_k = timeit(dkf, iters=5, reps=1, flush=flush)
            fl = 4.0 * n * n * dim
            res["dk"].append({"n": n, "dim": dim, "fwd_ms": t_f, "bwd_q_ms": t_b, "key_grad_ms": t_k,
                              "fwd_frac": fl / t_f / 1e9 / tf_burst, "key_grad_frac_4NND": fl / t_k / 1e9 / tf_burst,
                              "full_chain_frac_6NND_algorithmic": 6.0 * n * n * dim / (t_f + t_b + t_k) / 1e9 / tf_burst})
            del q, k, q_hat, k_hat, ws, dws, dk, gz, gq
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
