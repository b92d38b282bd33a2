"""c5 microbench: fused cross-resolution InfoNCE fwd(+bwd) vs the tensor roofline.

    python tools/bench_infonce.py [--n 4096 16384 65536] [--d 128 256] [--iters 20]

Timing: CUDA events on the launching stream, >= 3 warm-ups; inputs for N >= 16k exceed nothing near L2 (126 MB)
only at the largest sizes, so an L2 flush (256 MB memset) runs between timed iterations.  Algorithmic FLOPs =
4*Nq*N*D (keys detached); peak = MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L  # noqa: E402
from msfwsi_b200 import ops  # noqa: E402


def peak_tflops():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"], "measured"
    except Exception:
        return 1590.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1024, 4096, 16384, 65536])
    ap.add_argument("--d", type=int, nargs="+", default=[128, 256])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tau", type=float, default=0.07)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = "cuda:0"
    peak, how = peak_tflops()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for d in args.d:
        for n in args.n:
            g = torch.Generator(device=dev).manual_seed(3407)
            k = torch.randn(n, d, device=dev, generator=g)
            q = (0.3 * k + torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
            k = k.to(torch.bfloat16)
            q_hat, q_inv = ops.rownorm(q, torch.bfloat16)
            k_hat, _ = ops.rownorm(k, torch.bfloat16)
            prec = L.MSF_BF16
            ws_bytes = L.lib().msf_infonce_workspace_bytes(n, n, d, prec)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            gout = torch.ones((), device=dev)
            gq = torch.empty_like(q)
            st = L.stream_ptr()

            def fwd():
                L.check(L.lib().msf_infonce_fwd(q_hat.data_ptr(), k_hat.data_ptr(), n, n, d, 0, args.tau, prec, loss.data_ptr(), 0,
                                                ws.data_ptr(), ws_bytes, st), "fwd")

            def chain():
                qh, qi = ops.rownorm(q, torch.bfloat16)
                kh, _ = ops.rownorm(k, torch.bfloat16)
                L.check(L.lib().msf_infonce_fwd(qh.data_ptr(), kh.data_ptr(), n, n, d, 0, args.tau, prec, loss.data_ptr(), 0,
                                                ws.data_ptr(), ws_bytes, st), "fwd")
                L.check(L.lib().msf_infonce_bwd(qh.data_ptr(), kh.data_ptr(), qi.data_ptr(), n, n, d, 0, args.tau, prec, gout.data_ptr(),
                                                1.0 / n, ws.data_ptr(), ws_bytes, gq.data_ptr(), L.MSF_BF16, st), "bwd")

            # the grouped entry points (what the training step calls; one pair here): same flash pipeline behind a problem table
            pair = (L.NcePair * 1)(L.NcePair(q_hat.data_ptr(), 0, k_hat.data_ptr(), gq.data_ptr(), 0, n, n, 1, d, 0, 1.0))
            gws_bytes = L.lib().msf_nce_grouped_workspace_bytes(pair, 1)
            gws = torch.empty(gws_bytes, dtype=torch.uint8, device=dev)
            gloss = torch.empty((), dtype=torch.float32, device=dev)

            def gfwd():
                L.check(L.lib().msf_nce_grouped_fwd(pair, 1, prec, args.tau, 1e-8, gloss.data_ptr(), gws.data_ptr(), gws_bytes, st), "gfwd")

            def gchain():
                qh, qi = ops.rownorm(q, torch.bfloat16)
                kh, _ = ops.rownorm(k, torch.bfloat16)
                pr = (L.NcePair * 1)(L.NcePair(qh.data_ptr(), 0, kh.data_ptr(), gq.data_ptr(), 0, n, n, 1, d, 0, 1.0))
                L.check(L.lib().msf_nce_grouped_fwd(pr, 1, prec, args.tau, 1e-8, gloss.data_ptr(), gws.data_ptr(), gws_bytes, st), "gfwd")
                L.check(L.lib().msf_nce_grouped_bwd(pr, 1, prec, args.tau, 1e-8, gout.data_ptr(), gws.data_ptr(), gws_bytes, st), "gbwd")

            res = {}
            for name, fn in (("fwd_kernels", fwd), ("fwd_bwd_chain", chain), ("grouped_fwd", gfwd), ("grouped_fwd_bwd", gchain)):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(args.iters):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ts.sort()
                res[name] = ts[len(ts) // 2]
            flops = 4.0 * n * n * d
            row = {"N": n, "D": d, "loss": float(loss.item()) / n, "ms_fwd": res["fwd_kernels"], "ms_fwd_bwd": res["fwd_bwd_chain"],
                   "tflops_fwd": flops / res["fwd_kernels"] / 1e9, "tflops_fwd_bwd": flops / res["fwd_bwd_chain"] / 1e9,
                   "peak_tflops": peak, "peak_source": how}
            row["frac_fwd"] = row["tflops_fwd"] / peak
            row["frac_fwd_bwd"] = row["tflops_fwd_bwd"] / peak
            row.update({"grouped_ms_fwd": res["grouped_fwd"], "grouped_ms_fwd_bwd": res["grouped_fwd_bwd"], "grouped_frac_fwd": flops / res["grouped_fwd"] / 1e9 / peak,
                        "grouped_frac_fwd_bwd": flops / res["grouped_fwd_bwd"] / 1e9 / peak, "grouped_loss": float(gloss.item()),
                        "exp_mode": os.environ.get("MSF_NCE_EXP_MODE", "default")})
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
