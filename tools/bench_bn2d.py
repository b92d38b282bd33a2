"""HBM-roofline microbench of the channels-last BatchNorm2d kernels (N1) at the ResNet-18 activation shapes of a
4096-image target batch, each kernel called through the C ABI and timed alone with CUDA events (3 warm-ups, median
of 10, tensors >> the 126 MB L2).  achieved = algorithmic bytes / time; peak = MEASURED_PEAKS.json hbm_gbs.
ATen's kernels for the same work are timed next to them."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L  # noqa: E402

dev = "cuda:0"


ONCE = False


def timeit(fn, iters=10):
    if ONCE:  # ncu capture mode: one launch per kernel
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--aten", action="store_true", help="also time ATen's kernels for the same work")
    ap.add_argument("--once", action="store_true", help="launch every kernel exactly once (for an ncu capture)")
    ap.add_argument("--stem-n", type=int, default=2048)
    ap.add_argument("--only", default=None, help="comma-separated shape tags (stem,layer1,...)")
    args = ap.parse_args()
    global ONCE
    ONCE = args.once
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    lib, st = L.lib(), L.stream_ptr()
    rows = []

    def rec(name, nbytes, ms, note=""):
        r = {"kernel": name, "bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "note": note}
        rows.append(r)
        print(json.dumps(r), flush=True)

    dt, code, e = torch.bfloat16, L.MSF_BF16, 2
    for (N, C, H, W, tag) in ((args.stem_n, 64, 112, 112, "stem"), (4096, 64, 56, 56, "layer1"), (4096, 128, 28, 28, "layer2"),
                              (4096, 256, 14, 14, "layer3"), (4096, 512, 7, 7, "layer4")):
        if args.only and tag not in args.only.split(","):
            continue
        x = torch.randn(N, H, W, C, device=dev).to(dt)
        dy = torch.randn(N, H, W, C, device=dev).to(dt)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        rows_ = N * H * W
        nb = x.numel() * e
        wsb = lib.msf_bn2d_workspace_bytes(rows_, C)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        sums, sums2 = torch.empty(2 * C + 1, dtype=torch.float64, device=dev), torch.empty(2 * C, dtype=torch.float64, device=dev)
        mean, invstd = torch.empty(C, device=dev), torch.empty(C, device=dev)
        gamma, beta = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) - 0.5
        cnt = sums.data_ptr() + 16 * C
        shape = f"{tag} ({N},{C},{H},{W}) bf16 NHWC"
        f_stats = lambda: L.check(lib.msf_bn2d_stats(x.data_ptr(), rows_, C, code, sums.data_ptr(), ws.data_ptr(), wsb, st), "stats")
        rec("bn_stats", nb, timeit(f_stats), shape)
        L.check(lib.msf_bn2d_finalize(sums.data_ptr(), C, 1e-5, 0.1, mean.data_ptr(), invstd.data_ptr(), None, None, st), "fin")
        f_apply = lambda: L.check(lib.msf_bn2d_apply(x.data_ptr(), None, y.data_ptr(), None, rows_, C, code, mean.data_ptr(), invstd.data_ptr(),
                                                     gamma.data_ptr(), beta.data_ptr(), 1, st), "apply")
        rec("bn_apply+relu", 2 * nb, timeit(f_apply), shape)
        f_red = lambda: L.check(lib.msf_bn2d_bwd_reduce(x.data_ptr(), dy.data_ptr(), None, 0, rows_, C, code, mean.data_ptr(), invstd.data_ptr(),
                                                        gamma.data_ptr(), beta.data_ptr(), 1, None, 0, sums2.data_ptr(), ws.data_ptr(), wsb, st), "red")
        rec("bn_bwd_reduce (relu mask recomputed)", 2 * nb, timeit(f_red), shape)
        f_el = lambda: L.check(lib.msf_bn2d_bwd_elemt(x.data_ptr(), dy.data_ptr(), None, 0, dx.data_ptr(), None, rows_, C, code, mean.data_ptr(),
                                                      invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, None, 0, sums2.data_ptr(), cnt, st), "el")
        rec("bn_bwd_elemt (relu mask recomputed)", 3 * nb, timeit(f_el), shape)
        if tag != "stem":
            res, dres = torch.randn_like(x), torch.empty_like(x)
            bits = torch.empty(x.numel() // 8, dtype=torch.uint8, device=dev)
            f_apply_r = lambda: L.check(lib.msf_bn2d_apply(x.data_ptr(), res.data_ptr(), y.data_ptr(), bits.data_ptr(), rows_, C, code, mean.data_ptr(),
                                                           invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, st), "apply")
            rec("bn_apply+residual+relu (+mask bits)", 3 * nb + nb // 16, timeit(f_apply_r), shape)
            f_red_r = lambda: L.check(lib.msf_bn2d_bwd_reduce(x.data_ptr(), dy.data_ptr(), bits.data_ptr(), 1, rows_, C, code, mean.data_ptr(),
                                                              invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, None, 0, sums2.data_ptr(), ws.data_ptr(),
                                                              wsb, st), "red")
            rec("bn_bwd_reduce (mask bits)", 2 * nb + nb // 16, timeit(f_red_r), shape)
            f_el_r = lambda: L.check(lib.msf_bn2d_bwd_elemt(x.data_ptr(), dy.data_ptr(), bits.data_ptr(), 1, dx.data_ptr(), dres.data_ptr(), rows_, C, code,
                                                            mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, None, 0, sums2.data_ptr(),
                                                            cnt, st), "el")
            rec("bn_bwd_elemt (mask bits, + dres)", 4 * nb + nb // 16, timeit(f_el_r), shape)
            del res, dres, bits
        else:
            PH, PW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            yp = torch.empty(N, PH, PW, C, device=dev, dtype=dt)
            tap = torch.empty(N, PH, PW, C, device=dev, dtype=torch.uint8)
            dp = torch.randn(N, PH, PW, C, device=dev).to(dt)
            npool = yp.numel()
            wsb2 = lib.msf_bn2d_workspace_bytes(N * PH * PW, C)
            ws2 = torch.empty(wsb2, dtype=torch.uint8, device=dev)
            xarg = torch.empty(N, PH, PW, C, device=dev, dtype=dt)
            f_ap = lambda: L.check(lib.msf_bn2d_apply_pool(x.data_ptr(), yp.data_ptr(), tap.data_ptr(), xarg.data_ptr(), N, H, W, C, code,
                                                           mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), st), "apply_pool")
            rec("bn_apply+relu+maxpool", nb + npool * (2 * e + 1), timeit(f_ap), shape)
            f_pr = lambda: L.check(lib.msf_bn2d_bwd_reduce(xarg.data_ptr(), dp.data_ptr(), yp.data_ptr(), 0, N * PH * PW, C, code, mean.data_ptr(),
                                                           invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1, None, 0, sums2.data_ptr(), ws2.data_ptr(),
                                                           wsb2, st), "pool_red")
            rec("bn_bwd_reduce on the pooled grid (x_arg, dpool, y)", 3 * npool * e, timeit(f_pr), shape)
            f_pe = lambda: L.check(lib.msf_bn2d_pool_bwd_elemt(x.data_ptr(), dp.data_ptr(), tap.data_ptr(), dx.data_ptr(), N, H, W, C, code,
                                                               mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), sums2.data_ptr(), cnt, st), "pool_el")
            rec("bn_pool_bwd_elemt", 2 * nb + npool * (e + 1), timeit(f_pe), shape)
            del yp, tap, dp, xarg
        if args.aten:
            xa = x.permute(0, 3, 1, 2)  # NCHW view of the NHWC buffer = channels_last
            dya = dy.permute(0, 3, 1, 2)
            if xa.numel() < 2 ** 31:
                rec("ATen batch_norm fwd (stats + apply)", 3 * nb, timeit(lambda: F.batch_norm(xa, None, None, gamma, beta, True, 0.1, 1e-5)), shape)
                xg = xa.detach().requires_grad_(True)
                yg = F.batch_norm(xg, None, None, gamma, beta, True, 0.1, 1e-5)
                rec("ATen batch_norm bwd (reduce + elemt)", 5 * nb, timeit(lambda: torch.autograd.grad(yg, xg, dya, retain_graph=True)), shape)
                del xg, yg
                ya = y.permute(0, 3, 1, 2)
                rec("ATen relu_", 2 * nb, timeit(lambda: ya.relu_()), shape)
                if tag == "stem":
                    rec("ATen max_pool2d fwd", nb + nb // 4 * 5, timeit(lambda: F.max_pool2d(xa, 3, 2, 1)), shape + " (+int64 indices)")
        del x, dy, y, dx
    if not args.only or "extra" in args.only.split(","):
        # S1: space-to-depth layout of a 4096-image (or --stem-n) bf16 NHWC batch of 224^2 views
        from msfwsi_b200 import FusedAdam, ops
        n_img = args.stem_n
        img = torch.randn(n_img, 3, 224, 224, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        out_bytes = n_img * 115 * 115 * 16 * e
        rec("stem_s2d", img.numel() * e + out_bytes, timeit(lambda: ops.stem_s2d(img, dt)), f"({n_img},3,224,224) bf16 NHWC -> ({n_img},16,115,115)")
        del img
        # O1: Adam over the reference model's 123.6 M fp32 parameters (two ResNet-18 encoders + heads), three lr groups
        sizes = [64 * 3 * 49, 64, 64] + [64 * 64 * 9] * 4 + [128 * 64 * 9, 128 * 128 * 9 * 3] + [256 * 256 * 9] * 3 + [512 * 512 * 9] * 3 + \
                [4608 * 4608] * 3 + [2304 * 2304] * 3 + [1152 * 1152] * 3 + [576 * 576] * 3 + [512, 256, 128, 64] * 8
        ps = [torch.randn(n, device=dev).requires_grad_(True) for n in sizes]
        for p in ps:
            p.grad = torch.randn_like(p)
        third = len(ps) // 3
        opt = FusedAdam([{"params": ps[:third]}, {"params": ps[third:2 * third], "lr": 3e-3}, {"params": ps[2 * third:]}], lr=1e-3)
        rec("adam_multi", 28 * sum(sizes), timeit(opt.step), f"{len(sizes)} fp32 tensors, {sum(sizes) / 1e6:.1f} M parameters, 3 lr groups")
    if args.out:
        json.dump({"peak_GBps": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
