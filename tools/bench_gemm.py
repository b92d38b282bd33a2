"""G1 microbench: persistent tcgen05 GEMM vs cuBLAS (torch.matmul) on head-shaped and large problems."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import ops

dev = "cuda:0"


def t(fn, reps=5, iters=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts)


rows = []
for (M, N, K, tag) in ((8192, 8192, 8192, "square"), (16384, 4096, 512, "InfoNCE pass1 shape (Nq x N x D)"), (16384, 512, 4096, "InfoNCE pass2 shape"),
                       (4096, 512, 512, "target projector c2"), (256, 4608, 4608, "inter projector c2"), (16384, 512, 512, "target projector c4"),
                       (1024, 4608, 4608, "inter projector c4")):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    mine = t(lambda: ops.gemm_bf16(A, B, M, N, K))
    ref = t(lambda: torch.matmul(A, B.t()))
    fl = 2.0 * M * N * K
    r = {"M": M, "N": N, "K": K, "what": tag, "tcgen05_ms": mine, "tcgen05_tflops": fl / mine / 1e9, "cublas_ms": ref, "cublas_tflops": fl / ref / 1e9}
    rows.append(r)
    print(json.dumps(r), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
