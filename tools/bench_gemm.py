"""G1 microbench: the grouped persistent tcgen05 GEMM (msf_gemm_grouped, one problem per launch here) and the round-1
kernel (msf_gemm_bf16) vs cuBLAS (torch.matmul) on head-shaped and large problems; plus whole head-stage launches
(24 problems) vs the same 24 cuBLAS calls."""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops

dev = "cuda:0"


def prepared(specs, want_stats=False):
    """The ctypes problem table built once: the timed call is the bare C-ABI launch (no Python-side table building)."""
    arr = (L.GemmProblem * len(specs))()
    keep = []
    for i, g in enumerate(specs):
        cs = torch.empty(((g.M + 127) // 128, 2, g.N), dtype=torch.float32, device=dev) if want_stats else None
        keep.append(cs)
        arr[i] = L.GemmProblem(g.A.data_ptr(), g.A.stride(0), g.B.data_ptr(), g.B.stride(0), g.C.data_ptr(), g.C.stride(0), g.M, g.N, g.K, int(g.a_is_km),
                               int(g.b_is_kn), L.dtype_code(g.C.dtype), 1.0, 0, L.ptr(cs), 0, 0, 0, 0, 0, 0, 0, 0.0, 0, 0)
    wsb = L.lib().msf_gemm_grouped_workspace_bytes(arr, len(specs))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    ctr = ops._counters_for(torch.device(dev))
    st = L.stream_ptr()
    n = len(specs)
    info = (L.C.c_int32 * 6)()
    L.lib().msf_gemm_grouped_plan_info(arr, info)
    def run():
        L.check(L.lib().msf_gemm_grouped(arr, n, L.MSF_BF16, ws.data_ptr(), wsb, ctr.data_ptr(), st), "gemm")
    run.keep = (arr, keep, ws)
    run.plan = list(info)
    return run


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
                       (1024, 4608, 4608, "inter projector c4"), (512, 512, 8192, "target dW c2 (tn, fp32 out)")):
    tn = "dW" in tag
    A = torch.randn((K, M) if tn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if tn else (N, K), device=dev).to(torch.bfloat16)
    C = torch.empty((M, N), dtype=torch.float32 if tn else torch.bfloat16, device=dev)
    spec = ops.GemmSpec(A, B, M, N, K, a_is_km=tn, b_is_kn=tn, out_dtype=C.dtype, C=C)
    run = prepared([spec])
    new = t(run)
    old = t(lambda: ops.gemm_bf16(A, B, M, N, K, a_is_km=tn, b_is_kn=tn, out_dtype=C.dtype))
    ref = t((lambda: torch.matmul(A.t(), B)) if tn else (lambda: torch.matmul(A, B.t())))
    fl = 2.0 * M * N * K
    r = {"M": M, "N": N, "K": K, "what": tag, "grouped_ms": new, "grouped_tflops": fl / new / 1e9, "r1_kernel_ms": old, "cublas_ms": ref,
         "cublas_tflops": fl / ref / 1e9, "grouped_over_cublas": ref / new, "plan_tile_n_tm_tn_ks_kbps_kb": run.plan}
    rows.append(r)
    print(json.dumps(r), flush=True)

# whole head stage: first Linear of 12 heads x 2 views (c2: B = 256) in one launch vs 24 cuBLAS calls
for Bt in (256, 1024):
    specs, pairs = [], []
    for rows_, mult in ((Bt, 1), (Bt * 16, 1), (Bt, 9)):
        for d in (64, 128, 256, 512):
            dim = d * mult
            W = (torch.randn(dim, dim, device=dev) / dim ** 0.5).to(torch.bfloat16)
            for v in range(2):
                X = torch.randn(rows_, dim, device=dev).abs().to(torch.bfloat16)
                specs.append(ops.GemmSpec(X, W, rows_, dim, dim, C=torch.empty((rows_, dim), dtype=torch.bfloat16, device=dev)))
                pairs.append((X, W))
    fl = sum(2.0 * g.M * g.N * g.K for g in specs)
    new = t(prepared(specs, want_stats=True))
    ref = t(lambda: [torch.matmul(X, W.t()) for X, W in pairs])
    r = {"what": f"head stage depth 1, 24 problems, B={Bt}: ONE grouped launch (+ BN statistics epilogue) vs 24 cuBLAS calls", "grouped_ms": new,
         "cublas_ms": ref, "gflop": fl / 1e9, "grouped_tflops": fl / new / 1e9, "grouped_over_cublas": ref / new}
    rows.append(r)
    print(json.dumps(r), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
