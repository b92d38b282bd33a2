"""Diag: forced (tile_n, split_k) sweep of the grouped GEMM on one shape -- where the launch planner's cost model is right or wrong."""
import os, sys, statistics, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from msfwsi_b200 import ops
dev = "cuda:0"
def timeit(fn, iters=7, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / reps * 1e3)
    return statistics.median(ts)
shapes = [tuple(int(v) for v in s.split("x")) for s in os.environ.get("SHAPES", "256x4608x4608,256x2304x2304,4608x4608x512").split(",")]
for (M, N, K) in shapes:
    tn = M == N  # the dW shape: A stored [K][M], B stored [K][N], fp32 out
    A = torch.randn((K, M) if tn else (M, K), device=dev).to(torch.bfloat16)
    B = torch.randn((K, N) if tn else (N, K), device=dev).to(torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for bn in (64, 128, 256):
        for ks in (-1, 2, 3, 4, 6, 8):
            if ks > 0 and (K // 64) // ks < 4: continue
            spec = lambda: ops.GemmSpec(A, B, M, N, K, a_is_km=tn, b_is_kn=tn, out_dtype=torch.float32 if tn else None, tile_n=bn, split_k=ks)
            try:
                t = timeit(lambda: ops.gemm_grouped([spec()]))
                def cold():
                    flush.zero_(); ops.gemm_grouped([spec()])
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ts = []
                for _ in range(5):
                    flush.zero_(); torch.cuda.synchronize(); a.record(); ops.gemm_grouped([spec()]); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
                print(f"{M}x{N}x{K} bn={bn} ks={ks}: warm {t:.1f} us, cold-L2 {statistics.median(ts):.1f} us", flush=True)
            except Exception as e:
                print(f"{M}x{N}x{K} bn={bn} ks={ks}: {e!r:.100}")
    t = timeit(lambda: ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, a_is_km=tn, b_is_kn=tn, out_dtype=torch.float32 if tn else None)]))
    print(f"{M}x{N}x{K} planner: warm {t:.1f} us")
    if not tn:
        Af, Bf = A, B
        t = timeit(lambda: torch.matmul(Af, Bf.t()))
        print(f"{M}x{N}x{K} cuBLAS: warm {t:.1f} us")
    else:
        t = timeit(lambda: torch.matmul(A.t(), B))
        print(f"{M}x{N}x{K} cuBLAS (bf16 out): warm {t:.1f} us")
