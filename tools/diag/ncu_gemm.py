"""ncu target: a few launches of the grouped GEMM (single tensor-bound problem, short-K problem, 24-problem head stage)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from msfwsi_b200 import ops
dev = "cuda:0"
def mk(M, N, K):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    return ops.GemmSpec(A, B, M, N, K, C=torch.empty((M, N), dtype=torch.bfloat16, device=dev))
cases = [[mk(16384, 512, 512)], [mk(8192, 8192, 8192)], [mk(256, 4608, 4608)]]
specs = []
for rows_, mult in ((256, 1), (4096, 1), (256, 9)):
    for d in (64, 128, 256, 512):
        dim = d * mult
        W = (torch.randn(dim, dim, device=dev) / dim ** 0.5).to(torch.bfloat16)
        for v in range(2):
            X = torch.randn(rows_, dim, device=dev).abs().to(torch.bfloat16)
            specs.append(ops.GemmSpec(X, W, rows_, dim, dim, C=torch.empty((rows_, dim), dtype=torch.bfloat16, device=dev)))
cases.append(specs)
for _ in range(3):
    for c in cases:
        ops.gemm_grouped(c, want_col_stats=len(c) > 1)
torch.cuda.synchronize()
print("ok")
