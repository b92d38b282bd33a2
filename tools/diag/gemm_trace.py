"""Per-CTA timeline of the grouped GEMM (needs the temporary trace build of gemm_grouped.cu)."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
dev = "cuda:0"
lib = L.lib()
fn = lib.msf_debug_gemm_trace
fn.argtypes = [C.c_void_p]; fn.restype = C.c_int
buf = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
for (M, N, K) in ((16384, 512, 512), (16384, 4096, 512), (256, 4608, 4608)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    spec = ops.GemmSpec(A, B, M, N, K, C=torch.empty((M, N), dtype=torch.bfloat16, device=dev))
    for _ in range(3):
        ops.gemm_grouped([spec])
    torch.cuda.synchronize()
    assert fn(buf.data_ptr()) == 0
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_grouped([spec]); e1.record()
    torch.cuda.synchronize()
    assert fn(0) == 0
    t = buf.view(148, 32).cpu()
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    print(f"--- {M}x{N}x{K}: event time {e0.elapsed_time(e1) * 1000:.1f} us; CTA start spread {(int(t[:, 0].max()) - t0) / 1000:.1f} us; last epilogue end {(int(t[:, 20].max()) - t0) / 1000:.1f} us")
    names = {0: "start"}
    for u in range(4):
        names.update({1 + 4 * u: f"u{u} first operands landed", 2 + 4 * u: f"u{u} MMAs issued", 3 + 4 * u: f"u{u} accumulator ready", 4 + 4 * u: f"u{u} epilogue done"})
    names[20] = "stores drained"
    for cta in (0, 73, 147):
        row = [(int(t[cta, i]) - t0) / 1000 for i in range(32)]
        print(f"CTA {cta}: " + "; ".join(f"{names[i]} {row[i]:.1f}" for i in sorted(names) if t[cta, i] > 0))
