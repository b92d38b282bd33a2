"""Where the c5 forward+backward chain spends its time: every piece timed alone (CUDA events, L2 flushed)."""
import os
import statistics
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for n, d in ((65536, 128), (65536, 256)):
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
    fwd = lambda: L.check(L.lib().msf_nce_grouped_fwd(pair, 1, L.MSF_BF16, 0.07, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, st), "fwd")
    bwd = lambda: L.check(L.lib().msf_nce_grouped_bwd(pair, 1, L.MSF_BF16, 0.07, 1e-8, gout.data_ptr(), ws.data_ptr(), wsb, st), "bwd")
    rn = lambda: ops.rownorm(q, torch.bfloat16)

    def both():
        fwd()
        bwd()
    print(f"D={d} (reverse order): fwd+bwd {timeit(both):.3f} ms, fwd {timeit(fwd):.3f} ms")
    print(f"D={d}: workspace {wsb / 1e6:.0f} MB; fwd {timeit(fwd):.3f} ms, bwd {timeit(bwd):.3f} ms, fwd+bwd {timeit(both):.3f} ms, rownorm {timeit(rn):.3f} ms")
print("ok")
