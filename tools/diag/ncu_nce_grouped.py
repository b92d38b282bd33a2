"""ncu target: the grouped flash InfoNCE forward at c5 (N = Nq = 65536, D = 128 and 256), one launch each after warm-up."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
dev = "cuda:0"
n = 65536
for d in (128, 256):
    g = torch.Generator(device=dev).manual_seed(3407)
    k = torch.randn(n, d, device=dev, generator=g)
    q = (0.3 * k + torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
    kh, _ = ops.rownorm(k.to(torch.bfloat16), torch.bfloat16)
    qh, _ = ops.rownorm(q, torch.bfloat16)
    gq = torch.empty_like(qh)
    pr = (L.NcePair * 1)(L.NcePair(qh.data_ptr(), 0, kh.data_ptr(), gq.data_ptr(), n * d, n, n, 1, d, 0, 1.0))
    wsb = L.lib().msf_nce_grouped_workspace_bytes(pr, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    loss = torch.empty((), device=dev)
    gout = torch.ones((), device=dev)
    for _ in range(2):
        L.check(L.lib().msf_nce_grouped_fwd(pr, 1, L.MSF_BF16, 0.07, 1e-8, loss.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr()), "fwd")
        L.check(L.lib().msf_nce_grouped_bwd(pr, 1, L.MSF_BF16, 0.07, 1e-8, gout.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    print(d, float(loss))
print("ok")
