"""ncu target: the A2 forward strip kernel (4x zoom, bf16) and the InfoNCE key-gradient pass (transposed flash kernel, D = 128 and
256 at N = Nq = 65536), a couple of launches each after warm-up."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
dev = "cuda:0"
Bc, Cc, H, W, oh, ow = 64, 128, 128, 128, 128, 128
feat = torch.randn(Bc, Cc, H, W, device=dev).to(torch.bfloat16)
boxes = ops.footprint_boxes(Bc, 4, H, W, dev)
outp = torch.empty((Bc, 16, Cc, oh, ow), dtype=torch.bfloat16, device=dev)
for _ in range(3):
    L.check(L.lib().msf_crop_resample_fwd(feat.data_ptr(), Bc, Cc, H, W, boxes.data_ptr(), 16, oh, ow, L.MSF_BF16, outp.data_ptr(), L.stream_ptr()), "crop")
torch.cuda.synchronize()
del feat, outp
n = 65536
g1 = torch.ones((), device=dev)
for d in (128, 256):
    g = torch.Generator(device=dev).manual_seed(3407)
    k = torch.randn(n, d, device=dev, generator=g)
    q = (0.3 * k + torch.randn(n, d, device=dev, generator=g)).to(torch.bfloat16)
    kh, _ = ops.rownorm(k.to(torch.bfloat16), torch.bfloat16)
    qh, _ = ops.rownorm(q, torch.bfloat16)
    prec = L.MSF_BF16
    wsb = L.lib().msf_infonce_workspace_bytes(n, n, d, prec)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    dwb = L.lib().msf_infonce_dk_workspace_bytes(n, n, d, prec)
    dws = torch.empty(dwb, dtype=torch.uint8, device=dev)
    loss = torch.empty((), device=dev)
    dk = torch.empty((n, d), device=dev)
    L.check(L.lib().msf_infonce_fwd(qh.data_ptr(), kh.data_ptr(), n, n, d, 0, 0.07, prec, loss.data_ptr(), 0, ws.data_ptr(), wsb, L.stream_ptr()), "fwd")
    for _ in range(2):
        L.check(L.lib().msf_infonce_dk(qh.data_ptr(), kh.data_ptr(), n, n, d, 0, 0.07, prec, g1.data_ptr(), 1.0 / n, ws.data_ptr(), wsb, dk.data_ptr(),
                                       dws.data_ptr(), dwb, L.stream_ptr()), "dk")
    torch.cuda.synchronize()
    print(d, float(loss))
    del ws, dws, dk
print("ok")
