"""ncu target: the gather-form crop backward at the two bench shapes (4x zoom, integer copy)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops
dev = "cuda:0"
for (B, C, H, W, oh, ow) in ((64, 128, 128, 128, 128, 128), (256, 128, 128, 128, 32, 32)):
    boxes = ops.footprint_boxes(B, 4, H, W, dev)
    go = torch.randn(B, 16, C, oh, ow, device=dev).to(torch.bfloat16)
    gf = torch.empty(B, C, H, W, device=dev)
    for _ in range(2):
        L.check(L.lib().msf_crop_resample_bwd(go.data_ptr(), B, C, H, W, boxes.data_ptr(), 16, oh, ow, L.MSF_BF16, gf.data_ptr(), L.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.check(L.lib().msf_crop_resample_bwd(go.data_ptr(), B, C, H, W, boxes.data_ptr(), 16, oh, ow, L.MSF_BF16, gf.data_ptr(), L.stream_ptr()), "bwd")
    e1.record()
    torch.cuda.synchronize()
    nb = go.numel() * 2 + gf.numel() * 4
    ms = e0.elapsed_time(e1) / 5
    print(f"oh={oh}: {ms:.3f} ms, {nb / ms / 1e6:.0f} GB/s")
    del go, gf
print("ok")
