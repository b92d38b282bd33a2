"""ncu target: D1b (msf_view_crops_s2d) at the bench shape: 64 source tiles of 1024^2 -> 2 context + 32 target views each."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from msfwsi_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
Bs, S = 64, 1024
src = torch.randint(0, 256, (Bs, S, S, 3), dtype=torch.uint8, device=dev, generator=g)
cr = []
for v in range(2):
    side = torch.randint(724, 1025, (Bs,), device=dev, generator=g)
    y0 = (torch.rand(Bs, device=dev, generator=g) * (S - side + 1)).long()
    x0 = (torch.rand(Bs, device=dev, generator=g) * (S - side + 1)).long()
    cr.append(torch.stack((torch.arange(Bs, device=dev), y0, x0, y0 + side, x0 + side, torch.randint(0, 2, (Bs,), device=dev, generator=g)), 1))
    perm = torch.stack([torch.randperm(16, device=dev, generator=g) for _ in range(Bs)])
    tside = torch.randint(114, 257, (Bs, 16), device=dev, generator=g)
    ty = (torch.rand(Bs, 16, device=dev, generator=g) * (256 - tside + 1)).long()
    tx = (torch.rand(Bs, 16, device=dev, generator=g) * (256 - tside + 1)).long()
    cr.append(ops.jigsaw_view_crops(perm, torch.stack((ty, tx, ty + tside, tx + tside), 2), torch.randint(0, 2, (Bs, 16), device=dev, generator=g), S, S, 4))
crops = torch.cat(cr).to(torch.int32)
mean, std = (0.6998, 0.4785, 0.6609), (0.2203, 0.2407, 0.1983)
for _ in range(3):
    out = ops.view_crops_s2d(src, crops, (224, 224), mean, std, torch.bfloat16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = ops.view_crops_s2d(src, crops, (224, 224), mean, std, torch.bfloat16)
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 10:.3f} ms for {crops.shape[0]} views")
# target views only / context views only
for name, sel in (("context", crops[(crops[:, 3] - crops[:, 1]) > 256]), ("target", crops[(crops[:, 3] - crops[:, 1]) <= 256])):
    e0.record()
    for _ in range(10):
        ops.view_crops_s2d(src, sel.contiguous(), (224, 224), mean, std, torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms for {sel.shape[0]} views")
print("ok")
