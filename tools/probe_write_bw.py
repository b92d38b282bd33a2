import torch, statistics
dev="cuda:0"
def t(fn, reps=5, iters=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(iters):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)/reps)
    return statistics.median(ts)
x=torch.empty(2<<30, dtype=torch.bfloat16, device=dev)  # 4 GiB
y=torch.empty_like(x)
ms=t(lambda: x.fill_(1.0)); print("fill 4GiB bf16: %.3f ms  %.0f GB/s"%(ms, x.numel()*2/ms/1e6))
ms=t(lambda: x.zero_()); print("zero 4GiB: %.3f ms  %.0f GB/s"%(ms, x.numel()*2/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy 4GiB->4GiB: %.3f ms  %.0f GB/s (r+w)"%(ms, 2*x.numel()*2/ms/1e6))
ms=t(lambda: x.sum()); print("read-only sum 4GiB: %.3f ms  %.0f GB/s"%(ms, x.numel()*2/ms/1e6))
s=torch.empty(64,128,128,128,dtype=torch.bfloat16,device=dev).normal_()
ms=t(lambda: torch.nn.functional.interpolate(s[:, :, :32, :32], scale_factor=4, mode="bilinear", align_corners=False))
print("torch interpolate (64,128,32,32)->128x128: %.3f ms, out %.0f MB -> x16 tiles would be %.2f ms"%(ms, 64*128*128*128*2/1e6, ms*16))
