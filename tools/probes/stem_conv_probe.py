"""Development probe: does cuDNN run the 7x7/2 stem convolution faster when the 3 input channels are pre-padded to 4 / 8?"""
import torch, torch.nn.functional as F, statistics
dev = "cuda:0"
torch.backends.cudnn.benchmark = True
N = 4096
def t(fn, it=5):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
x3 = torch.randn(N, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
w3 = torch.randn(64, 3, 7, 7, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
gy = torch.randn(N, 64, 112, 112, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for C in (3, 4, 8):
    x = x3 if C == 3 else F.pad(x3, (0, 0, 0, 0, 0, C - 3)).contiguous(memory_format=torch.channels_last)
    w = w3 if C == 3 else F.pad(w3.detach(), (0, 0, 0, 0, 0, C - 3)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    fwd = lambda: F.conv2d(x, w, None, 2, 3)
    y = fwd()
    bwd = lambda: torch.autograd.grad(y, w, gy, retain_graph=True)
    print(f"C={C}: fwd {t(fwd):.2f} ms, wgrad {t(bwd):.2f} ms", flush=True)
    if C != 3:
        print(f"      pad input 3->{C}: {t(lambda: F.pad(x3, (0, 0, 0, 0, 0, C - 3))):.2f} ms")
    del y
# space-to-depth alternative: 7x7/2 on 3 channels == 4x4/1 on 12 channels of the 2x2 pixel-unshuffled input (zero-padded taps)
x12 = F.pixel_unshuffle(F.pad(x3, (3, 5, 3, 5)), 2).contiguous(memory_format=torch.channels_last)  # (N,12,116,116)
w12 = torch.randn(64, 12, 4, 4, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
f12 = lambda: F.conv2d(x12, w12, None, 1, 0)
y12 = f12(); print("space-to-depth out", tuple(y12.shape))
gy12 = torch.randn_like(y12)
print(f"s2d C=12 4x4/1: fwd {t(f12):.2f} ms, wgrad {t(lambda: torch.autograd.grad(y12, w12, gy12, retain_graph=True)):.2f} ms")
x16 = F.pad(x12, (0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last)
w16 = torch.randn(64, 16, 4, 4, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
f16 = lambda: F.conv2d(x16, w16, None, 1, 0)
y16 = f16()
print(f"s2d C=16 4x4/1: fwd {t(f16):.2f} ms, wgrad {t(lambda: torch.autograd.grad(y16, w16, gy12, retain_graph=True)):.2f} ms")
