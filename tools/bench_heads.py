"""Head path (gather/concat + projectors/predictors + fused loss, forward + backward, no encoders) at the
config-2 sizes (B=256: context 256 rows, target 4096 rows, inter 256 rows), bf16 autocast, with the head Linears on
this repo's tcgen05 GEMM vs on cuBLAS, and the fused loss vs the reference's 24 CosineSimilarity calls."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import msfwsi_b200 as M  # noqa: E402

dev = "cuda:0"


class _Null(torch.nn.Module):
    def __init__(self, **_):
        super().__init__()
        self.fc = torch.nn.Identity()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    B, K = args.batch, 16
    model = M.MSFWSI(lambda **kw: _Null(**kw), 4).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda r, d: torch.randn(r, d, device=dev, generator=g).abs().to(torch.bfloat16).requires_grad_(True)
    cf = [tuple(mk(B, d) for d in (64, 128, 256, 512)) for _ in range(2)]
    tf = [tuple(mk(B * K, d) for d in (64, 128, 256, 512)) for _ in range(2)]
    rev = [torch.stack([torch.randperm(K) for _ in range(B)]).to(dev) for _ in range(2)]
    cos = torch.nn.CosineSimilarity(dim=1)
    W = M.DEFAULT_FUSER_WEIGHTS

    def step(mode):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model.heads(cf[0], cf[1], tf[0], tf[1], rev)
            if mode == "torch_cosine":
                loss = 0
                for br in out:
                    for i, (p1, p2, z1, z2) in enumerate(zip(*br)):
                        loss = loss + (-(cos(p1, z2).mean() + cos(p2, z1).mean()) * 0.5) * W[i]
            else:
                loss = M.ssl_loss(out, W, mode=mode)
        loss.backward()

    rows = []
    for use_tc in (True,):
        for mode in ("cosine", "infonce", "torch_cosine"):
            for _ in range(3):
                step(mode)
            ts = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                step(mode)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            r = {"linears": "tcgen05 (this repo)" if use_tc else "cuBLAS", "loss": mode, "ms_fwd_bwd": statistics.median(ts), "batch": B}
            rows.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
