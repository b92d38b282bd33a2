"""D1b (msf_view_crops_s2d): all views of a step from the uint8 source tiles in one launch, in the stem convolution's
space-to-depth layout, against the oracle (crop -> F.interpolate bilinear -> flip -> normalise -> pad/pixel-unshuffle).
Tiling / permutation / crop coordinates are integer and exact (jigsaw_view_crops uses blockshaped's raster formula,
pinned by tests/golden/blockshaped.npz); the resampling is float (tolerance 2e-5 fp32, one bf16 ulp in bf16)."""
import pytest
import torch

from msfwsi_b200 import ops
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN, STD = (0.6998, 0.4785, 0.6609), (0.2203, 0.2407, 0.1983)  # scripts/bcss.sh:13-14


def _src(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)


def _random_crops(n, B, H, W, g, min_side=8):
    rows = []
    for _ in range(n):
        ch = int(torch.randint(min_side, H + 1, (1,), generator=g))
        cw = int(torch.randint(min_side, W + 1, (1,), generator=g))
        y0 = int(torch.randint(0, H - ch + 1, (1,), generator=g))
        x0 = int(torch.randint(0, W - cw + 1, (1,), generator=g))
        rows.append([int(torch.randint(0, B, (1,), generator=g)), y0, x0, y0 + ch, x0 + cw, int(torch.randint(0, 2, (1,), generator=g))])
    return torch.tensor(rows, dtype=torch.int32)


@pytest.mark.parametrize("H,W,oh,ow", [(64, 64, 16, 16), (96, 80, 32, 24), (256, 256, 224, 224)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_matches_oracle(H, W, oh, ow, dtype):
    g = torch.Generator().manual_seed(H + ow)
    B = 3
    src = _src(B, H, W, 1)
    crops = _random_crops(7, B, H, W, g)
    crops[0] = torch.tensor([1, 0, 0, H, W, 0], dtype=torch.int32)          # whole image (context view without a crop)
    crops[1] = torch.tensor([2, H - 8, W - 8, H, W, 1], dtype=torch.int32)  # bottom-right corner: the last bytes of the buffer
    out = ops.view_crops_s2d(src.to(DEV), crops.to(DEV), (oh, ow), MEAN, STD, dtype, validate=True)
    ref = O.view_crops_s2d(src, crops.tolist(), (oh, ow), MEAN, STD)
    assert out.shape == ref.shape == (7, 16, (oh + 6) // 2, (ow + 6) // 2) and out.is_contiguous(memory_format=torch.channels_last)
    err = (out.double().cpu() - ref).abs().max().item()
    assert err <= (3e-5 if dtype == torch.float32 else 4e-2), err  # values reach |(0-0.7*255)/(0.2*255)| ~ 3.5: one bf16 ulp = 1.6e-2
    assert torch.equal(out[:, 12:], torch.zeros_like(out[:, 12:])), "channels 12..15 are zero"
    assert torch.equal(out[:, :, 0, :].float(), torch.zeros_like(out[:, :, 0, :].float())), "3 rows of zero padding -> the first s2d row is zero"


def test_identity_scale_is_an_exact_copy_and_matches_stem_s2d():
    """Crop size == view size: bilinear weights vanish, so the result is exactly normalise(crop) in the stem layout -- the
    same tensor msf_stem_s2d produces from the normalised view (S1), which ties D1b to the pinned tiling of blockshaped."""
    H = W = 64
    src = _src(2, H, W, 5)
    g = torch.Generator().manual_seed(3)
    perm = torch.stack([torch.randperm(16, generator=g) for _ in range(2)])
    boxes = torch.tensor([0, 0, 16, 16]).view(1, 1, 4).expand(2, 16, 4)
    crops = ops.jigsaw_view_crops(perm.to(DEV), boxes.to(DEV), torch.zeros(2, 16, dtype=torch.int64, device=DEV), H, W, 4)
    out = ops.view_crops_s2d(src.to(DEV), crops, (16, 16), MEAN, STD, torch.float32)
    m = torch.tensor(MEAN).view(1, 3, 1, 1) * 255.0
    s = torch.tensor(STD).view(1, 3, 1, 1) * 255.0
    for b in range(2):
        tiles = O.blockshaped(src[b], 16, 16)[perm[b]]  # the reference's tiling + shuffle (bcss.py:171-177)
        view = ((tiles.permute(0, 3, 1, 2).float() - m) / s).to(DEV)
        want = ops.stem_s2d(view, torch.float32)
        got = out[b * 16:(b + 1) * 16]
        assert torch.allclose(got, want, rtol=0, atol=2e-6), float((got - want).abs().max())


def test_out_of_range_crop_is_flagged():
    src = _src(1, 32, 32, 2).to(DEV)
    bad = torch.tensor([[0, 0, 0, 40, 32, 0]], dtype=torch.int32, device=DEV)
    with pytest.raises(IndexError):
        ops.view_crops_s2d(src, bad, (16, 16), MEAN, STD, torch.float32, validate=True)
    with pytest.raises(ValueError):
        ops.view_crops_s2d(src, bad.to(torch.int64), (16, 16))


def test_encoder_accepts_the_s2d_views():
    import warnings
    import msfwsi_b200 as M
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc = M.resnet18(return_features=True, zero_init_residual=True).to(DEV).to(memory_format=torch.channels_last).train()
    enc.fc = torch.nn.Identity()
    src = _src(2, 128, 128, 9)
    crops = torch.tensor([[0, 0, 0, 128, 128, 0], [1, 10, 20, 100, 120, 1]], dtype=torch.int32)
    s2d = ops.view_crops_s2d(src.to(DEV), crops.to(DEV), (64, 64), MEAN, STD, torch.bfloat16)
    views = O.view_crops_s2d(src, crops.tolist(), (64, 64), MEAN, STD)  # only to rebuild the plain (n,3,64,64) views below
    plain = torch.nn.functional.pixel_shuffle(views[:, :12], 2)[:, :, 3:-3, 3:-3].float().to(DEV)
    enc2 = __import__("copy").deepcopy(enc)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        fa = enc(s2d)
        fb = enc2(plain)
    for a, b in zip(fa, fb):
        assert (a.float() - b.float()).norm() <= 3e-2 * b.float().norm()
