"""GPU parity of D1 (on-device tiling + jigsaw shuffle + resize + normalise, src/utils/data/bcss.py:171-177, 203-216)
through the C ABI against the oracle: tile coordinates and the permutation are exact (oracle pinned to the reference's
blockshaped by the golden fixture), the float bilinear resampling within 1e-5 in fp32 / one ulp in bf16."""
import pytest
import torch

from msfwsi_b200 import ops
from oracle import msf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN, STD = (0.7, 0.55, 0.68), (0.17, 0.21, 0.15)  # BCSS-like statistics (scripts/bcss.sh:13-14 pass the real ones)


def _src(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2 ** -7)])
@pytest.mark.parametrize("B,H,W,grid,oh,ow", [(3, 64, 64, 4, 14, 14), (2, 256, 128, 4, 56, 24), (2, 96, 96, 2, 64, 64), (1, 1024, 1024, 4, 224, 224),
                                              (2, 48, 80, 1, 24, 40)])
def test_jigsaw_tiles_vs_oracle(dtype, tol, B, H, W, grid, oh, ow):
    K = grid * grid
    g = torch.Generator().manual_seed(B + H)
    src = _src(B, H, W, 7)
    perm = torch.stack([torch.randperm(K, generator=g) for _ in range(B)])
    out = ops.jigsaw_tiles(src.to(DEV), perm.to(DEV), grid, (oh, ow), MEAN, STD, dtype)
    assert out.shape == (B * K, 3, oh, ow) and out.is_contiguous(memory_format=torch.channels_last)
    for b in range(B):
        ref = O.jigsaw_tile_views(src[b], perm[b], grid, (oh, ow), MEAN, STD)
        got = out[b * K:(b + 1) * K].double().cpu()
        assert torch.allclose(got, ref, rtol=tol, atol=tol), float((got - ref).abs().max())


def test_tiles_are_exactly_blockshaped_and_shuffled_when_no_resampling():
    """out size == tile size: every output pixel is one source pixel; order = blockshaped(img)[jigsaw_idx] (bcss.py:175-177)."""
    B, H, W, grid = 2, 32, 48, 4
    src = _src(B, H, W, 3)
    g = torch.Generator().manual_seed(1)
    perm = torch.stack([O.jigsaw_indices(g, 16)[0] for _ in range(B)])
    out = ops.jigsaw_tiles(src.to(DEV), perm.to(DEV), grid, (H // grid, W // grid), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), torch.float32)
    for b in range(B):
        tiles = O.blockshaped(src[b], H // grid, W // grid)[perm[b]]            # (16, th, tw, 3) uint8
        want = tiles.permute(0, 3, 1, 2).float() / 255.0
        assert torch.allclose(out[b * 16:(b + 1) * 16].cpu(), want, rtol=0, atol=1e-6)
    ident = ops.jigsaw_tiles(src.to(DEV), None, grid, (H // grid, W // grid), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), torch.float32)
    assert torch.allclose(ident[:16].cpu(), O.blockshaped(src[0], H // grid, W // grid).permute(0, 3, 1, 2).float() / 255.0, atol=1e-6)


def test_jigsaw_tiles_errors_mirror_the_reference():
    src = _src(1, 30, 32, 1).to(DEV)
    with pytest.raises(AssertionError):  # bcss.py:212-213: not evenly divisible
        ops.jigsaw_tiles(src, None, 4)
    src = _src(1, 32, 32, 1).to(DEV)
    bad = torch.arange(16).view(1, 16).to(DEV)
    bad[0, 3] = 16
    with pytest.raises(IndexError):
        ops.jigsaw_tiles(src, bad, 4, (8, 8), validate=True)
    neg = torch.arange(16).view(1, 16).to(DEV)
    neg[0, 0] = -16  # Python-style negative index = tile 0
    a = ops.jigsaw_tiles(src, neg, 4, (8, 8), validate=True)
    b = ops.jigsaw_tiles(src, torch.arange(16).view(1, 16).to(DEV), 4, (8, 8))
    assert torch.equal(a, b)
    with pytest.raises(AssertionError):
        ops.jigsaw_tiles(src, torch.zeros(1, 4, dtype=torch.int64, device=DEV), 4)
