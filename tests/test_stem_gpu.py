"""GPU parity of S1 (space-to-depth input layout for the stem convolution, src/models/resnet.py:155, 244) through the
C ABI: the layout itself is an exact copy / cast (bit-exact against torch's pad + pixel_unshuffle), and the 4x4/1
convolution over it reproduces Conv2d(3, 64, 7, stride 2, padding 3) forward and in the weight gradient."""
import pytest
import torch
import torch.nn.functional as F

from msfwsi_b200 import ops
from msfwsi_b200 import resnet as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_s2d(x, dtype):
    xs = F.pixel_unshuffle(F.pad(x, (3, 3, 3, 3)), 2)  # channel = c*4 + dy*2 + dx
    return F.pad(xs, (0, 0, 0, 0, 0, 16 - xs.shape[1])).to(dtype)


@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                                (torch.float16, torch.float16)])
@pytest.mark.parametrize("shape,cl", [((3, 3, 16, 20), True), ((2, 3, 224, 224), True), ((5, 3, 32, 18), False), ((2, 4, 8, 8), False), ((1, 1, 2, 2), True)])
def test_stem_s2d_layout_bit_exact(in_dtype, out_dtype, shape, cl):
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = torch.randn(shape, device=DEV, generator=g).to(in_dtype)
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    out = ops.stem_s2d(x, out_dtype)
    assert out.shape == (shape[0], 16, (shape[2] + 6) // 2, (shape[3] + 6) // 2) and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, _ref_s2d(x, out_dtype))


def test_stem_s2d_rejects_odd_sizes_and_wide_inputs():
    with pytest.raises(RuntimeError):
        ops.stem_s2d(torch.zeros(1, 3, 15, 16, device=DEV))
    with pytest.raises(RuntimeError):
        ops.stem_s2d(torch.zeros(1, 5, 16, 16, device=DEV))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_stem_convolution_in_space_to_depth_form_matches_conv2d(dtype, tol):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(6, 3, 64, 48, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w1 = (torch.randn(64, 3, 7, 7, device=DEV, generator=g) * 0.1).to(dtype).requires_grad_(True)
    w2 = w1.detach().clone().requires_grad_(True)
    y_ref = F.conv2d(x, w1, None, 2, 3)
    y = F.conv2d(ops.stem_s2d(x), ops.stem_s2d_weight(w2))
    assert y.shape == y_ref.shape
    gy = torch.randn(y.shape, device=DEV, generator=g).to(dtype)
    y_ref.backward(gy)
    y.backward(gy)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert rel(y, y_ref) <= tol and rel(w2.grad, w1.grad) <= tol
    assert w2.grad.shape == (64, 3, 7, 7)


def test_encoder_uses_the_fast_stem_and_keeps_the_parameter_shape():
    torch.backends.cudnn.allow_tf32 = False
    enc = R.resnet18(return_features=True).to(DEV).to(memory_format=torch.channels_last).train()
    x = torch.randn(4, 3, 64, 64, device=DEV).contiguous(memory_format=torch.channels_last)
    a, b = enc._stem_conv(x), enc.conv1(x)
    assert float((a - b).norm() / b.norm()) <= 1e-5
    assert enc.state_dict()["conv1.weight"].shape == (64, 3, 7, 7)
    xo = torch.randn(2, 3, 63, 64, device=DEV)  # odd height: plain conv path
    assert torch.equal(enc._stem_conv(xo), enc.conv1(xo))
