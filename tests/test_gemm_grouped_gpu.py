"""G1 grouped tcgen05 GEMM (msf_gemm_grouped) against fp64 torch on the same 16-bit operands: plain products in all four
operand layouts (y = x W^T, dX = dY W, dW = dY^T X), ragged edges, per-problem tiles, deterministic split-K, the batch-norm
statistics / row-norm epilogues, the batch-norm-apply + ReLU A prologue, fp16 operands, the TMA-store and the direct
store path, and a 24-problem head-stage launch.  Tolerance (stated): 16-bit outputs within one output ulp of the fp64
product (relative Frobenius <= 3e-3), fp32 outputs <= 2e-5 relative (fp32 accumulation in a different order);
statistics are compared with fp64 sums over the kernel's OWN rounded outputs (<= 2e-6 relative to the column scale)."""
import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def _mk(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(dtype)


def _ref(A, B, a_is_km, b_is_kn):
    a = A.double().t() if a_is_km else A.double()
    b = B.double() if b_is_kn else B.double().t()
    return a @ b


@pytest.mark.parametrize("M,N,K", [(256, 64, 64), (4096, 512, 512), (256, 4608, 4608), (304, 200, 136), (8, 16, 16), (1000, 576, 144), (128, 256, 4096)])
@pytest.mark.parametrize("layout", ["nt", "nn", "tn"])
@pytest.mark.parametrize("out", ["op", "f32"])
def test_single_problem_matches_fp64(M, N, K, layout, out):
    if M * N * K > 3e9 and layout != "nt":
        pytest.skip("large shape checked in the nt layout only")
    a_is_km, b_is_kn = layout == "tn", layout in ("nn", "tn")
    A = _mk((K, M) if a_is_km else (M, K), torch.bfloat16, 1)
    B = _mk((K, N) if b_is_kn else (N, K), torch.bfloat16, 2)
    odt = torch.float32 if out == "f32" else torch.bfloat16
    (C,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, a_is_km=a_is_km, b_is_kn=b_is_kn, out_dtype=odt)])
    ref = _ref(A, B, a_is_km, b_is_kn)
    assert C.shape == (M, N) and C.dtype == odt
    assert _rel(C, ref) <= (2e-5 if out == "f32" else 3e-3)


@pytest.mark.parametrize("no_tma_store", [False, True])
@pytest.mark.parametrize("tile_n", [0, 64, 128, 256])
def test_tiles_and_store_paths_agree_bit_for_bit(no_tma_store, tile_n):
    M, N, K = 520, 328, 264  # ragged in every dimension
    A, B = _mk((M, K), torch.bfloat16, 3), _mk((N, K), torch.bfloat16, 4)
    bias = _mk((N,), torch.float32, 5)
    (C0,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, bias=bias, tile_n=64, no_tma_store=True)])
    (C1,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, bias=bias, tile_n=tile_n, no_tma_store=no_tma_store)])
    ref = A.double() @ B.double().t() + bias.double()
    assert _rel(C0, ref) <= 3e-3
    assert torch.equal(C0, C1)  # same fp32 accumulation order per output element whatever the N tile / store path


@pytest.mark.parametrize("split_k", [2, 5, 32])
def test_split_k_is_deterministic_and_close(split_k):
    M, N, K = 512, 192, 8192  # the target heads' dW shape: K = rows of both views
    A, B = _mk((K, M), torch.bfloat16, 6), _mk((K, N), torch.bfloat16, 7)
    spec = lambda sk: ops.GemmSpec(A, B, M, N, K, a_is_km=True, b_is_kn=True, out_dtype=torch.float32, split_k=sk)
    (C,) = ops.gemm_grouped([spec(split_k)])
    (C2,) = ops.gemm_grouped([spec(split_k)])
    (C1,) = ops.gemm_grouped([spec(-1)])
    ref = A.double().t() @ B.double()
    assert torch.equal(C, C2), "split-K must be bit-reproducible (fixed-order reduction, no float atomics)"
    assert _rel(C, ref) <= 2e-5 and _rel(C1, ref) <= 2e-5
    (Cauto,) = ops.gemm_grouped([spec(0)])
    assert _rel(Cauto, ref) <= 2e-5


def test_split_k_counters_are_left_zero():
    M, N, K = 64, 64, 4096
    A, B = _mk((K, M), torch.bfloat16, 8), _mk((K, N), torch.bfloat16, 9)
    for _ in range(3):
        ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, a_is_km=True, b_is_kn=True, out_dtype=torch.float32, split_k=8)])
    torch.cuda.synchronize()
    assert int(ops._counters_for(torch.device(DEV)).abs().sum()) == 0


@pytest.mark.parametrize("M,N,K,tile_n", [(256, 64, 64, 0), (4096, 512, 128, 0), (96, 1152, 576, 0), (45, 200, 72, 0),
                                             # every tile width x split the launch planner may choose for the head-stage shapes
                                             (32, 1152, 4608, 64), (32, 1152, 4608, 128), (32, 1152, 4608, 256), (256, 4608, 4608, 256),
                                             (512, 512, 512, 128), (300, 576, 1152, 256)])
@pytest.mark.parametrize("split_k", [-1, 3, 0, 9])
def test_statistics_epilogues(M, N, K, tile_n, split_k):
    if split_k > 0 and K < 64 * 4 * split_k:
        pytest.skip("needs >= 4 k-blocks per split")
    A, B = _mk((M, K), torch.bfloat16, 10), _mk((N, K), torch.bfloat16, 11, 0.3)
    bias = _mk((N,), torch.float32, 12)
    g = ops.GemmSpec(A, B, M, N, K, bias=bias, split_k=split_k, tile_n=tile_n)
    (C,) = ops.gemm_grouped([g], want_col_stats=True, want_row_sumsq=True)
    y = C.double()
    assert _rel(C, A.double() @ B.double().t() + bias.double()) <= 3e-3
    groups = (M + 127) // 128  # one entry per 128-row M tile
    assert g.col_stats.shape == (groups, 2, N) and g.row_sumsq.shape == ((N + 63) // 64, M)
    pad = torch.zeros(groups * 128, N, dtype=torch.float64, device=DEV)
    pad[:M] = y
    pad = pad.view(groups, 128, N)
    s1, s2 = pad.sum(1), (pad * pad).sum(1)
    scale = s2.max().item() + 1e-30
    assert float((g.col_stats[:, 0].double() - s1).abs().max()) <= 2e-6 * max(1.0, s1.abs().max().item())
    assert float((g.col_stats[:, 1].double() - s2).abs().max()) <= 2e-6 * scale
    blocks = (N + 63) // 64
    padc = torch.zeros(M, blocks * 64, dtype=torch.float64, device=DEV)
    padc[:, :N] = y
    rs = (padc * padc).view(M, blocks, 64).sum(2).t()
    assert float((g.row_sumsq.double() - rs).abs().max()) <= 2e-6 * (rs.max().item() + 1e-30)


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (4096, 64, 512), (100, 1152, 288), (512, 16, 16), (64, 576, 144)])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_apply_relu_prologue(M, N, K, relu):
    """A' = relu?(bf16(A * scale + shift)) on the fly == the GEMM of the explicitly normalised activation."""
    A, B = _mk((M, K), torch.bfloat16, 13), _mk((N, K), torch.bfloat16, 14)
    scale, shift = _mk((K,), torch.float32, 15).abs() + 0.5, _mk((K,), torch.float32, 16)
    (C,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, a_scale=scale, a_shift=shift, a_relu=relu)])
    An = (A.double() * scale.double() + shift.double()).float().to(torch.bfloat16)  # = fp32 fma, then the 16-bit rounding of a BN output
    An = torch.relu(An) if relu else An
    (Cx,) = ops.gemm_grouped([ops.GemmSpec(An, B, M, N, K)])
    ref = An.double() @ B.double().t()
    assert _rel(C, ref) <= 3e-3
    assert torch.equal(C, Cx), "prologue must reproduce the explicit normalise -> GEMM sequence exactly"


def test_fp16_operands():
    M, N, K = 300, 136, 200
    A, B = _mk((M, K), torch.float16, 17), _mk((N, K), torch.float16, 18)
    scale, shift = _mk((K,), torch.float32, 19).abs() + 0.5, _mk((K,), torch.float32, 20)
    g = ops.GemmSpec(A, B, M, N, K, a_scale=scale, a_shift=shift, a_relu=True)
    (C,) = ops.gemm_grouped([g], want_col_stats=True)
    An = torch.relu((A.double() * scale.double() + shift.double()).float().to(torch.float16))
    assert C.dtype == torch.float16
    assert _rel(C, An.double() @ B.double().t()) <= 1e-3
    (Cw,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, out_dtype=torch.float32)])
    assert _rel(Cw, A.double() @ B.double().t()) <= 2e-5


def test_head_stage_launch_24_problems():
    """One launch = the first Linear of all 12 heads x 2 views at B = 64 (context rows 64, target rows 1024, fuser rows 64)."""
    B_, K16 = 64, 16
    specs, refs = [], []
    seed = 100
    for branch, rows, mult in (("ctx", B_, 1), ("tgt", B_ * K16, 1), ("inter", B_, 9)):
        for d in (64, 128, 256, 512):
            dim = d * mult
            Wt = _mk((dim, dim), torch.bfloat16, seed, 1.0 / dim ** 0.5)
            seed += 1
            for v in range(2):
                X = _mk((rows, dim), torch.bfloat16, seed).abs()
                seed += 1
                specs.append(ops.GemmSpec(X, Wt, rows, dim, dim))
                refs.append(X.double() @ Wt.double().t())
    before = L.launch_count
    outs = ops.gemm_grouped(specs, want_col_stats=True)
    assert L.launch_count - before == 1 and len(outs) == 24
    for g, C, ref in zip(specs, outs, refs):
        assert _rel(C, ref) <= 3e-3, (g.M, g.N)
        s1 = C.double().sum(0)
        got = g.col_stats[:, 0].double().sum(0)
        assert float((got - s1).abs().max()) <= 1e-5 * max(1.0, s1.abs().max().item())


def test_mixed_backward_launch_dx_and_dw():
    """dX = dY W and dW = dY^T X of several heads in one launch (mixed operand layouts, mixed output dtypes)."""
    specs, refs = [], []
    for i, (rows, din, dout) in enumerate(((256, 64, 64), (2048, 512, 128), (128, 1152, 288), (40, 576, 576))):
        X, Wt, dY = _mk((rows, din), torch.bfloat16, 200 + i), _mk((dout, din), torch.bfloat16, 210 + i), _mk((rows, dout), torch.bfloat16, 220 + i)
        specs.append(ops.GemmSpec(dY, Wt, rows, din, dout, b_is_kn=True))                                             # dX
        refs.append(dY.double() @ Wt.double())
        specs.append(ops.GemmSpec(dY, X, dout, din, rows, a_is_km=True, b_is_kn=True, out_dtype=torch.float32))        # dW
        refs.append(dY.double().t() @ X.double())
    outs = ops.gemm_grouped(specs)
    for g, C, ref in zip(specs, outs, refs):
        assert _rel(C, ref) <= (2e-5 if C.dtype == torch.float32 else 3e-3), (g.M, g.N, g.K)


def test_linear_bnstat_single_problem_abi():
    rows, fin, fout = 200, 128, 72
    x, w = _mk((rows, fin), torch.bfloat16, 300), _mk((fout, fin), torch.bfloat16, 301)
    y = torch.empty((rows, fout), dtype=torch.bfloat16, device=DEV)
    st = torch.empty(((rows + 127) // 128, 2, fout), dtype=torch.float32, device=DEV)
    L.check(L.lib().msf_linear_bnstat(x.data_ptr(), w.data_ptr(), y.data_ptr(), rows, fin, fout, L.MSF_BF16, st.data_ptr(), 0, 0, 0, L.stream_ptr()),
            "msf_linear_bnstat")
    assert _rel(y, x.double() @ w.double().t()) <= 3e-3
    assert float((st[:, 0].double().sum(0) - y.double().sum(0)).abs().max()) <= 1e-4


def test_argument_validation():
    A, B = _mk((64, 64), torch.bfloat16, 1), _mk((64, 64), torch.bfloat16, 2)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        sc = torch.ones(64, device=DEV)
        ops.gemm_grouped([ops.GemmSpec(A[:, :60], B[:, :60], 64, 64, 60, a_scale=sc, a_shift=sc)])
    (C60,) = ops.gemm_grouped([ops.GemmSpec(A[:, :60], B[:, :60], 64, 64, 60)])  # K only needs the 16-byte row pitch, not a multiple of 8
    assert _rel(C60, A[:, :60].double() @ B[:, :60].double().t()) <= 3e-3
    with pytest.raises(TypeError):
        ops.gemm_grouped([ops.GemmSpec(A.double(), B.double(), 64, 64, 64)])
    with pytest.raises(RuntimeError, match="prologue"):
        s = torch.ones(64, device=DEV)
        ops.gemm_grouped([ops.GemmSpec(A, B, 64, 64, 64, a_is_km=True, a_scale=s, a_shift=s)])


@pytest.mark.parametrize("M,N,K", [(256, 64, 64), (300, 200, 136), (7, 9, 5), (512, 576, 144)])
@pytest.mark.parametrize("layout", ["nt", "nn", "tn"])
def test_fp32_simt_path_is_exact_fp32(M, N, K, layout):
    """fp32 operands take the plain-FMA kernel: <= 1e-6 relative to fp64 (no tf32 rounding), bias and alpha honoured."""
    a_is_km, b_is_kn = layout == "tn", layout in ("nn", "tn")
    A = _mk((K, M) if a_is_km else (M, K), torch.float32, 31)
    B = _mk((K, N) if b_is_kn else (N, K), torch.float32, 32)
    bias = _mk((N,), torch.float32, 33)
    (C,) = ops.gemm_grouped([ops.GemmSpec(A, B, M, N, K, a_is_km=a_is_km, b_is_kn=b_is_kn, bias=bias, alpha=0.5)])
    ref = 0.5 * _ref(A, B, a_is_km, b_is_kn) + bias.double()
    assert C.dtype == torch.float32 and _rel(C, ref) <= 1e-6
