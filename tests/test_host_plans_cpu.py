"""Host-side planning logic of the C ABI, callable without a GPU: the grouped GEMM's launch planner
(msf_gemm_grouped_plan_info: tile width, deterministic split-K) and the workspace sizes of the InfoNCE entry points."""
import ctypes as C

import pytest

from msfwsi_b200 import _lib as L


def _plan(M, N, K, tile_n=0, split_k=0):
    arr = (L.GemmProblem * 1)()
    g = arr[0]
    g.M, g.N, g.K, g.out_dtype, g.tile_n, g.split_k = M, N, K, L.MSF_BF16, tile_n, split_k
    info = (C.c_int32 * 6)()
    L.check(L.lib().msf_gemm_grouped_plan_info(arr, info), "plan_info")
    return dict(zip(("bn", "tiles_m", "tiles_n", "ks", "kbps", "kb"), info))


@pytest.mark.parametrize("M,N,K", [(8192, 8192, 8192), (16384, 512, 512), (4096, 64, 64), (256, 4608, 4608), (1024, 4608, 4608),
                                   (512, 512, 8192), (8, 16, 16), (300, 200, 136), (4608, 4608, 512), (256, 576, 576)])
def test_plan_is_a_valid_cover(M, N, K):
    p = _plan(M, N, K)
    assert p["bn"] in (64, 128, 256)
    assert p["tiles_m"] == (M + 127) // 128 and p["tiles_n"] == (N + p["bn"] - 1) // p["bn"]
    assert p["kb"] == (K + 63) // 64
    assert 1 <= p["ks"] <= 32 and p["ks"] * p["kbps"] >= p["kb"] > (p["ks"] - 1) * p["kbps"]  # every split has work
    assert p["ks"] == 1 or p["kbps"] >= 4                                                      # at least 4 k-blocks per split
    assert p["bn"] == 64 or p["bn"] // 2 < N                                                   # no tile twice as wide as the problem


def test_plan_prefers_wide_tiles_for_weight_streaming_problems():
    """A 256-row problem under a 4608 x 4608 weight matrix: narrow tiles re-read the A panel per column tile, so the planner keeps
    bn >= 128 and fills the 148 SMs in one wave (with split-K if needed) instead of 288 units of bn = 64."""
    p = _plan(256, 4608, 4608)
    assert p["bn"] >= 128 and p["tiles_m"] * p["tiles_n"] * p["ks"] <= 148
    assert _plan(8192, 8192, 8192) == {"bn": 256, "tiles_m": 64, "tiles_n": 32, "ks": 1, "kbps": 128, "kb": 128}


def test_forced_tile_and_split_are_respected():
    p = _plan(512, 512, 8192, tile_n=128, split_k=8)
    assert p["bn"] == 128 and p["ks"] == 8 and p["kbps"] == 16
    assert _plan(512, 512, 8192, split_k=-1)["ks"] == 1


@pytest.mark.parametrize("nq,n_keys,dim,prec", [(4096, 4096, 128, L.MSF_BF16), (512, 4096, 256, L.MSF_BF16), (300, 300, 64, L.MSF_F32),
                                               (256, 2048, 576, L.MSF_BF16), (5, 5, 128, L.MSF_BF16)])
def test_infonce_workspaces_cover_their_buffers(nq, n_keys, dim, prec):
    lib = L.lib()
    fwd = lib.msf_infonce_workspace_bytes(nq, n_keys, dim, prec)
    dk = lib.msf_infonce_dk_workspace_bytes(nq, n_keys, dim, prec)
    info = (C.c_int64 * 8)()
    L.check(lib.msf_infonce_plan_info(nq, n_keys, dim, prec, info), "plan_info")
    splits, nq_pad = info[0], info[1]
    assert fwd >= splits * nq_pad * (dim + 1) * 4  # O partials + row sums
    # key gradient: per-column terms + the transposed pass's partials (rows = all keys), or Q_hat / sum + the GEMM output for wide rows
    assert dk >= ((nq + 127) // 128) * 128 * 4 + n_keys * dim * 4
    assert lib.msf_infonce_dk_workspace_bytes(0, n_keys, dim, prec) == 0
