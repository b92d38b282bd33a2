"""CPU-only checks: the C-ABI library loads, exports every symbol include/*.h declares, and validates
arguments before touching a device.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest
import torch

from msfwsi_b200 import _lib as L
from msfwsi_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "msfwsi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msf_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(L.LIB_PATH), "run `python -m msfwsi_b200.build` (or __graft_entry__.build())"
    assert L.LIB_PATH.startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 16
    handle = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/msfwsi_b200.h but not exported"
    assert sorted(L.EXPORTS) == names, "ctypes binding and header disagree"
    assert L.lib().msf_abi_version() == 2


def test_struct_layouts_match_header():
    # sizes the C side assumes (LP64): see include/msfwsi_b200.h
    assert C.sizeof(L.GatherItem) == 48 and C.sizeof(L.GatherGradItem) == 48
    assert C.sizeof(L.CosPair) == 48 and C.sizeof(L.EmaEntry) == 24


def test_argument_validation_without_device():
    lib = L.lib()
    items = (L.GatherItem * 1)()
    assert lib.msf_gather_concat_fwd(items, 99, 1, 16, 8, L.MSF_F32, None, None) == -1
    assert b"n_items" in lib.msf_last_error()
    items[0] = L.GatherItem(0, 0, 0, 0, 0, 64, 0)
    assert lib.msf_gather_concat_fwd(items, 1, 4, 16, 8, L.MSF_F32, None, None) == -1  # NULL pointers
    assert lib.msf_gather_concat_fwd(items, 1, 4, 16, 99, L.MSF_F32, None, None) == -1  # n_keep > K
    items[0] = L.GatherItem(16, 16, 16, 16, 16, 6, 0)
    assert lib.msf_gather_concat_fwd(items, 1, 4, 16, 8, L.MSF_BF16, None, None) == -1  # d % 8
    # InfoNCE: tau below the supported bound, positives outside the key range, unsupported width
    assert lib.msf_infonce_fwd(16, 16, 8, 8, 64, 0, 0.001, L.MSF_F32, 16, None, 16, 1 << 20, None) == -2
    assert b"tau" in lib.msf_last_error()
    assert lib.msf_infonce_fwd(16, 16, 8, 8, 64, 4, 0.07, L.MSF_F32, 16, None, 16, 1 << 20, None) == -1
    assert lib.msf_infonce_fwd(16, 16, 8, 8, 72, 0, 0.07, L.MSF_BF16, 16, None, 16, 1 << 20, None) == -2
    assert lib.msf_infonce_fwd(16, 16, 8, 8, 64, 0, 0.07, L.MSF_F32, 16, None, 16, 8, None) == -3  # workspace too small
    assert lib.msf_rownorm(16, 4, 12, L.MSF_F32, 1e-8, 16, L.MSF_BF16, None, None) == -1


def test_plans_are_pure_host_functions():
    lib = L.lib()
    numels = (C.c_int64 * 4)(1, 8192, 8193, 0)
    prefix = (C.c_int32 * 5)()
    assert lib.msf_ema_plan(numels, 4, prefix) == 0
    assert list(prefix) == [0, 1, 2, 4, 4]
    small = lib.msf_infonce_workspace_bytes(256, 256, 128, L.MSF_BF16)
    big = lib.msf_infonce_workspace_bytes(65536, 65536, 128, L.MSF_BF16)
    assert 0 < small < big
    assert big >= 65536 * 128 * 4  # at least one fp32 O partial
    assert lib.msf_infonce_workspace_bytes(0, 16, 64, L.MSF_F32) == 0


def test_ops_refuse_cpu_tensors():
    p, z = torch.randn(4, 64), torch.randn(4, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cosine_loss([p], [z], [-0.5])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.infonce_loss(p, z)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gather_concat([torch.randn(2, 64)], [torch.randn(32, 64)], [torch.zeros(2, 16, dtype=torch.int64)])
    with pytest.raises(TypeError):
        L.dtype_code(torch.float64)
    with pytest.raises(ValueError):
        ops.cosine_loss([p] * 33, [z] * 33, [1.0] * 33)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libmsfwsi_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        L.lib()
