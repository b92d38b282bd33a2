"""ctypes binding of libmsfwsi_b200.so (the C ABI in include/msfwsi_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails this module
raises, loudly.  Build it with ``python -m msfwsi_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmsfwsi_b200.so")

MSF_F32, MSF_BF16, MSF_F16 = 0, 1, 2
MSF_GATHER_MAX_ITEMS = 16
MSF_COS_MAX_PAIRS = 32
MSF_EMA_CHUNK = 8192

_DTYPE = {torch.float32: MSF_F32, torch.bfloat16: MSF_BF16, torch.float16: MSF_F16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE[dt]
    except KeyError:
        raise TypeError(f"msfwsi_b200 supports float32/bfloat16/float16 tensors, got {dt}") from None


class GatherItem(C.Structure):
    _fields_ = [("tgt_f", C.c_void_p), ("ctx_f", C.c_void_p), ("rev", C.c_void_p), ("tgt_sorted", C.c_void_p),
                ("ms_f", C.c_void_p), ("d", C.c_int32), ("reserved", C.c_int32)]


class GatherGradItem(C.Structure):
    _fields_ = [("g_sorted", C.c_void_p), ("g_ms", C.c_void_p), ("rev", C.c_void_p), ("g_tgt_f", C.c_void_p),
                ("g_ctx_f", C.c_void_p), ("d", C.c_int32), ("reserved", C.c_int32)]


class CosPair(C.Structure):
    _fields_ = [("p", C.c_void_p), ("z", C.c_void_p), ("row_stats", C.c_void_p), ("grad_p", C.c_void_p),
                ("rows", C.c_int64), ("dim", C.c_int32), ("coef", C.c_float)]


class ProfRecord(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("launches", C.c_int32), ("work", C.c_double), ("ms", C.c_double)]


MSF_K_COUNT = 30
MSF_NCE_MAX_PAIRS = 32


class NcePair(C.Structure):
    _fields_ = [("q", C.c_void_p), ("q_rowsq", C.c_void_p), ("keys", C.c_void_p), ("grad_q", C.c_void_p), ("rank_stride", C.c_int64),
                ("nq", C.c_int32), ("rows_per_rank", C.c_int32), ("world", C.c_int32), ("D", C.c_int32), ("pos_rank", C.c_int32), ("coef", C.c_float)]

MSF_HEAD_MAX_ITEMS = 48
MSF_HEAD_MAX_MATS = 96
MSF_HEAD_SYNC_MAX_CTAS = 256


class HeadBnItem(C.Structure):
    _fields_ = [("col_stats", C.c_void_p * 2), ("scale", C.c_void_p * 2), ("shift", C.c_void_p * 2), ("mean", C.c_void_p * 2),
                ("invstd", C.c_void_p * 2), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("rows", C.c_int32), ("C", C.c_int32), ("n_views", C.c_int32), ("centered", C.c_int32), ("group_rows", C.c_int32)]


class HeadMat(C.Structure):
    _fields_ = [("x", C.c_void_p), ("col_stats", C.c_void_p), ("rows", C.c_int32), ("C", C.c_int32)]


class HeadApplyItem(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("y_hat", C.c_void_p), ("inv_norm", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("rows", C.c_int32), ("C", C.c_int32), ("relu", C.c_int32), ("reserved", C.c_int32)]


class HeadBwdItem(C.Structure):
    _fields_ = [("g", C.c_void_p), ("y", C.c_void_p), ("dy", C.c_void_p), ("partial", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p), ("c1", C.c_void_p), ("c2", C.c_void_p),
                ("rows", C.c_int32), ("C", C.c_int32), ("relu", C.c_int32), ("centered", C.c_int32)]


class HeadBwdFinItem(C.Structure):
    _fields_ = [("partial", C.c_void_p * 2), ("c1", C.c_void_p * 2), ("c2", C.c_void_p * 2), ("d_gamma", C.c_void_p), ("d_beta", C.c_void_p),
                ("rows", C.c_int32), ("C", C.c_int32), ("n_views", C.c_int32), ("plain", C.c_int32)]

MSF_GEMM_MAX_PROBLEMS = 48
MSF_GEMM_MAX_COUNTERS = 8192


class GemmProblem(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64), ("C", C.c_void_p), ("ldc", C.c_int64),
                ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("a_is_km", C.c_int32), ("b_is_kn", C.c_int32), ("out_dtype", C.c_int32),
                ("alpha", C.c_float), ("bias", C.c_void_p), ("col_stats", C.c_void_p), ("row_sumsq", C.c_void_p), ("a_scale", C.c_void_p),
                ("a_shift", C.c_void_p), ("a_relu", C.c_int32), ("tile_n", C.c_int32), ("split_k", C.c_int32), ("no_tma_store", C.c_int32),
                ("exp_a", C.c_float), ("row_sum_ld", C.c_int32), ("row_scale", C.c_void_p)]

MSF_ADAM_CHUNK = 4096
MSF_PEER_MAX_WORLD = 32


class AdamEntry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("ema", C.c_void_p),
                ("numel", C.c_int64), ("group", C.c_int32), ("shadow_dtype", C.c_int32), ("shadow", C.c_void_p)]


class EmaEntry(C.Structure):
    _fields_ = [("teacher", C.c_void_p), ("student", C.c_void_p), ("numel", C.c_int64)]


_SIGS = {
    "msf_abi_version": (C.c_int, []),
    "msf_last_error": (C.c_char_p, []),
    "msf_device_check": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "msf_gather_concat_fwd": (C.c_int, [C.POINTER(GatherItem), C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_gather_concat_bwd": (C.c_int, [C.POINTER(GatherGradItem), C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "msf_cosine_loss_workspace_bytes": (C.c_size_t, [C.POINTER(CosPair), C.c_int]),
    "msf_cosine_loss_fwd": (C.c_int, [C.POINTER(CosPair), C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_cosine_loss_bwd": (C.c_int, [C.POINTER(CosPair), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_rownorm": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_infonce_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "msf_infonce_plan_info": (C.c_int, [C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "msf_infonce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_infonce_fwd_timed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msf_infonce_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float,
                                  C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "msf_infonce_dk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "msf_infonce_dk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int, C.c_void_p, C.c_float,
                                 C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_infonce_dk_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_int,
                                        C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "msf_gemm_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "msf_gemm_grouped_workspace_bytes": (C.c_size_t, [C.POINTER(GemmProblem), C.c_int]),
    "msf_gemm_grouped": (C.c_int, [C.POINTER(GemmProblem), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "msf_gemm_grouped_f32": (C.c_int, [C.POINTER(GemmProblem), C.c_int, C.c_void_p]),
    "msf_gemm_grouped_plan_info": (C.c_int, [C.POINTER(GemmProblem), C.POINTER(C.c_int32)]),
    "msf_linear_bnstat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p]),
    "msf_nce_grouped_workspace_bytes": (C.c_size_t, [C.POINTER(NcePair), C.c_int]),
    "msf_nce_grouped_fwd": (C.c_int, [C.POINTER(NcePair), C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_nce_grouped_bwd": (C.c_int, [C.POINTER(NcePair), C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_head_sync_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "msf_head_bn_finalize": (C.c_int, [C.POINTER(HeadBnItem), C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                       C.c_int64, C.c_int, C.c_void_p]),
    "msf_head_bn_stats": (C.c_int, [C.POINTER(HeadMat), C.c_int, C.c_int, C.c_void_p]),
    "msf_head_bn_apply": (C.c_int, [C.POINTER(HeadApplyItem), C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "msf_head_bn_bwd_reduce": (C.c_int, [C.POINTER(HeadBwdItem), C.c_int, C.c_int, C.c_void_p]),
    "msf_head_bn_bwd_finalize": (C.c_int, [C.POINTER(HeadBwdFinItem), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64,
                                           C.c_int, C.c_void_p]),
    "msf_head_bn_bwd_elemt": (C.c_int, [C.POINTER(HeadBwdItem), C.c_int, C.c_int, C.c_void_p]),
    "msf_crop_resample_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "msf_crop_resample_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "msf_ema_plan": (C.c_int, [C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int32)]),
    "msf_ema_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "msf_bn2d_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "msf_bn2d_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_bn2d_stats_finalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_bn2d_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msf_bn2d_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_void_p]),
    "msf_bn2d_bwd_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "msf_bn2d_bwd_elemt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msf_bn2d_apply_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msf_bn2d_pool_bwd_elemt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "msf_adam_plan": (C.c_int, [C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int32)]),
    "msf_grad_check_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_adam_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "msf_stem_s2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                               C.c_int, C.c_void_p]),
    "msf_peer_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "msf_peer_allreduce_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
    "msf_jigsaw_tiles": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_view_crops_s2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "msf_prof_begin": (C.c_int, [C.c_int]),
    "msf_prof_end": (C.c_int, [C.POINTER(ProfRecord), C.POINTER(C.c_int)]),
    "msf_prof_kernel_name": (C.c_char_p, [C.c_int]),
    "msf_prof_kernel_bound": (C.c_int, [C.c_int]),
}
EXPORTS = tuple(_SIGS)

_lib = None
launch_count = 0  # number of C-ABI compute calls made (bench.py reports kernel launches from it)

# Parameters are also written through raw pointers (msf_adam_multi, msf_ema_multi), which never moves a tensor's
# `_version`.  Every such writer bumps this epoch; anything derived from parameter values (the 16-bit operand copies of
# the head Linears) keys its cache on it -- unless the writer itself keeps the copy in sync (FusedAdam shadows).
param_epoch = 0


def bump_param_epoch() -> None:
    global param_epoch
    param_epoch += 1


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is not built (run `python -m msfwsi_b200.build`). "
                "msfwsi_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().msf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("msfwsi_b200 ops run on CUDA tensors only (there is no CPU fallback)")


def prof_begin(capacity: int = 1 << 16) -> None:
    """Switch the library's launch profiler on (CUDA event pairs around every main kernel; see msf_prof_begin)."""
    check(lib().msf_prof_begin(capacity), "msf_prof_begin")


def prof_end():
    """Switch it off; returns ({kernel name: {"launches", "work", "ms", "bound"}}, dropped)."""
    recs = (ProfRecord * MSF_K_COUNT)()
    dropped = C.c_int(0)
    check(lib().msf_prof_end(recs, C.byref(dropped)), "msf_prof_end")
    out = {}
    for r in recs:
        if r.launches:
            out[lib().msf_prof_kernel_name(r.kernel).decode()] = {"launches": r.launches, "work": r.work, "ms": r.ms,
                                                                  "bound": chr(lib().msf_prof_kernel_bound(r.kernel))}
    return out, dropped.value
