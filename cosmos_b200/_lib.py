"""ctypes binding of libcosmos_b200.so (the C ABI in include/cosmos_b200.h).

There is no CPU fallback: if the library cannot be loaded, or a call returns a non-zero status,
a RuntimeError is raised.  The library is built in-tree by `cosmos_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COSMOS_B200_LIB") or os.path.join(HERE, "libcosmos_b200.so")     # (override: A/B of two builds)

DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
ABI_VERSION = 3
CLAMP_MAX = 8   # COSMOS_CLAMP_MAX

_DEBUG_SYNC = os.environ.get("COSMOS_B200_DEBUG_SYNC", "0") == "1"
_lock = threading.Lock()
_lib = None


class EmaChunk(C.Structure):
    _fields_ = [("teacher", C.c_uint64), ("student", C.c_uint64), ("count", C.c_uint32), ("aligned", C.c_uint32)]


def _declare(lib):
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    lib.cosmos_abi_version.restype = i32
    lib.cosmos_abi_version.argtypes = []
    lib.cosmos_status_string.restype = C.c_char_p
    lib.cosmos_status_string.argtypes = [i32]
    lib.cosmos_last_cuda_error.restype = i32
    lib.cosmos_last_cuda_error.argtypes = []
    lib.cosmos_cuda_error_string.restype = C.c_char_p
    lib.cosmos_cuda_error_string.argtypes = [i32]
    lib.cosmos_device_check.restype = i32
    lib.cosmos_device_check.argtypes = [i32]
    lib.cosmos_ema_table_entries.restype = i64
    lib.cosmos_ema_table_entries.argtypes = [i64, C.POINTER(i64)]
    lib.cosmos_ema_table_fill.restype = i32
    lib.cosmos_ema_table_fill.argtypes = [i64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(i64), i32, vp]
    lib.cosmos_ema_apply.restype = i32
    lib.cosmos_ema_apply.argtypes = [vp, i64, f64, i32, i32, vp]
    lib.cosmos_clamp_scalars.restype = i32
    lib.cosmos_clamp_scalars.argtypes = [C.POINTER(C.c_uint64), i32, f64, f64, i32, i32, vp]
    for name, args in _POOL_SIGS.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = args
    for name, args in _INFONCE_SIGS.items():
        fn = getattr(lib, name)
        fn.restype = i64 if name in ("cosmos_infonce_workspace_bytes", "cosmos_infonce_e_bytes",
                                     "cosmos_infonce_bwd_e_workspace_bytes") else i32
        fn.argtypes = args


vp_, i32_, i64_, f32_ = C.c_void_p, C.c_int, C.c_int64, C.c_float


class InfoNceProblem(C.Structure):
    """Mirror of cosmos_infonce_problem (include/cosmos_b200.h)."""
    _fields_ = [
        ("x", C.c_uint64), ("y", C.c_uint64),
        ("gx", C.c_int32), ("gy", C.c_int32),
        ("n_rows", C.c_int32), ("n_cols", C.c_int32),
        ("dim", C.c_int32), ("label_offset", C.c_int32),
        ("dtype", C.c_int32), ("reserved", C.c_int32),
        ("scale", C.c_uint64),
    ]


_INFONCE_SIGS = {
    "cosmos_infonce_workspace_bytes": [C.POINTER(InfoNceProblem)],
    "cosmos_infonce_fwd": [C.POINTER(InfoNceProblem), vp_, vp_, vp_, vp_, i64_, i32_, vp_],
    "cosmos_scale16": [vp_, vp_, vp_, f32_, i32_, i64_, i32_, vp_],
    "cosmos_lse2_merge": [vp_, vp_, i32_, i64_, i32_, vp_],
    "cosmos_infonce_loss_sums": [C.POINTER(InfoNceProblem), vp_, vp_, vp_, i32_, i32_, vp_, vp_, i32_, vp_],
    "cosmos_infonce_bwd": [C.POINTER(InfoNceProblem), vp_, vp_, f32_, f32_, f32_, f32_, f32_, vp_, vp_, vp_, vp_, i64_, i32_, vp_],
    "cosmos_infonce_e_bytes": [C.POINTER(InfoNceProblem)],
    "cosmos_infonce_bwd_e_workspace_bytes": [C.POINTER(InfoNceProblem), i32_],
    "cosmos_infonce_fwd_e": [C.POINTER(InfoNceProblem), vp_, vp_, vp_, vp_, vp_, vp_, i64_, i32_, vp_],
    "cosmos_infonce_bwd_e": [C.POINTER(InfoNceProblem), vp_, vp_, vp_, vp_, vp_, f32_, f32_, f32_, f32_, f32_, vp_, vp_, vp_, vp_,
                             i64_, vp_, i64_, i32_, vp_],
    "cosmos_infonce_bwd_e_cols_splits": [C.POINTER(InfoNceProblem), i32_],
    "cosmos_infonce_bwd_e_cols": [C.POINTER(InfoNceProblem), vp_, vp_, vp_, vp_, vp_, f32_, f32_, vp_, i32_, i32_, vp_],
    "cosmos_infonce_bwd_g": [C.POINTER(InfoNceProblem), vp_, vp_, f32_, f32_, f32_, f32_, f32_, vp_, vp_, vp_, vp_, i64_, vp_, i64_,
                             i32_, vp_],
}


class GemmDesc(C.Structure):
    """cosmos_gemm_desc (include/cosmos_b200.h)"""
    _fields_ = ([(n, C.c_void_p) for n in ("a", "b", "d", "bias", "a2", "b2")] +
                [(n, C.c_int32) for n in ("M", "N", "K", "K2")] +
                [(n, C.c_int64) for n in ("lda", "ldb", "ldd", "lda2", "ldb2")] +
                [(n, C.c_int32) for n in ("batch", "batch_in")] +
                [(n, C.c_int64) for n in ("stride_a", "stride_b", "stride_d", "stride_bias", "stride_a2", "stride_b2",
                                          "stride_a_in", "stride_b_in", "stride_d_in", "stride_bias_in", "stride_a2_in", "stride_b2_in")] +
                [(n, C.c_int32) for n in ("a_kmajor", "b_kmajor", "in_dtype", "out_dtype", "splits", "accumulate")] +
                [("alpha", C.c_float), ("reserved", C.c_int32)])


_POOL_SIGS = {
    "cosmos_gemm_ex": [vp_, i32_, vp_],
    "cosmos_gemm": [vp_, vp_, vp_, vp_, i32_, i32_, i32_, i64_, i64_, i64_, i32_, i32_, i32_, i32_, i32_, f32_, i32_, vp_],
    "cosmos_gemm_batched": [vp_, vp_, vp_, vp_, i32_, i32_, i32_, i64_, i64_, i64_, i32_, i64_, i64_, i64_, i64_, i32_, i32_, i32_,
                            i32_, i32_, i32_, f32_, vp_, vp_, i32_, i64_, i64_, i64_, i64_, i32_, vp_],
    "cosmos_colsoftmax_fwd": [vp_, i64_, i32_, vp_, i64_, i32_, i32_, i32_, i32_, i32_, i32_, i32_, vp_],
    "cosmos_colsoftmax_bwd": [vp_, i64_, i32_, vp_, i64_, i32_, vp_, i64_, i32_, i32_, i32_, i32_, i32_, i32_, vp_],
    "cosmos_layernorm_fwd": [vp_, i32_, vp_, vp_, vp_, i32_, vp_, vp_, i64_, i32_, f32_, i32_, vp_],
    "cosmos_layernorm_bwd": [vp_, i32_, vp_, i32_, vp_, vp_, vp_, vp_, i32_, i32_, vp_, vp_, i64_, i32_, i32_, vp_],
    "cosmos_attn_core_fwd": [vp_, vp_, vp_, vp_, i32_, i32_, i32_, i32_, i32_, i32_, i64_, i64_, i32_, vp_],
    "cosmos_attn_core_bwd": [vp_, vp_, vp_, vp_, vp_, vp_, i32_, i32_, i32_, i32_, i32_, i32_, i64_, i64_, i32_, vp_],
    "cosmos_addnorm_fwd": [vp_, i32_, vp_, vp_, vp_, i64_, i32_, i32_, vp_],
    "cosmos_addnorm_bwd": [vp_, vp_, i32_, vp_, vp_, vp_, i32_, i64_, i32_, i32_, vp_],
    "cosmos_colsum": [vp_, i32_, vp_, i64_, i32_, i64_, i32_, vp_],
    "cosmos_retrieval_ranks": [vp_, vp_, i32_, i32_, i32_, i32_, i64_, i64_, vp_, vp_, vp_, vp_, vp_, i32_, vp_],
}


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"cosmos_b200: CUDA extension not built ({LIB_PATH} missing). Run `python -m cosmos_b200.build` "
                "(needs nvcc); there is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        _declare(handle)
        if handle.cosmos_abi_version() != ABI_VERSION:
            raise RuntimeError("cosmos_b200: libcosmos_b200.so ABI version mismatch; rebuild with cosmos_b200.build")
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().cosmos_status_string(status).decode()
        if status == 3:
            code = lib().cosmos_last_cuda_error()
            msg += f" [cuda error {code}: {lib().cosmos_cuda_error_string(code).decode()}]"
        raise RuntimeError(f"cosmos_b200: {what} failed: {msg} (status {status})")
    if _DEBUG_SYNC:
        import torch
        torch.cuda.synchronize()


def torch_dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return DTYPE_F32
    if dtype == torch.bfloat16:
        return DTYPE_BF16
    if dtype == torch.float16:
        return DTYPE_F16
    raise RuntimeError(f"cosmos_b200: unsupported dtype {dtype}")


def require_cuda(t, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"cosmos_b200: {what} must live on a CUDA device (got {t.device}); there is no CPU fallback")
