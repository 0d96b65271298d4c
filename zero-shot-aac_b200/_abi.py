"""ctypes binding of libzsaac_b200.so — the C ABI declared in include/zsaac.h.

This is the only place the Python host touches native code.  There is no fallback: if the
shared library is missing, or the machine has no sm_100 GPU, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading

_LIB_NAME = "libzsaac_b200.so"
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", _LIB_NAME)

ZS_OK = 0
ZS_ERR_INVALID = -1
ZS_ERR_CUDA = -2
ZS_ERR_NO_DEVICE = -3
ZS_ERR_STATE = -4
ZS_ERR_KERNEL = -5

ZS_F32 = 0
ZS_BF16 = 1
ZS_PASS_K = 32      # list length per pass of the fused kernel
ZS_MAX_K = 1024     # k > 32 runs ceil(k / 32) passes over the bank
ZS_DIM_MULTIPLE = 64

_c_ctx = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_ptr = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/zsaac.h one to one (tests/test_host_cpu.py checks it)
SIGNATURES = {
    "zs_abi_version": (_int, []),
    "zs_last_error": (ctypes.c_char_p, []),
    "zs_kernel_name": (ctypes.c_char_p, []),
    "zs_create": (_int, [ctypes.POINTER(_c_ctx), _int]),
    "zs_destroy": (_int, [_c_ctx]),
    "zs_bank_alloc": (_int, [_c_ctx, _i64, _int]),
    "zs_bank_upload": (_int, [_c_ctx, _ptr, _i64, _i64, _int, _int, _ptr]),
    "zs_bank_window": (_int, [_c_ctx, _i64, _i64]),
    "zs_normalize_rows_f32": (_int, [_c_ctx, _ptr, _ptr, _i64, _int, _ptr]),
    "zs_bank_rows": (_i64, [_c_ctx]),
    "zs_bank_dim": (_int, [_c_ctx]),
    "zs_reserve": (_int, [_c_ctx, _i64, _int]),
    "zs_search": (_int, [_c_ctx, _ptr, _i64, _int, _int, _int, _ptr, _i64, _ptr, _ptr, _ptr]),
    "zs_rank_count": (_int, [_c_ctx, _ptr, _i64, _int, _int, _ptr, _int, _i64, _ptr, _ptr, _ptr]),
    "zs_memory_project": (_int, [_c_ctx, _ptr, _i64, _ptr, _i64, _int, ctypes.c_float, _ptr, _ptr]),
    "zs_memory_bank_prepare": (_int, [_c_ctx, _ptr, _i64, _int, _ptr]),
    "zs_memory_project_batched": (_int, [_c_ctx, _ptr, _i64, ctypes.c_float, _ptr, _ptr]),
    "zs_merge": (_int, [_c_ctx, _ptr, _ptr, _int, _i64, _i64, _i64, _int, _ptr, _ptr, _ptr]),
    "zs_rescore_f32": (_int, [_c_ctx, _ptr, _i64, _int, _ptr, _i64, _int, _i64, _ptr, _int, _int,
                              _ptr, _ptr, _ptr]),
    "zs_exact_topk_f32": (_int, [_c_ctx, _ptr, _i64, _ptr, _i64, _int, _int, _int, _ptr, _i64, _ptr,
                                 _ptr, _ptr]),
    "zs_exact_rank_f32": (_int, [_c_ctx, _ptr, _i64, _ptr, _i64, _int, _int, _ptr, _int, _i64, _ptr,
                                 _ptr, _ptr]),
    "zs_gather_rows_f32": (_int, [_c_ctx, _ptr, _i64, _int, _ptr, _i64, _ptr, _ptr]),
    "zs_plan": (_int, [_c_ctx, _i64, _int, ctypes.POINTER(_int), ctypes.POINTER(_int),
                       ctypes.POINTER(_int)]),
    "zs_plan_dry": (_int, [_int, _i64, _i64, _int, _int, ctypes.POINTER(_int), ctypes.POINTER(_int),
                           ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "zs_launch_count": (_i64, [_c_ctx]),
    "zs_kernel_error": (_int, [_c_ctx]),
    "zs_profile_enable": (_int, [_c_ctx, _int]),
    "zs_profile_read": (_int, [_c_ctx, ctypes.POINTER(ctypes.c_float), _int, ctypes.POINTER(_int)]),
    "zs_debug_scores": (_int, [_c_ctx, _ptr, _i64, _int, _int, _ptr, _ptr]),
    "zs_debug_trace": (_int, [_c_ctx, _ptr]),
}

_lib = None
_lock = threading.Lock()


class ZsaacError(RuntimeError):
    """A non-zero status from the native library (message from zs_last_error)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libzsaac_b200 status {status}: {message}")
        self.status = status


def library_path() -> str:
    return os.environ.get("ZSAAC_B200_LIB", _LIB_PATH)


def load_library() -> ctypes.CDLL:
    """Load the shared library (once).  Raises OSError with build instructions if absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise OSError(
                f"{path} not found: the CUDA extension is not built.  Run `make lib` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`) at the repo root.  "
                "There is no CPU fallback for this path.")
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(status: int) -> None:
    if status == ZS_OK:
        return
    msg = load_library().zs_last_error()
    text = msg.decode("utf-8", "replace") if msg else ""
    # like torch, argument errors (k out of range, bad shapes) surface as RuntimeError
    raise ZsaacError(status, text)
