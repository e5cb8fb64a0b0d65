"""ctypes binding of libu2b200.so (declared in include/u2b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised. PyTorch only supplies device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libu2b200.so"
_lib = None


class GemmDesc(C.Structure):
    """Mirror of ``u2_gemm_desc`` (include/u2b200.h)."""
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("zi", C.c_int32), ("zo", C.c_int32), ("b_zi_div", C.c_int32),
        ("lda", C.c_int64), ("a_stride_zi", C.c_int64), ("a_stride_zo", C.c_int64),
        ("ldb", C.c_int64), ("b_stride_zi", C.c_int64), ("b_stride_zo", C.c_int64),
        ("ldc", C.c_int64), ("c_stride_zi", C.c_int64), ("c_stride_zo", C.c_int64),
        ("c_dtype", C.c_int32), ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("res_row_mod", C.c_int32),
        ("row_div", C.c_int32), ("row_stride", C.c_int32), ("row_off", C.c_int32),
        ("block_n", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/u2b200.h declares must be listed here
# (tests/test_abi.py cross-checks this table against the header and the built library).
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "u2_version": (C.c_int, []),
    "u2_last_error": (C.c_char_p, []),
    "u2_device_sm_count": (C.c_int, []),
    "u2_gemm_bf16": (C.c_int, [_P, _P, _P, C.POINTER(GemmDesc), _P]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the library (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m u2tokenizer_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().u2_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
