"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports exactly
the symbols include/u2b200.h declares; the ctypes table mirrors the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "u2b200.h")).read() + open(os.path.join(ROOT, "include", "u2b200_train.h")).read()
    return sorted(set(re.findall(r"U2_API\s+[\w\s\*]+?\b(u2_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "u2_gemm_bf16" in syms and "u2_gemv_bf16" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol():
    from u2tokenizer_b200 import _lib
    if not _lib.lib_path().exists():
        from u2tokenizer_b200 import build
        build.build()
    lib = ctypes.CDLL(str(_lib.lib_path()))
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in u2b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.load().u2_version() == 1


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly off-GPU."""
    import torch
    from u2tokenizer_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        from u2tokenizer_b200.engine import U2Engine
        with pytest.raises(RuntimeError):
            U2Engine(None, {})
