"""Python-side operator wrappers: torch tensors in, C-ABI calls (libu2b200.so) underneath.

Every function here launches hand-written sm_100a kernels on the current torch CUDA stream; none
of them has a PyTorch fallback. Shapes/strides are validated here, arithmetic happens in csrc/.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmDesc

BF16 = torch.bfloat16
F32 = torch.float32

ACT_NONE, ACT_GELU, ACT_SILU = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("u2tokenizer_b200 ops run on CUDA tensors only (no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, *, M: int, N: int, K: int,
         lda: int, ldb: int, ldc: int,
         zi: int = 1, zo: int = 1, b_zi_div: int = 1,
         a_strides=(0, 0), b_strides=(0, 0), c_strides=(0, 0),
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         residual: Optional[torch.Tensor] = None, ldr: int = 0, res_row_mod: int = 0,
         row_remap=(0, 0, 0), block_n: int = 0) -> torch.Tensor:
    """Raw (batched, strided) GEMM: C[z] = act(alpha * A[z] @ B[z']^T + bias) + residual."""
    _need_cuda(a, b, c, bias, residual)
    if a.dtype != BF16 or b.dtype != BF16:
        raise TypeError("gemm operands must be bf16")
    if c.dtype not in (BF16, F32):
        raise TypeError("gemm output must be bf16 or fp32")
    if bias is not None and bias.dtype != F32:
        raise TypeError("gemm bias must be fp32")
    if residual is not None and residual.dtype != BF16:
        raise TypeError("gemm residual must be bf16")
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.zi, d.zo, d.b_zi_div = zi, zo, b_zi_div
    d.lda, d.a_stride_zi, d.a_stride_zo = lda, a_strides[0], a_strides[1]
    d.ldb, d.b_stride_zi, d.b_stride_zo = ldb, b_strides[0], b_strides[1]
    d.ldc, d.c_stride_zi, d.c_stride_zo = ldc, c_strides[0], c_strides[1]
    d.c_dtype = DT_BF16 if c.dtype == BF16 else DT_F32
    d.alpha = alpha
    d.bias = _ptr(bias)
    d.act = act
    d.residual = _ptr(residual)
    d.ldr = ldr
    d.res_row_mod = res_row_mod
    d.row_div, d.row_stride, d.row_off = row_remap
    d.block_n = block_n
    lib = _lib.load()
    _lib.check(lib.u2_gemm_bf16(a.data_ptr(), b.data_ptr(), c.data_ptr(), C.byref(d), _stream()),
               "u2_gemm_bf16")
    return c


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = BF16,
           alpha: float = 1.0, block_n: int = 0) -> torch.Tensor:
    """y = act(x @ w^T + bias) + residual for x [..., K] (last dim contiguous), w [N, K]."""
    K = x.shape[-1]
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"linear: weight {tuple(w.shape)} does not match input K={K}")
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    if w.stride(1) != 1:
        w = w.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=out_dtype)
    o2 = out.view(-1, N) if out.is_contiguous() else out
    r2 = None
    ldr = 0
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(1) != 1:
            r2 = r2.contiguous()
        ldr = r2.stride(0)
    gemm(x2, w, o2, M=M, N=N, K=K, lda=x2.stride(0), ldb=w.stride(0), ldc=o2.stride(0),
         alpha=alpha, bias=bias, act=act, residual=r2, ldr=ldr, block_n=block_n)
    return out
