"""Where does the patch-embedding GEMM (M = 65536, N = 768, K = 1024) lose its time? Variants of the epilogue / tile width."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from u2tokenizer_b200 import ops

M, N, K, P, Sp = 65536, 768, 1024, 2048, 2056
a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(2)]
w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda")
pos = torch.randn(P, N, device="cuda").bfloat16()
out = [torch.empty(32, Sp, N, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
flat = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(2)]


def t(fn, reps=10):
    for i in range(2):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        fn(r % 2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


fl = 2.0 * M * N * K
for name, fn in [
    ("plain bn=auto", lambda i: ops.gemm(a[i], w, flat[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N)),
    ("plain bn=128", lambda i: ops.gemm(a[i], w, flat[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N, block_n=128)),
    ("plain bn=256", lambda i: ops.gemm(a[i], w, flat[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N, block_n=256)),
    ("bias", lambda i: ops.gemm(a[i], w, flat[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias)),
    ("bias+pos", lambda i: ops.gemm(a[i], w, flat[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, residual=pos, ldr=N, res_row_mod=P)),
    ("bias+pos+remap", lambda i: ops.gemm(a[i], w, out[i], M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, residual=pos, ldr=N,
                                          res_row_mod=P, row_remap=(P, Sp, 1))),
    ("same A (L2-warm operands)", lambda i: ops.gemm(a[0][:8192], w, flat[0][:8192], M=8192, N=N, K=K, lda=K, ldb=K, ldc=N)),
]:
    us = t(fn)
    f = fl if "same A" not in name else fl / 8
    print(f"{name:28s} {us:8.1f} us  {f / us / 1e6:7.1f} TFLOP/s")
c = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
x = torch.randn(8192, 8192, device="cuda").bfloat16()
us = t(lambda i: ops.gemm(x, x, c, M=8192, N=8192, K=8192, lda=8192, ldb=8192, ldc=8192), reps=4)
print(f"{'8192^3':28s} {us:8.1f} us  {2 * 8192 ** 3 / us / 1e6:7.1f} TFLOP/s")
