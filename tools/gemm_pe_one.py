import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from u2tokenizer_b200 import ops
M, N, K, P, Sp = 65536, 768, 1024, 2048, 2056
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda")
pos = torch.randn(P, N, device="cuda").bfloat16()
out = torch.empty(32, Sp, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bias, residual=pos, ldr=N, res_row_mod=P, row_remap=(P, Sp, 1))
torch.cuda.synchronize()
