"""Quick TFLOP/s probe of u2_gemm_bf16 (not the product bench; used while tuning)."""
import sys
import torch
sys.path.insert(0, ".")
from u2tokenizer_b200 import ops

def run(M, N, K, block_n=0, iters=20):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.linear(a, b, out=out, block_n=block_n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.linear(a, b, out=out, block_n=block_n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K} bn={block_n}: {ms:.3f} ms {fl/ms/1e9:.1f} TF/s | cublas {ms_t:.3f} ms {fl/ms_t/1e9:.1f} TF/s", flush=True)

if __name__ == "__main__":
    for bn in (128, 256):
        run(8192, 8192, 8192, bn)
        run(8192, 4096, 4096, bn)
        run(65568, 2304, 768, bn)
        run(65568, 768, 3072, bn)
        run(1152, 6144, 4096, bn)
