"""Short-K (ViT, K = 768) GEMM shapes: tile width 128 vs 256 vs cuBLAS, timed from CUDA graphs (device time only)."""
import sys
import torch
sys.path.insert(0, ".")
from u2tokenizer_b200 import ops


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps)


for (M, N, K, gelu) in ((8224, 3072, 768, True), (8224, 2304, 768, False), (8224, 768, 3072, False), (65792, 3072, 768, True),
                        (65792, 2304, 768, False), (8192, 4096, 4096, False)):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    line = f"M={M} N={N} K={K} gelu={gelu}:"
    for bn in (64, 128, 256):
        ms = timed(lambda: ops.linear(a, b, bias, act=ops.ACT_GELU if gelu else ops.ACT_NONE, out=out, block_n=bn))
        line += f"  bn{bn} {ms * 1e3:.1f} us {fl / ms / 1e9:.0f} TF"
    ms = timed(lambda: ops.linear(a, b, out=out, block_n=256))
    line += f" | bn256 plain {ms * 1e3:.1f} us"
    ms = timed(lambda: torch.matmul(a, b.t(), out=out))
    line += f" | cublas {ms * 1e3:.1f} us {fl / ms / 1e9:.0f} TF"
    print(line, flush=True)
