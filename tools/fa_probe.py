import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from u2tokenizer_b200 import ops
b, S, h, dh = 32, 2049, 12, 64
Sp = 2056
qkv = torch.randn(b, Sp, 3, h, dh, device="cuda").bfloat16()
q, k, v = qkv[:, :S, 0], qkv[:, :S, 1], qkv[:, :S, 2]
out = torch.zeros(b, Sp, h * dh, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.flash_attention_d64(q, k, v, out[:, :S], dh ** -0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    ops.flash_attention_d64(q, k, v, out[:, :S], dh ** -0.5)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 5
print(f"flash attention 32 x 12 heads, S = 2049: {us:.1f} us, {4.0 * b * h * S * S * dh / us / 1e6:.1f} TFLOP/s")
