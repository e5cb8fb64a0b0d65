import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from u2tokenizer_b200 import ops
Fr, P, Hd, K = 32, 2048, 768, 1024
vol = torch.rand(Fr, 32, 256, 256, device="cuda")
w = (torch.randn(Hd, K, device="cuda") * K ** -0.5).bfloat16()
b = torch.randn(Hd, device="cuda")
pos = torch.randn(P, Hd, device="cuda").bfloat16()
out = torch.empty(Fr, 2056, Hd, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.patch_embed(vol, [4, 16, 16], w, b, pos, out)
torch.cuda.synchronize()
