import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from u2tokenizer_b200 import ops
Fr, P, Hd, K = 32, 2048, 768, 1024
vols = [torch.rand(Fr, 32, 256, 256, device="cuda") for _ in range(2)]
w = (torch.randn(Hd, K, device="cuda") * K ** -0.5).bfloat16()
b = torch.randn(Hd, device="cuda")
pos = torch.randn(P, Hd, device="cuda").bfloat16()
out = torch.empty(Fr, 2056, Hd, device="cuda", dtype=torch.bfloat16)
for i in range(3):
    ops.patch_embed(vols[i % 2], [4, 16, 16], w, b, pos, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(8):
    ops.patch_embed(vols[i % 2], [4, 16, 16], w, b, pos, out)
e1.record()
torch.cuda.synchronize()
print("U2_PE_DBG", os.environ.get("U2_PE_DBG", "0"), round(e0.elapsed_time(e1) * 1e3 / 8, 1), "us")
