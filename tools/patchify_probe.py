import sys, torch
sys.path.insert(0, ".")
from u2tokenizer_b200 import ops
vol = torch.rand(32, 32, 256, 256, device="cuda")
out = ops.patchify(vol, (4, 16, 16))
ref = vol.view(32, 1, 8, 4, 16, 16, 16, 16).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(32 * 2048, 1024).bfloat16()
print("exact", torch.equal(out, ref))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
vols = [torch.rand(32, 32, 256, 256, device="cuda") for _ in range(3)]   # 268 MB each: > L2
outs = [torch.empty(32 * 2048, 1024, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
for i in range(3): ops.patchify(vols[i], (4, 16, 16), out=outs[i])
torch.cuda.synchronize(); e0.record()
for r in range(9): ops.patchify(vols[r % 3], (4, 16, 16), out=outs[r % 3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 9
by = 32 * 32 * 256 * 256 * 6
print(f"patchify 32 frames: {ms * 1e3:.1f} us, {by / ms / 1e6:.0f} GB/s algorithmic ({by / ms / 1e6 / 6581.6:.2%} of measured HBM peak)")
