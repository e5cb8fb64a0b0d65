"""Timing of the GPU volume preprocessing on a CT-sized study (512 x 512 x 212 voxels, like the reference's example
amos_7284.nii.gz, u2Transform.py:128-131) with CUDA events; prints ms per study and the HBM traffic it implies."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from u2tokenizer_b200 import ops
D, H, W = 212, 512, 512
g = torch.Generator(device="cuda").manual_seed(0)
vol = torch.full((D, H, W), -1000.0, device="cuda")
vol[10:200, 60:450, 40:470] = (torch.rand(190, 390, 430, device="cuda", generator=g) * 1400 - 200).round()
ws = torch.empty(int(__import__("u2tokenizer_b200._lib", fromlist=["x"]).load().u2_preprocess_ws_bytes(D, H, W)), device="cuda", dtype=torch.uint8)
for _ in range(3):
    out, info = ops.preprocess_volume(vol, ws=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out, info = ops.preprocess_volume(vol, ws=ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
n = D * H * W
i = ops.preprocess_info(info)
crop = 1
for a in range(3):
    crop *= i["hi"][a] - i["lo"][a]
# algorithmic traffic: 3 histogram reads + 1 scale+box read and 1 write of the volume, 3 smoothing passes (read + write) over the crop,
# the resize reading the crop once and writing the padded output
by = 4 * n * 4 + n * 4 + 3 * 2 * crop * 4 + crop * 4 + out.numel() * 4
print(f"preprocess {D}x{H}x{W}: {ms:.3f} ms per study; info {i}")
print(f"algorithmic bytes {by / 1e6:.1f} MB -> {by / ms / 1e6:.1f} GB/s")
