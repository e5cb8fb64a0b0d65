"""cfg2-shaped vision + prefill forward (for ncu captures of the GEMM / flash-attention kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
cfg, geom, spec = bench.make_geometry("cfg2")
model = bench.build_model(cfg, geom)
from u2tokenizer_b200.synthetic import synthetic_inputs
images, ids, qids = synthetic_inputs(geom, batch=1, frames=4, n_question=32, lt=512)
images, ids, qids = images.cuda(), ids.cuda(), qids.cuda()
model(images=images, input_ids=ids, question_ids=qids)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(images=images, input_ids=ids, question_ids=qids)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
