"""Build the cfg3 model, prefill, run a few eager decode steps (for ncu captures of the decode kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from u2tokenizer_b200 import ops
wl = os.environ.get("U2_PROBE_WORKLOAD", "cfg3")
cfg, geom, spec = bench.make_geometry(wl)
model = bench.build_model(cfg, geom)
eng = model.engine()
B = spec["batch"]
L = geom.num_3d_query_token + spec["n_question"]
emb = (torch.randn(B, L, geom.hidden_size, device="cuda") * 0.02).bfloat16()
cache = eng.new_cache(B, L + 16)
eng.prefill(emb, cache)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("U2_PROBE_STEPS", "2"))):
    eng.decode_step(cache)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
