"""One cfg 4 training step under the CUDA profiler range (for `ncu --profile-from-start off` launch lists).
usage: python tools/train_profile.py [workload] [no-opt]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from u2tokenizer_b200.synthetic import synthetic_state_dict  # noqa: E402
from u2tokenizer_b200.train import TrainEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cfg, geom, spec = bench.make_geometry(wl)
sd = synthetic_state_dict(geom, seed=0, device="cuda", dtype=torch.bfloat16)
te = TrainEngine(geom, sd, device="cuda")
del sd
torch.cuda.empty_cache()
te.init_optimizer(lr=4e-6, moment_dtype=torch.bfloat16)
images, ids, qids, labels, mask = [t.cuda() for t in bench.train_batch(geom, spec, 0, 1)]
for it in range(2):
    if it == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    te.zero_grad()
    loss = te.forward_loss(images, ids, qids, labels)
    te.backward()
    te.optimizer_step()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
