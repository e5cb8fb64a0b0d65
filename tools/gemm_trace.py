"""Per-shape GEMM times inside one training step (CUDA events around every ops.gemm call, warm, in stream order):
which contraction classes of the step sit how far from the tensor peak.
usage: python tools/gemm_trace.py [workload]"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from u2tokenizer_b200 import ops  # noqa: E402
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
phase_of = {}
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        ops.GEMM_TRACE = []
    te.zero_grad()
    loss = te.forward_loss(images, ids, qids, labels)
    n_fwd = len(ops.GEMM_TRACE) if ops.GEMM_TRACE is not None else 0
    te.backward()
    te.optimizer_step()
    torch.cuda.synchronize()
trace, ops.GEMM_TRACE = ops.GEMM_TRACE, None
agg = defaultdict(lambda: [0, 0.0])
for i, (key, e0, e1) in enumerate(trace):
    k = ("fwd" if i < n_fwd else "bwd",) + key
    agg[k][0] += 1
    agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"loss {float(loss):.4f}; {len(trace)} GEMM calls, {tot:.1f} ms inside the events")
print(f"{'ph':3s} {'M':>7s} {'N':>7s} {'K':>7s} {'z':>5s} aT bT {'out':>4s} b a r e {'n':>4s} {'ms':>8s} {'share':>6s} {'TF/s':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    ph, M, N, K, z, amn, bmn, od, bi, ac, rs, ep = k
    fl = 2.0 * M * N * K * z * n
    print(f"{ph:3s} {M:7d} {N:7d} {K:7d} {z:5d} {amn:2d} {bmn:2d} {od:>4s} {bi} {ac} {rs} {ep} {n:4d} {ms:8.2f} {ms / tot:6.1%} {fl / ms / 1e9:7.0f}")
