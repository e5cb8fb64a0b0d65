"""In-kernel timeline (globaltimer stamps) of the 4-op decode chain launch at Qwen3-8B shapes."""
import os, sys, torch
sys.path.insert(0, ".")
from u2tokenizer_b200 import ops
B, E, I, NQ = 4, 4096, 12288, 6144
dev = "cuda"
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc)
nlay = 6
W = [dict(wo=rnd(E, E, sc=E ** -0.5).bfloat16(), wgu=rnd(2 * I, E, sc=E ** -0.5).bfloat16(), wdn=rnd(E, I, sc=I ** -0.5).bfloat16(),
          wqkv=rnd(NQ, E, sc=E ** -0.5).bfloat16()) for _ in range(nlay)]
ln = 1 + 0.1 * rnd(E)
tiles = (2 * I + 127) // 128
ws = torch.zeros(max(ops.dlinear_ws_elems(n, k) for n, k in ((E, E), (2 * I, E), (E, I), (NQ, E))), device=dev); cnt = torch.zeros(tiles * 2, device=dev, dtype=torch.int32)
gridbar = torch.zeros(4 * nlay, device=dev, dtype=torch.int32); step = torch.zeros(1, device=dev, dtype=torch.int32)
ssq_a, ssq_b = torch.zeros(16, device=dev), torch.zeros(16, device=dev)
x = rnd(B, E).bfloat16(); xg = torch.empty_like(x); ctx = rnd(B, E).bfloat16()
act = torch.empty(B, I, device=dev, dtype=torch.bfloat16); qkv = torch.empty(B, NQ, device=dev, dtype=torch.bfloat16)
dbg = torch.zeros(148 * 4 * 8, device=dev, dtype=torch.int64)
def chain(l, d=None):
    w = W[l]; c = dict(ws=ws, counters=cnt, sched=int(os.environ.get("U2_DL_SCHED", "0")))
    return [(ctx, w["wo"], x, dict(residual=x, gamma_next=ln, xg=xg, ssq_out=ssq_a, ssq_zero=ssq_b, dbg=d, **c)),
            (xg, w["wgu"], act, dict(ssq_in=ssq_a, silu_pair=True, **c)),
            (act, w["wdn"], x, dict(residual=x, gamma_next=ln, xg=xg, ssq_out=ssq_b, ssq_zero=ssq_a, **c)),
            (xg, w["wqkv"], qkv, dict(ssq_in=ssq_b, **c))]
for it in range(3):
    step += 1
    for l in range(nlay):
        ops.dlinear_multi(chain(l), gridbar=gridbar[4 * l:4 * l + 4], step_dev=step)
torch.cuda.synchronize()
import os
LA = int(os.environ.get("U2_L2_LOOKAHEAD", "24")); NX = int(os.environ.get("U2_L2_NEXT", "20"))
nxw = lambda l: ((W[(l + 1) % nlay]["wo"], 1 << 20), (W[(l + 1) % nlay]["wgu"], NX)) if NX >= 0 else ()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
step += 1
e0.record()
for l in range(nlay):
    ops.dlinear_multi(chain(l), gridbar=gridbar[4 * l:4 * l + 4], step_dev=step, lookahead_units=LA, next_weights=nxw(l))
e1.record(); torch.cuda.synchronize()
print(f"chain launch: {e0.elapsed_time(e1) * 1e3 / nlay:.1f} us each (stream-only bound {2 * (E * E + 2 * I * E + E * I + NQ * E) / 6.4e6:.1f} us)")
step += 1
ops.dlinear_multi(chain(2, dbg), gridbar=gridbar[8:12], step_dev=step, lookahead_units=LA, next_weights=nxw(2)); torch.cuda.synchronize()
d = dbg.view(148, 4, 8).cpu()
t0 = d[:, 0, 0].min().item()
rel = (d - t0).float() / 1e3
names = ["Wpre", "dep ok", "1st full", "last commit", "last acc", "epi done"]
for oi, on in enumerate(["o_proj", "gate_up", "down", "qkv"]):
    print(f"{on:8s} " + " | ".join(f"{n} {rel[:, oi, i].min():.1f}/{rel[:, oi, i].median():.1f}/{rel[:, oi, i].max():.1f}" for i, n in enumerate(names)))
