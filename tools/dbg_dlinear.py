import sys, torch
sys.path.insert(0, ".")
from u2tokenizer_b200 import ops
DEV = "cuda"
N, K, B = 4096, 12288, 1
g = torch.Generator(device=DEV).manual_seed(1)
x = torch.randn(B, K, device=DEV, generator=g).bfloat16()
w = (torch.randn(N, K, device=DEV, generator=g) * K ** -0.5).bfloat16()
ws = ops.dlinear_new_ws(ops.dlinear_ws_elems(N, K), device=DEV)
cnt = torch.zeros((N + 63) // 64, device=DEV, dtype=torch.int32)
ref = x.float() @ w.float().t()
print("ws elems", ws.numel(), "expected slots", ws.numel() // (32 * 128 * 16))
for it in range(3):
    out = torch.empty(B, N, device=DEV)
    ops.dlinear(x, w, out, ws=ws, counters=cnt, sched=0)
    torch.cuda.synchronize()
    bad = torch.isnan(out[0]).nonzero().view(-1)
    err = (out - ref)[~torch.isnan(out)].abs().max().item()
    wsi = ws.view(torch.int32)
    nonsent = (wsi != -1).sum().item()
    print(f"call {it}: nan rows {bad.numel()} first {bad[:8].tolist()} tiles {sorted(set((bad // 128).tolist()))[:10]} max err(other) {err:.4f} non-sentinel words left {nonsent}")
