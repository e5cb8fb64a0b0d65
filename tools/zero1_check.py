"""ZeRO-1 gradient exchange on real GPUs (launch: torchrun --nproc-per-node 2 tools/zero1_check.py).
Every rank trains on its own batch with TrainEngine(world_size = W): bucketed NCCL reduce-scatter overlapped with the
backward, fused AdamW on the local slices, all-gather. Rank 0 replays the same two steps on a single-GPU engine that sees
all W batches (gradient = mean over the batches) and compares parameters and optimizer state."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import tiny_geometry  # noqa: E402
from u2tokenizer_b200.synthetic import synthetic_inputs, synthetic_state_dict  # noqa: E402
from u2tokenizer_b200.train import TrainEngine  # noqa: E402


def batch_for(g, r):
    images, ids, qids = synthetic_inputs(g, batch=2, frames=2, n_question=6, lt=10, seed=100 + r)
    labels = ids.clone()
    labels[:, :g.num_3d_query_token + 1] = -100
    return [t.cuda() for t in (images, ids, qids, labels)]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    g = tiny_geometry()
    sd = synthetic_state_dict(g, seed=4, device="cpu", dtype=torch.bfloat16)
    sd["model.u2tokenizer.query_tokens"] = (sd["model.u2tokenizer.query_tokens"].float() * 50).to(torch.bfloat16)
    te = TrainEngine(g, sd, device="cuda", world_size=world, rank=rank, bucket_elems=40_000)
    te.init_optimizer(lr=1e-2, weight_decay=0.01, max_grad_norm=1.0)
    losses = []
    for _ in range(2):
        te.zero_grad()
        losses.append(float(te.forward_backward(*batch_for(g, rank)[:3], batch_for(g, rank)[3])))
        te.optimizer_step()
    te.sync_params()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ref = TrainEngine(g, sd, device="cuda", world_size=1, rank=0, bucket_elems=40_000)
        ref.init_optimizer(lr=1e-2, weight_decay=0.01, max_grad_norm=1.0)
        for _ in range(2):
            ref.zero_grad()
            for r in range(world):
                b = batch_for(g, r)
                ref.forward_backward(b[0], b[1], b[2], b[3], grad_scale=1.0 / world)
            ref.optimizer_step()
        torch.cuda.synchronize()
        used = te.lay.mat_used
        a, b_ = te.W[:used].float(), ref.W[:used].float()
        dw = (a - b_).abs().max().item()
        scale = b_.abs().max().item()
        va = te.W[te.lay.mat_total:].float()
        vb = ref.W[ref.lay.mat_total:].float()
        dv = (va - vb).abs().max().item()
        # fp32 master weights of rank 0's slices against the single-GPU run: the two-step AdamW displacement (about 2 * lr
        # per element) must agree - a mis-routed bucket would show up as displacements that differ by O(lr)
        lr = te.opt["lr"]
        L = te.lay
        agree = tot = 0
        for i in range(L.n_buckets):
            lo = i * L.bucket
            n = min(L.piece, max(0, used - lo))
            if n <= 0:
                continue
            mine = te.opt["m_master"][i * L.piece:i * L.piece + n]
            theirs = ref.opt["m_master"][lo:lo + n]
            if i == 0:  # flat start values through the layout (8-aligned slots)
                start = torch.zeros(L.mat_total, device="cuda")
                for k in L.mat_names:
                    start[L.mat_off[k]:L.mat_off[k] + L._numel(k)] = sd[k].float().flatten().cuda()
            d_mine, d_ref = mine - start[lo:lo + n], theirs - start[lo:lo + n]
            # elements whose gradient is above the noise floor moved ~lr in the same direction in both steps; the others
            # (k-projection weights behind a nearly uniform softmax, ...) carry cancellation noise of the order of Adam's
            # eps, where the update is not a stable function of the gradient in ANY implementation
            sig = d_ref.abs() >= 1.5 * lr
            good = ((d_mine - d_ref).abs() < 0.25 * lr) & sig
            agree += int(good.sum())
            tot += int(sig.sum())
            n = int(sig.sum())
            if int(good.sum()) < 0.99 * n:
                names = [k for k in L.bucket_names[i] if L.mat_off[k] < lo + n and L.mat_off[k] + L._numel(k) > lo]
                print(f"   bucket {i}: {int(good.sum())}/{n} agree; rank-0 slice holds {names}; "
                      f"|d_mine| {d_mine.abs().max().item():.3g} |d_ref| {d_ref.abs().max().item():.3g}", flush=True)
        frac = agree / max(tot, 1)
        ok = frac > 0.995 and tot > 0.25 * used / world
        print(f"ZERO1 world={world} buckets={te.lay.n_buckets} overlap={te.overlap} losses={losses} "
              f"max|dW|={dw:.3g} (max|W| {scale:.3g}) max|dV|={dv:.3g}; rank-0 master displacement agrees on "
              f"{frac:.3%} of the {tot} elements with a significant gradient -> {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
