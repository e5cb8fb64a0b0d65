"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from u2tokenizer_b200 import parallel


def test_shard_range_partitions_everything():
    for n in (0, 1, 4, 7, 16):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                seen += list(parallel.shard_range(n, r, w))
            assert seen == list(range(n))
            sizes = [len(parallel.shard_range(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert parallel.world() == (rank, world_size)
        imgs = torch.arange(6 * 3, dtype=torch.float32).view(6, 3)
        ids = torch.arange(6).view(6, 1)
        my_imgs, my_ids = parallel.shard_batch([imgs, ids], rank, world_size)
        assert my_imgs.shape[0] == 3 and my_ids[0, 0].item() == 3 * rank
        ms = parallel.max_over_ranks(10.0 + rank)            # slowest rank defines the step time
        got = parallel.gather_ids(my_ids.view(-1) + 100 * rank)
        q.put((rank, ms, [g.tolist() for g in got]))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, gathered in res:
        assert ms == 11.0
        assert gathered == [[0, 1, 2], [103, 104, 105]]


# ------------------------------------------------------------------------------------------------
# ZeRO-1 gradient exchange (host logic of the training step)
# ------------------------------------------------------------------------------------------------
def test_flat_layout_and_buckets():
    lay = parallel.FlatLayout([5, 3, 12], world_size=4, align=2)
    assert lay.offsets == [0, 5, 8] and lay.used == 20 and lay.total == 24 and lay.shard == 6
    assert [lay.shard_bounds(r) for r in range(4)] == [(0, 6), (6, 12), (12, 18), (18, 24)]
    ts = [torch.arange(5.0), None, torch.arange(12.0).view(3, 4)]
    flat = lay.flatten(ts, torch.full((24,), 9.0))
    assert flat[:5].tolist() == [0, 1, 2, 3, 4] and flat[5:8].abs().sum() == 0 and flat[20:].abs().sum() == 0
    outs = [torch.empty(5), torch.empty(3), torch.empty(3, 4)]
    lay.unflatten(flat, outs)
    assert torch.equal(outs[2], ts[2]) and outs[1].abs().sum() == 0
    with pytest.raises(ValueError):
        lay.flatten(ts, torch.zeros(23))
    # buckets walk the parameter list backwards; an oversized parameter sits alone
    assert parallel.plan_buckets([4, 4, 4, 4], 8) == [[3, 2], [1, 0]]
    assert parallel.plan_buckets([3, 20, 2, 2], 5) == [[3, 2], [1], [0]]
    assert parallel.plan_buckets([], 8) == []


def _adamw(lr=1e-2, b1=0.9, b2=0.999, eps=1e-8, wd=0.1):
    """torch.optim.AdamW's update on a flat shard (the CPU stand-in for the fused kernel)."""
    def update(p, g, state, step):
        if "m" not in state:
            state["m"], state["v"] = torch.zeros_like(p), torch.zeros_like(p)
        p.mul_(1 - lr * wd)
        state["m"].mul_(b1).add_(g, alpha=1 - b1)
        state["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (state["v"].sqrt() / (1 - b2 ** step) ** 0.5).add_(eps)
        p.addcdiv_(state["m"], denom, value=-lr / (1 - b1 ** step))
    return update


def _make_params():
    g = torch.Generator().manual_seed(0)
    return [torch.randn(7, 3, generator=g), torch.randn(5, generator=g), torch.randn(2, 2, generator=g), torch.randn(11, generator=g)]


def _rank_grads(step, rank):
    g = torch.Generator().manual_seed(100 * step + rank)
    gs = [torch.randn(7, 3, generator=g), torch.randn(5, generator=g), None, torch.randn(11, generator=g)]  # param 2 is unused
    return gs


def _zero1_worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        params = [p.clone().requires_grad_() for p in _make_params()]
        z = parallel.Zero1Step(params, _adamw())
        assert z.master.numel() == z.layout.shard and z.master.dtype == torch.float32
        for step in range(1, 4):
            for p, gr in zip(params, _rank_grads(step, rank)):
                p.grad = gr
            z.step()
        q.put((rank, [p.detach().clone() for p in params], z.state["m"].numel()))
    finally:
        dist.destroy_process_group()


def test_zero1_step_matches_single_process_adamw():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_zero1_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: torch.optim.AdamW on the mean of the two ranks' gradients (unused parameter: zero gradient)
    ref = [p.clone().requires_grad_() for p in _make_params()]
    opt = torch.optim.AdamW(ref, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    for step in range(1, 4):
        g0, g1 = _rank_grads(step, 0), _rank_grads(step, 1)
        for p, a, b in zip(ref, g0, g1):
            p.grad = torch.zeros_like(p) if a is None else (a + b) / 2
        opt.step()
    total = sum(p.numel() for p in ref)
    for rank, params, m_elems in res:
        assert m_elems == (total + 15) // 16 * 16 // 2          # optimiser state: half of the (padded) flat vector per rank
        for got, want in zip(params, ref):
            assert torch.allclose(got, want.detach(), atol=1e-6, rtol=1e-5)
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)                                  # both ranks end with identical parameters
