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
