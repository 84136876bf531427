"""N>1 path on CPU: two gloo ranks shard the window list, each 'transcribes' its share, host gather restores order."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manual_whisper_b200.distributed import shard_batches, gather_ordered


def test_shard_batches_partition():
    for n, bs, world in [(136, 32, 2), (136, 32, 8), (5, 32, 4), (0, 16, 2), (64, 16, 3)]:
        seen = []
        for r in range(world):
            seen += [i for a, b in shard_batches(n, bs, world, r) for i in range(a, b)]
        assert sorted(seen) == list(range(n))
    assert shard_batches(136, 32, 2, 1) == [(32, 64), (96, 128)]
    with pytest.raises(ValueError):
        shard_batches(10, 4, 2, 2)


def _worker(rank, world, port, n, bs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = []
        for a, b in shard_batches(n, bs, world, rank):
            for i in range(a, b):
                local.append((i, {"text": f"w{i}", "rank": rank, "tokens": [i, i + 1]}))
        out = gather_ordered(local, n)
        ok = [o["text"] for o in out] == [f"w{i}" for i in range(n)]
        ranks = sorted({o["rank"] for o in out})
        q.put((rank, ok, ranks))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_restores_window_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 70, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(ranks == [0, 1] for _, _, ranks in res)


def test_gather_detects_missing_and_duplicate_windows():
    with pytest.raises(RuntimeError, match="not produced"):
        gather_ordered([(0, "a")], 2)
    with pytest.raises(RuntimeError, match="two ranks"):
        gather_ordered([(0, "a"), (0, "b")], 1)
