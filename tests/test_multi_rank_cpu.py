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


def _fake_align(transcript, model, metadata, audio, _keep_index=False, **kw):
    """Stands in for the GPU alignment: splits every segment into one sub-segment per sentence-ish chunk."""
    segs = []
    for seg in transcript:
        parts = seg["text"].split(".")
        for k, p in enumerate(x for x in parts if x):
            s = {"text": p, "start": seg["start"] + k, "end": seg["start"] + k + 1, "words": [{"word": w} for w in p.split()]}
            if _keep_index:
                s["_idx"] = seg.get("_idx")
            segs.append(s)
    return {"segments": segs, "word_segments": [w for s in segs for w in s["words"]]}


def _align_worker(rank, world, port, q):
    from manual_whisper_b200.distributed import align_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        transcript = [{"text": f"a{i} b{i}.c{i}" if i % 3 == 0 else f"x{i}", "start": 10.0 * i, "end": 10.0 * i + 5} for i in range(11)]
        got = align_sharded(transcript, None, {}, None, rank, world, batch_size=2, align_fn=_fake_align)
        want = _fake_align(transcript, None, {}, None)
        q.put((rank, got == want, len(got["segments"])))
    finally:
        dist.destroy_process_group()


def test_two_rank_align_sharded_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_align_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(n == 15 for _, _, n in res)
