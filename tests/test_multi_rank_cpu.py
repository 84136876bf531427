"""N>1 path on CPU: two gloo ranks shard the window list, each 'transcribes' its share, host gather restores order."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manual_whisper_b200.distributed import shard_batches, shard_windows, gather_ordered


def test_shard_batches_partition():
    for n, bs, world in [(136, 32, 2), (136, 32, 8), (5, 32, 4), (0, 16, 2), (64, 16, 3)]:
        seen = []
        for r in range(world):
            seen += [i for a, b in shard_batches(n, bs, world, r) for i in range(a, b)]
        assert sorted(seen) == list(range(n))
    assert shard_batches(136, 32, 2, 1) == [(32, 64), (96, 128)]
    with pytest.raises(ValueError):
        shard_batches(10, 4, 2, 2)


def test_shard_windows_is_a_balanced_partition():
    import numpy as np
    rng = np.random.default_rng(0)
    for n, world in [(136, 1), (136, 2), (136, 4), (136, 8), (5, 8), (0, 2), (17, 3)]:
        lens = rng.integers(16000, 480001, size=n)
        shards = shard_windows(lens, world)
        assert len(shards) == world and sorted(i for sh in shards for i in sh) == list(range(n))
        assert all(sh == sorted(sh) for sh in shards)
        counts = [len(sh) for sh in shards]
        assert max(counts) - min(counts) <= 1                               # equal work first: every window costs the same decode
        if n >= 4 * world:
            loads = [int(lens[sh].sum()) for sh in shards]
            assert max(loads) - min(loads) <= 480000                        # then audio seconds, within one window
        assert shards == shard_windows(lens, world)                         # deterministic: every rank computes the same table


class _FakePipeline:
    """CPU stand-in with the attributes transcribe_sharded touches; a window's 'ids' are a function of its audio."""
    preset_language = None
    _vad_params = {"vad_onset": 0.5, "vad_offset": 0.363}

    def __init__(self, turns):
        from manual_whisper_b200.vad import InjectedVad
        self.vad_model = InjectedVad(turns)
        self.tokenizer = None
        self.calls = 0

    def transcribe_windows_host(self, audio, windows, batch_size=None, language=None, **kw):
        import types
        self.calls += 1
        self.tokenizer = types.SimpleNamespace(language_code=language or "detected-en")
        out = []
        for w in windows:
            a = audio[int(w["start"] * 16000): int(w["end"] * 16000)]
            out.append({"text": "", "start": round(w["start"], 3), "end": round(w["end"], 3), "tokens": [len(a), int(abs(a).sum() * 1e3) % 50000]})
        return out


def _sharded_worker(rank, world, port, q):
    from manual_whisper_b200.distributed import transcribe_sharded
    from manual_whisper_b200.vad import synthetic_speech
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        audio, turns = synthetic_speech(400.0, seed=4)
        pipe = _FakePipeline(turns)
        got = transcribe_sharded(pipe, audio, 4, rank, world)
        single = _FakePipeline(turns)
        from manual_whisper_b200.distributed import sharded_windows
        windows, _, _ = sharded_windows(single, audio)
        want = single.transcribe_windows_host(audio, windows)
        q.put((rank, got["segments"] == want, got["language"], pipe.calls, pipe.tokenizer is None, len(want)))
    finally:
        dist.destroy_process_group()


def test_two_rank_transcribe_sharded_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, language, calls, tok_reset, n in res:
        assert same and n >= 10
        assert language == "detected-en"           # the language actually used, not None, when it was detected
        assert calls == 1                          # ONE dispatch per rank over all of its windows (one upload, all streams busy)
        assert tok_reset                           # language=None pipelines forget the tokenizer after the call, as upstream


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
