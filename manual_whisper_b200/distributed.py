"""Multi-GPU plumbing: one process per GPU, windows sharded by batch, host gather (no data-path collective).

VAD windows are independent units (same fixed prompt, no cross-window state: SURVEY.md §8e), so the N>1 path
is: every rank computes the same window list, takes its share of the windows (longest-first bin packing, equal counts),
transcribes them on its own GPU, and the (index, result) pairs are gathered and restored to window order.  torch.distributed is
used only for that final object gather (NCCL or gloo — identical code path, which is what the CPU tests run).
"""
from __future__ import annotations

from typing import Any, List, Sequence, Tuple


def shard_batches(n_windows: int, batch_size: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """[start, end) window ranges owned by `rank`: batches dealt round-robin so every rank gets full batches first."""
    if batch_size <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad sharding arguments")
    batches = [(i, min(i + batch_size, n_windows)) for i in range(0, n_windows, batch_size)]
    return batches[rank::world]


def gather_ordered(local: Sequence[Tuple[int, Any]], n_windows: int, group=None) -> List[Any]:
    """All ranks contribute (window_index, result) pairs; every rank gets the full list in window order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        merged = list(local)
    else:
        bucket = [None] * dist.get_world_size(group)
        dist.all_gather_object(bucket, list(local), group=group)
        merged = [p for part in bucket for p in part]
    out: List[Any] = [None] * n_windows
    for idx, res in merged:
        if out[idx] is not None:
            raise RuntimeError(f"window {idx} was produced by two ranks")
        out[idx] = res
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError(f"windows {missing[:8]} were not produced by any rank")
    return out


def shard_windows(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Window indices owned by each rank: longest-first greedy bin packing (LPT) on window length with equal counts as the
    first criterion - every window costs the same encoder pass and (at the token cap) the same decode steps, so ranks must
    hold the same NUMBER of windows before their audio seconds are balanced.  Deterministic; indices
    of a rank are returned in window order."""
    if world <= 0:
        raise ValueError("bad sharding arguments")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    load = [0] * world
    for i in order:
        r = min(range(world), key=lambda k: (len(shards[k]), load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(sh) for sh in shards]


def sharded_windows(pipeline, audio, chunk_size: int = 30):
    """The window list every rank computes identically (VAD -> merge_chunks), with sample offsets and lengths."""
    import numpy as np
    import torch
    from .vad import merge_chunks
    from .config import SAMPLE_RATE
    turns = pipeline.vad_model({"waveform": torch.from_numpy(audio).unsqueeze(0), "sample_rate": SAMPLE_RATE})
    windows = merge_chunks(turns, chunk_size, onset=pipeline._vad_params["vad_onset"], offset=pipeline._vad_params["vad_offset"])
    offs = np.clip(np.array([int(w["start"] * SAMPLE_RATE) for w in windows], dtype=np.int64), 0, len(audio))
    ends = np.clip(np.array([int(w["end"] * SAMPLE_RATE) for w in windows], dtype=np.int64), offs, len(audio))
    return windows, offs, (ends - offs)


def transcribe_sharded(pipeline, audio, batch_size: int, rank: int, world: int, group=None, **kwargs):
    """model.transcribe across `world` single-GPU processes: same return value on every rank.

    Every rank computes the same window list, takes its LPT share (shard_windows), uploads the span of audio its windows
    cover ONCE and hands all of its batches to the replicas of its GPU in one call (so the streams_per_device batches in
    flight are used exactly as in the single-process path); only the final (index, segment) gather crosses ranks."""
    import numpy as np
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    windows, offs, lens = sharded_windows(pipeline, audio, kwargs.pop("chunk_size", 30))
    mine = shard_windows(lens, world)[rank]
    local = []
    language = kwargs.get("language")
    if mine:
        res = pipeline.transcribe_windows_host(audio, [windows[i] for i in mine], batch_size=batch_size, **kwargs)
        local = list(zip(mine, res))
    if pipeline.tokenizer is not None:
        language = pipeline.tokenizer.language_code
    segs = gather_ordered(local, len(windows), group)
    if pipeline.preset_language is None:
        pipeline.tokenizer = None
    return {"segments": segs, "language": language or pipeline.preset_language}


def align_sharded(transcript, model_a, metadata, audio, rank: int, world: int, group=None, *, batch_size: int = 16,
                  align_fn=None, **kwargs):
    """whisperx.align across `world` single-GPU processes (the step after the path): segments are
    independent units, so batches of `batch_size` segments are dealt round-robin, every rank aligns its share on its own
    GPU, and the per-segment results are gathered in segment order.  Same return value on every rank."""
    from .alignment import align
    align_fn = align_fn or align
    transcript = list(transcript)
    local = []
    for a, b in shard_batches(len(transcript), batch_size, world, rank):
        # one call per batch keeps the GPU batching; a segment may come back as several sentence sub-segments, so each
        # input carries its index through the call and the outputs are regrouped by it
        tagged = [dict(seg, _idx=i) for i, seg in zip(range(a, b), transcript[a:b])]
        per_seg = {i: [] for i in range(a, b)}
        out = align_fn(tagged, model_a, metadata, audio, _keep_index=True, **kwargs)["segments"]
        for s_ in out:
            per_seg[s_.pop("_idx")].append(s_)
        local += [(i, per_seg[i]) for i in range(a, b)]
    groups = gather_ordered(local, len(transcript), group) if transcript else []
    segments = [s for g in groups for s in g]
    return {"segments": segments, "word_segments": [w for s in segments for w in s["words"]]}
