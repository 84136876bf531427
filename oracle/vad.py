"""ORACLE (test infrastructure, not product code) — VAD window merging and the synthetic speech generator.

Restates whisperx ``Vad.merge_chunks(segments, chunk_size, onset, offset)``, driven by the
reference's ``vad_options`` (/root/reference/transcribe.py:43-46,112).  The pyannote segmentation network
is out of scope (weights unavailable); speech turns are injected.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def merge_chunks(segments: Sequence[Tuple[float, float]], chunk_size: float = 30.0) -> List[Dict]:
    segs = [(float(s), float(e)) for s, e in segments]
    if not segs:
        return []
    out = []
    curr_start = segs[0][0]
    curr_end = 0.0
    seg_idxs: List[Tuple[float, float]] = []
    for s, e in segs:
        if e - curr_start > chunk_size and curr_end - curr_start > 0:
            out.append({"start": curr_start, "end": curr_end, "segments": seg_idxs})
            curr_start = s
            seg_idxs = []
        curr_end = e
        seg_idxs.append((s, e))
    out.append({"start": curr_start, "end": curr_end, "segments": seg_idxs})
    return out
