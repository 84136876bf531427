"""ORACLE (test infrastructure, not product code) — the CTC forced-alignment dynamic programme of ``whisperx.align``
(/root/reference/transcribe.py:130-135).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

whisperx 3.7.6 is absent from the container, so ``get_trellis`` / ``backtrack`` / ``merge_repeats`` are restated from its
published algorithm [UPSTREAM-MEMORY: whisperx/alignment.py, itself the PyTorch "forced alignment with wav2vec2" tutorial
plus the '*' wildcard for out-of-dictionary characters].  PARITY UNPINNED: the reference holds no fixture for it and no
second implementation of this exact objective is importable (``torchaudio.functional.forced_align`` solves full CTC, where
a token may repeat over frames; the tutorial's trellis scores every "stay" frame as blank).  What IS checked: the trellis
optimum and the backtracked path against brute-force enumeration of every change-time assignment on small cases
(tests/test_oracle_align.py).
One documented deviation: the published ``get_trellis`` seeds the tail of column 0 with +inf
(``trellis[-num_tokens + 1:, 0] = inf``), a device that only stops the backtrack from reaching column 0 too late to fit
the remaining tokens - which the -inf cells already guarantee; it is omitted here and in the CUDA kernel so that every
cell is a plain maximum.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

WILDCARD = -1          # token id of '*' (a character outside the model's dictionary): scores as the best non-blank symbol


def _token_emission(frame: np.ndarray, tok: int, blank: int) -> float:
    if tok != WILDCARD:
        return float(frame[tok])
    m = frame.copy()
    m[blank] = -np.inf
    return float(m.max())


def get_trellis(emission: np.ndarray, tokens: Sequence[int], blank: int = 0) -> np.ndarray:
    """trellis[t, j] = best log-probability of having emitted tokens[0..j] after t frames, with tokens[j] started.
    emission [T, V] log-probabilities; returns [T, N] float32 (N = len(tokens)); row 0 is the start state."""
    T, N = emission.shape[0], len(tokens)
    tr = np.full((T, N), -np.inf, dtype=np.float32)
    tr[0, 0] = 0.0
    tr[1:, 0] = np.cumsum(emission[1:, blank], dtype=np.float32)
    for t in range(T - 1):
        for j in range(1, N):
            stay = np.float32(tr[t, j] + np.float32(emission[t, blank]))
            change = np.float32(tr[t, j - 1] + np.float32(_token_emission(emission[t], tokens[j], blank)))
            tr[t + 1, j] = max(stay, change)
    return tr


@dataclass
class Point:
    token_index: int
    time_index: int
    score: float


def backtrack(trellis: np.ndarray, emission: np.ndarray, tokens: Sequence[int], blank: int = 0) -> Optional[List[Point]]:
    """Walks back from (T-1, N-1): at each frame either the token stayed (blank emitted) or changed (token emitted)."""
    t, j = trellis.shape[0] - 1, trellis.shape[1] - 1
    path = [Point(j, t, float(np.exp(emission[t, blank])))]
    while j > 0:
        if t <= 0:
            return None                                    # more tokens than frames: not alignable
        p_stay = np.float32(emission[t - 1, blank])
        p_change = np.float32(_token_emission(emission[t - 1], tokens[j], blank))
        stayed = np.float32(trellis[t - 1, j] + p_stay)
        changed = np.float32(trellis[t - 1, j - 1] + p_change)
        t -= 1
        if changed > stayed:
            j -= 1
            path.append(Point(j, t, float(np.exp(p_change))))
        else:
            path.append(Point(j, t, float(np.exp(p_stay))))
    while t > 0:
        path.append(Point(j, t - 1, float(np.exp(emission[t - 1, blank]))))
        t -= 1
    return path[::-1]


@dataclass
class Segment:
    label: str
    start: int
    end: int
    score: float


def merge_repeats(path: List[Point], transcript: str) -> List[Segment]:
    """Runs of equal token_index become one character segment [first frame, last frame + 1) with the mean frame score."""
    i1, out = 0, []
    while i1 < len(path):
        i2 = i1
        while i2 < len(path) and path[i1].token_index == path[i2].token_index:
            i2 += 1
        score = sum(path[k].score for k in range(i1, i2)) / (i2 - i1)
        out.append(Segment(transcript[path[i1].token_index], path[i1].time_index, path[i2 - 1].time_index + 1, score))
        i1 = i2
    return out


def frame_tokens(path: List[Point], n_frames: int) -> np.ndarray:
    """token index occupied at every frame (the form the CUDA kernel returns)."""
    out = np.zeros(n_frames, dtype=np.int32)
    for p in path:
        out[p.time_index] = p.token_index
    return out
