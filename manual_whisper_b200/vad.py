"""VAD window merging (host) — mirror of ``whisperx.vads.Vad.merge_chunks`` plus injection points.

The reference passes ``vad_options={"vad_onset": 0.5, "vad_offset": 0.363}`` to ``whisperx.load_model``
(/root/reference/transcribe.py:43-46,112).  Upstream the speech turns come from the pyannote
segmentation network, whose weights are not available offline (SURVEY.md §2.2 "VAD": network out of
scope, merge logic on the path).  ``load_model(vad_model=...)`` accepts any callable
``f({"waveform": Tensor[1, N], "sample_rate": 16000}) -> [(start_s, end_s), ...]`` (or objects with
``.start`` / ``.end``); ``EnergyVad`` is the built-in stand-in used when nothing is injected.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

from .config import SAMPLE_RATE


@dataclass
class SegmentX:
    """whisperx.vads.vad.SegmentX: a speech turn."""
    start: float
    end: float
    speaker: str = "UNKNOWN"


def _as_pairs(segments) -> List[Tuple[float, float]]:
    out = []
    for s in segments or []:
        if isinstance(s, dict):
            out.append((float(s["start"]), float(s["end"])))
        elif hasattr(s, "start") and hasattr(s, "end"):
            out.append((float(s.start), float(s.end)))
        else:
            a, b = s
            out.append((float(a), float(b)))
    return out


def merge_chunks(segments, chunk_size: float = 30.0, onset: float = 0.5, offset: float = None) -> List[Dict]:
    """Greedy left-to-right merge of speech turns into windows of at most `chunk_size` seconds
    (SURVEY.md A.4).  A turn longer than `chunk_size` stays a window of its own (the upstream Binarize
    step is what bounds turn length).  `onset`/`offset` are accepted for signature parity; upstream uses
    them only in the binarisation that precedes this merge."""
    if chunk_size <= 0:
        raise ValueError("chunk_size must be positive")
    segs = _as_pairs(segments)
    if len(segs) == 0:
        print("No active speech found in audio")
        return []
    merged = []
    curr_start = segs[0][0]
    curr_end = 0.0
    seg_idxs: List[Tuple[float, float]] = []
    for s, e in segs:
        if e - curr_start > chunk_size and curr_end - curr_start > 0:
            merged.append({"start": curr_start, "end": curr_end, "segments": seg_idxs})
            curr_start = s
            seg_idxs = []
        curr_end = e
        seg_idxs.append((s, e))
    merged.append({"start": curr_start, "end": curr_end, "segments": seg_idxs})
    return merged


class InjectedVad:
    """Feeds a known list of speech turns (the synthetic generator's ground truth)."""

    def __init__(self, segments: Sequence[Tuple[float, float]]):
        self.segments = _as_pairs(segments)

    def __call__(self, audio: Dict) -> List[Tuple[float, float]]:
        return list(self.segments)


class EnergyVad:
    """Frame-energy VAD with onset/offset hysteresis and a maximum turn duration — a stand-in for the
    pyannote segmentation + Binarize(onset, offset, max_duration) stage.  NOT numerically related to
    pyannote; it only provides the same interface so the pipeline is runnable without gated weights."""

    def __init__(self, vad_onset: float = 0.5, vad_offset: float = 0.363, chunk_size: float = 30.0,
                 frame_s: float = 0.02, min_duration_on: float = 0.1, min_duration_off: float = 0.1):
        self.onset, self.offset = float(vad_onset), float(vad_offset)
        self.max_duration = float(chunk_size)
        self.frame = int(round(frame_s * SAMPLE_RATE))
        self.min_on, self.min_off = min_duration_on, min_duration_off

    def frame_rms(self, audio: Dict) -> np.ndarray:
        wav = audio["waveform"]
        wav = np.asarray(wav.detach().cpu().numpy() if hasattr(wav, "detach") else wav, dtype=np.float32).reshape(-1)
        n = len(wav) // self.frame
        if n == 0:
            return np.zeros(0, np.float32)
        return np.sqrt((wav[: n * self.frame].reshape(n, self.frame) ** 2).mean(axis=1) + 1e-12)

    def __call__(self, audio: Dict) -> List[SegmentX]:
        e = self.frame_rms(audio)
        n = len(e)
        if n == 0:
            return []
        db = 20 * np.log10(e)
        # map energy to a [0, 1] speech score between the noise floor and the loud percentile
        lo, hi = np.percentile(db, 10), np.percentile(db, 95)
        score = np.clip((db - lo) / max(hi - lo, 6.0), 0.0, 1.0)
        fs = self.frame / SAMPLE_RATE
        turns, active, start = [], False, 0.0
        for i, s in enumerate(score):
            t = i * fs
            if not active and s > self.onset:
                active, start = True, t
            elif active and (s < self.offset or t - start >= self.max_duration):
                turns.append([start, t])
                active = s >= self.offset
                start = t
        if active:
            turns.append([start, n * fs])
        merged: List[List[float]] = []
        for a, b in turns:     # fill short gaps, drop blips
            if merged and a - merged[-1][1] < self.min_off and b - merged[-1][0] <= self.max_duration:
                merged[-1][1] = b
            else:
                merged.append([a, b])
        return [SegmentX(a, b) for a, b in merged if b - a >= self.min_on]


class GpuEnergyVad(EnergyVad):
    """EnergyVad whose per-frame RMS is computed on the GPU (mw_frame_rms).  `wants_device = True` tells the pipeline to
    upload the waveform ONCE, run the VAD on the device copy and reuse that copy for the ASR batches (SURVEY.md §8f
    rank 1: the host never revisits the audio); only n/320 floats come back for the hysteresis pass."""
    wants_device = True

    def frame_rms(self, audio: Dict) -> np.ndarray:
        import ctypes as C
        import torch
        from . import _lib
        wav = audio["waveform"]
        if not (hasattr(wav, "is_cuda") and wav.is_cuda):
            wav = torch.as_tensor(np.asarray(wav, dtype=np.float32)).cuda()
        wav = wav.reshape(-1).contiguous()
        n = wav.numel() // self.frame
        if n == 0:
            return np.zeros(0, np.float32)
        out = torch.empty(n, dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _lib.check(_lib.load().mw_frame_rms(wav.data_ptr(), wav.numel(), self.frame, out.data_ptr(),
                                                C.c_void_p(torch.cuda.current_stream(wav.device).cuda_stream)), "mw_frame_rms")
        return out.cpu().numpy()


def synthetic_speech(duration_s: float, seed: int = 1, sr: int = SAMPLE_RATE):
    """SURVEY.md §8(d) C3 generator: speech-like bursts U(2,12) s of AM noise (amp 0.1) separated by gaps
    U(0.2,1.5) s of faint noise (amp 1e-4).  Returns (audio f32 [N], [(start_s, end_s), ...])."""
    rng = np.random.default_rng(seed)
    n = int(round(duration_s * sr))
    audio = (1e-4 * rng.standard_normal(n)).astype(np.float32)
    turns = []
    t = float(rng.uniform(0.2, 1.5))
    while t < duration_s - 0.5:
        dur = float(rng.uniform(2.0, 12.0))
        end = min(t + dur, duration_s)
        i0, i1 = int(t * sr), int(end * sr)
        tt = np.arange(i1 - i0, dtype=np.float32) / sr
        am = 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(2.0, 6.0) * tt + rng.uniform(0, 6.28))
        carrier = np.sin(2 * np.pi * rng.uniform(120.0, 300.0) * tt) + 0.5 * np.sin(2 * np.pi * rng.uniform(600.0, 2500.0) * tt)
        audio[i0:i1] += (0.1 * am * (0.5 * rng.standard_normal(i1 - i0) + 0.5 * carrier)).astype(np.float32)
        turns.append((round(t, 3), round(end, 3)))
        t = end + float(rng.uniform(0.2, 1.5))
    return audio, turns
