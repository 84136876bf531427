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


VAD_BINS, VAD_DB_LO, VAD_DB_HI = 2048, -140.0, 20.0      # csrc/vad.cu


def _hist_percentile(hist: np.ndarray, n: int, q: float) -> float:
    """Lower edge of the histogram bin in which the cumulative count first exceeds q*n (csrc/vad.cu: hist_percentile)."""
    cum = np.cumsum(hist.astype(np.int64))
    idx = int(np.searchsorted(cum.astype(np.float64), q * float(n), side="right"))
    return VAD_DB_HI if idx >= VAD_BINS else VAD_DB_LO + idx * ((VAD_DB_HI - VAD_DB_LO) / VAD_BINS)


class EnergyVad:
    """Frame-energy VAD with onset/offset hysteresis and a maximum turn duration — a stand-in for the
    pyannote segmentation + Binarize(onset, offset, max_duration) stage.  NOT numerically related to
    pyannote; it only provides the same interface so the pipeline is runnable without gated weights.

    Score: dB of the frame RMS mapped to [0, 1] between the 10 % and 95 % points of a 2048-bin histogram over
    [-140, 20] dB (float64 throughout), which is exactly what csrc/vad.cu computes on the device: this class is the host
    twin of ``GpuEnergyVad`` and the two produce identical turns and windows."""

    def __init__(self, vad_onset: float = 0.5, vad_offset: float = 0.363, chunk_size: float = 30.0,
                 frame_s: float = 0.02, min_duration_on: float = 0.1, min_duration_off: float = 0.1):
        self.onset, self.offset = float(vad_onset), float(vad_offset)
        self.max_duration = float(chunk_size)
        self.frame = int(round(frame_s * SAMPLE_RATE))
        self.min_on, self.min_off = float(min_duration_on), float(min_duration_off)

    def frame_rms(self, audio: Dict) -> np.ndarray:
        wav = audio["waveform"]
        wav = np.asarray(wav.detach().cpu().numpy() if hasattr(wav, "detach") else wav, dtype=np.float32).reshape(-1)
        n = len(wav) // self.frame
        if n == 0:
            return np.zeros(0, np.float32)
        return np.sqrt((wav[: n * self.frame].reshape(n, self.frame) ** 2).mean(axis=1) + 1e-12)

    def turns_from_rms(self, e: np.ndarray) -> List[List[float]]:
        n = len(e)
        if n == 0:
            return []
        db = 20.0 * np.log10(e.astype(np.float64))
        bins = np.clip(np.floor((db - VAD_DB_LO) * (VAD_BINS / (VAD_DB_HI - VAD_DB_LO))), 0, VAD_BINS - 1).astype(np.int64)
        hist = np.bincount(bins, minlength=VAD_BINS)
        lo, hi = _hist_percentile(hist, n, 0.10), _hist_percentile(hist, n, 0.95)
        score = np.clip((db - lo) / max(hi - lo, 6.0), 0.0, 1.0)
        fs = self.frame / SAMPLE_RATE
        turns, active, start = [], False, 0.0
        for i, s in enumerate(score.tolist()):
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
        return [[a, b] for a, b in merged if b - a >= self.min_on]

    def __call__(self, audio: Dict) -> List[SegmentX]:
        return [SegmentX(a, b) for a, b in self.turns_from_rms(self.frame_rms(audio))]


class GpuEnergyVad(EnergyVad):
    """The VAD front end on the GPU: per-frame RMS (mw_frame_rms), energy score, Binarize hysteresis,
    gap fill and Vad.merge_chunks (mw_vad_windows) all run on the device copy of the waveform.  `wants_device = True` tells
    the pipeline to upload the waveform ONCE and reuse that copy for the ASR batches; the host never revisits the audio and
    only the finished tables (turns and windows: a few hundred doubles) come back.  ``device_windows`` is what the pipeline
    calls; ``__call__`` (turns only, for API parity with other VAD objects) goes through the same kernels."""
    wants_device = True
    MAX_TURNS, MAX_WINDOWS = 1 << 16, 1 << 14

    def _run(self, audio: Dict, chunk_size: float):
        import ctypes as C
        import torch
        from . import _lib
        lib = _lib.load()
        wav = audio["waveform"]
        if not (hasattr(wav, "is_cuda") and wav.is_cuda):
            wav = torch.as_tensor(np.asarray(wav, dtype=np.float32)).cuda()
        wav = wav.reshape(-1).contiguous()
        n = wav.numel() // self.frame
        if n == 0:
            return np.zeros((0, 2)), np.zeros((0, 2))
        dev = wav.device
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rms = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.check(lib.mw_frame_rms(wav.data_ptr(), wav.numel(), self.frame, rms.data_ptr(), st), "mw_frame_rms")
            scratch = torch.empty(8192 + 16 * self.MAX_TURNS, dtype=torch.uint8, device=dev)
            windows = torch.empty((self.MAX_WINDOWS, 2), dtype=torch.float64, device=dev)
            counts = torch.empty(3, dtype=torch.int32, device=dev)
            _lib.check(lib.mw_vad_windows(rms.data_ptr(), n, self.frame / SAMPLE_RATE, self.onset, self.offset, self.max_duration,
                                          self.min_on, self.min_off, float(chunk_size), scratch.data_ptr(), self.MAX_TURNS,
                                          windows.data_ptr(), self.MAX_WINDOWS, counts.data_ptr(), st), "mw_vad_windows")
            c = counts.cpu().numpy()
            if c[2]:
                raise RuntimeError("mw_vad_windows: more speech turns or windows than the device tables hold")
            turns = scratch[8192: 8192 + 16 * int(c[0])].view(torch.float64).reshape(-1, 2).cpu().numpy()
            return turns, windows[: int(c[1])].cpu().numpy()

    def device_windows(self, audio: Dict, chunk_size: float = 30.0) -> List[Dict]:
        """What merge_chunks(self(audio), chunk_size) returns, computed on the device."""
        turns, wins = self._run(audio, chunk_size)
        if len(wins) == 0:
            print("No active speech found in audio")
            return []
        out = []
        for a, b in wins.tolist():
            inside = [(float(s), float(e)) for s, e in turns.tolist() if s >= a and e <= b]
            out.append({"start": float(a), "end": float(b), "segments": inside})
        return out

    def __call__(self, audio: Dict) -> List[SegmentX]:
        turns, _ = self._run(audio, self.max_duration)
        return [SegmentX(float(a), float(b)) for a, b in turns.tolist()]


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
