"""ORACLE (test infrastructure, not product code) — band-limited sinc resampling for the audio-decode "next" row
(/root/reference/transcribe.py:117 ``whisperx.load_audio`` = ``ffmpeg -ac 1 -ar 16000 -f s16le`` then
int16/32768).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

ffmpeg (libswresample) is absent from the container and its filter bank is not reproducible from memory, so parity with the
reference's decoder is UNPINNED.  The algorithm restated here is the published polyphase windowed-sinc of
``torchaudio.functional.resample`` (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99), which IS importable:
PARITY PINNED BY torchaudio (kernel taps bit-equal; output within 2e-5 of torchaudio's fp32 conv1d and within 1e-6 of the
float64 evaluation of the same taps; tests/test_oracle_resample.py).  The s16 quantisation
step mirrors the s16le pipe: round-half-even of x*32768, clipped to int16, then /32768.
"""
from __future__ import annotations

import math

import numpy as np


def sinc_kernel(orig: int, new: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """-> (kernels f32 [new, 2*width + orig], width) for the reduced ratio orig:new (torchaudio _get_sinc_resample_kernel)."""
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]   # fp32 division upstream
    t = (phase + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / orig)
    return k.astype(np.float32), width, orig, new


def resample(wave: np.ndarray, orig: int, new: int) -> np.ndarray:
    """float32 [n] at `orig` Hz -> float32 [ceil(n * new / orig)] at `new` Hz."""
    k, width, o, n = sinc_kernel(orig, new)
    if o == n:
        return wave.astype(np.float32).copy()
    length = len(wave)
    padded = np.concatenate([np.zeros(width, np.float32), wave.astype(np.float32), np.zeros(width + o, np.float32)])
    n_blocks = (len(padded) - k.shape[1]) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(padded, k.shape[1])[::o][:n_blocks]       # [blocks, taps]
    out = (win.astype(np.float32) @ k.T.astype(np.float32)).reshape(-1)                        # block-major, phase-minor
    return out[: int(math.ceil(n * length / o))].astype(np.float32)


def decode_pcm16(pcm: np.ndarray, channels: int, orig: int, new: int = 16000, quantize: bool = True) -> np.ndarray:
    """Interleaved int16 [frames * channels] -> mono float32 at `new` Hz the way the s16le pipe delivers it."""
    x = pcm.reshape(-1, channels).astype(np.float32).mean(axis=1) / np.float32(32768.0)
    y = resample(x, orig, new)
    if quantize:
        y = np.clip(np.rint(y * np.float32(32768.0)), -32768, 32767).astype(np.float32) / np.float32(32768.0)
    return y
