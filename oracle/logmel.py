"""ORACLE (test infrastructure, not product code) — CPU restatement of the log-mel front end.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Restates ``whisperx.audio.log_mel_spectrogram(audio, n_mels, padding, device)`` (whisperx 3.7.6, the
only pin: /root/reference/transcribe_colab.ipynb:47,80), reached from the reference through
``model.transcribe`` (/root/reference/transcribe.py:123).  whisperx is NOT installed here, so the
formula is SURVEY.md Appendix A.3 and is pinned against the in-container Hugging Face twin
(transformers/models/whisper/feature_extraction_whisper.py:135-163) by tests/test_oracle_logmel.py
and the committed vectors in tests/golden/.  PARITY PINNED BY: HF twin + golden vectors; the
reference itself holds no vectors ("parity unpinned" upstream).
"""
from __future__ import annotations

import numpy as np
import torch

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000


def _hz_to_mel(f):
    """Slaney mel scale (librosa.hz_to_mel, htk=False): linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters(n_mels: int, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> np.ndarray:
    """``librosa.filters.mel(sr=16000, n_fft=400, n_mels=n_mels)`` (slaney scale + slaney norm, fmax=sr/2),
    the content of whisperx ``assets/mel_filters.npz``.  Returns float32 [n_mels, 201]."""
    n_freq = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freq)
    mel_pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_freq), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_pts[2: n_mels + 2] - mel_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def log_mel_spectrogram(audio, n_mels: int, padding: int = 0, filters=None) -> torch.Tensor:
    """SURVEY.md A.3, line for line.  audio float32 [T] -> float32 [n_mels, (T+padding)//160]."""
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    audio = audio.to(torch.float32)
    if padding > 0:
        audio = torch.nn.functional.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    if filters is None:
        filters = mel_filters(n_mels)
    filters = torch.as_tensor(filters, dtype=torch.float32)
    mel_spec = filters @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec


def log_mel_chunks(audio: np.ndarray, offsets, lengths, n_mels: int) -> torch.Tensor:
    """The pipeline's per-chunk call (SURVEY.md A.5 ``preprocess``): every chunk padded to 30 s, own global max."""
    out = []
    for o, n in zip(offsets, lengths):
        out.append(log_mel_spectrogram(audio[o: o + n], n_mels, padding=N_SAMPLES - n))
    return torch.stack(out) if out else torch.zeros(0, n_mels, N_SAMPLES // HOP_LENGTH)


def log_mel_float64(audio: np.ndarray, n_mels: int, padding: int = 0) -> np.ndarray:
    """Independent float64 direct-DFT restatement (no torch.stft) used to bound the fp32 oracle's own error."""
    x = np.asarray(audio, dtype=np.float64)
    if padding > 0:
        x = np.concatenate([x, np.zeros(padding)])
    T = x.shape[0]
    n_frames = T // HOP_LENGTH
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n = np.arange(N_FFT)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)
    k = np.arange(N_FFT // 2 + 1)
    basis = np.exp(-2j * np.pi * np.outer(k, n) / N_FFT)
    idx = np.arange(n_frames)[:, None] * HOP_LENGTH + n[None, :]
    frames = xp[idx] * win[None, :]
    power = np.abs(frames @ basis.T) ** 2  # [frames, 201]
    mel = mel_filters(n_mels).astype(np.float64) @ power.T
    log = np.log10(np.maximum(mel, 1e-10))
    log = np.maximum(log, log.max() - 8.0)
    return (log + 4.0) / 4.0
