"""Host mirror of ``whisperx.audio`` for the hot path: constants, ``load_audio``, ``log_mel_spectrogram``.

The reference reaches these through ``whisperx.load_audio`` (/root/reference/transcribe.py:117) and,
inside ``model.transcribe`` (/root/reference/transcribe.py:123), ``whisperx.audio.log_mel_spectrogram``.
The arithmetic runs in the hand-written CUDA kernels of csrc/logmel.cu through the C ABI
(include/mw_b200.h: mw_logmel / mw_logmel_long); there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import shutil
import subprocess
import wave
from functools import lru_cache
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .config import SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH, N_SAMPLES, N_FRAMES, N_FREQ

N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """File -> mono float32 at 16 kHz in [-1, 1) (``ffmpeg ... -f s16le -ac 1 -ar 16000``,
    then int16/32768).  ffmpeg is used when present; otherwise 16-bit PCM WAV files (the format the reference's web
    recorder produces, /root/reference/web/audioRecorder.js:100-127) are read directly, and any other rate or channel
    count goes through the GPU decoder (mw_pcm_resample: channel mean + band-limited resampling + s16 quantisation)."""
    ffmpeg = shutil.which("ffmpeg")
    if ffmpeg:
        cmd = [ffmpeg, "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1",
               "-acodec", "pcm_s16le", "-ar", str(sr), "-"]
        try:
            out = subprocess.run(cmd, capture_output=True, check=True).stdout
        except subprocess.CalledProcessError as e:
            raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
        return np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0
    pcm, channels, rate = _read_wav_pcm16(file)
    if rate == sr:
        if channels > 1:
            pcm = pcm.reshape(-1, channels).astype(np.int32).sum(axis=1) // channels
        return pcm.astype(np.float32) / 32768.0
    if not torch.cuda.is_available():
        raise RuntimeError(
            f"Failed to load audio: without ffmpeg, {channels}-channel {rate} Hz WAV needs the GPU decoder "
            f"(mw_pcm_resample) and no CUDA device is available; only mono 16-bit PCM WAV at {sr} Hz is readable on the host")
    return decode_pcm_device(pcm, channels, rate, sr).cpu().numpy()


def _read_wav_pcm16(file: str):
    try:
        with wave.open(file, "rb") as w:
            if w.getsampwidth() != 2:
                raise RuntimeError(f"Failed to load audio: without ffmpeg only 16-bit PCM WAV is readable (got {8 * w.getsampwidth()}-bit)")
            pcm = np.frombuffer(bytearray(w.readframes(w.getnframes())), np.int16)      # writable: torch.from_numpy needs it
            return pcm, w.getnchannels(), w.getframerate()
    except (wave.Error, EOFError, OSError) as e:
        raise RuntimeError(f"Failed to load audio: {e} (ffmpeg is not installed)") from e


@lru_cache(maxsize=16)
def sinc_resample_kernel(orig: int, new: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Polyphase windowed-sinc taps of ``torchaudio.functional.resample`` (sinc_interp_hann) for the rate pair, as numpy:
    (kernels f32 [new', 2*width + orig'], lo_hi i32 [new', 2] non-zero tap range per phase, width, orig', new') with
    orig':new' the reduced ratio.  Evaluated in float64 and cast, as upstream does (its phase offsets are fp32)."""
    import math
    g = math.gcd(int(orig), int(new))
    orig, new = int(orig) // g, int(new) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = np.clip((phase + idx) * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = (np.where(t == 0, 1.0, np.sin(t) / t) * window * (base / orig)).astype(np.float32)
    lo_hi = np.zeros((new, 2), dtype=np.int32)
    for i in range(new):
        nz = np.nonzero(k[i])[0]
        lo_hi[i] = (nz[0], nz[-1] + 1) if len(nz) else (0, 0)
    return k, lo_hi, width, orig, new


def decode_pcm_device(pcm, channels: int, rate: int, sr: int = SAMPLE_RATE, device=None, quantize_s16: bool = True) -> torch.Tensor:
    """Interleaved PCM (int16 or float32; numpy or a CUDA tensor) at `rate` Hz -> mono float32 at `sr` Hz ON THE DEVICE:
    channel mean, band-limited resampling and the s16le pipe's quantisation in one kernel (include/mw_b200.h:
    mw_pcm_resample).  The result can be passed straight to ``model.transcribe``."""
    import math
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("decode_pcm_device needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    device = torch.device(device if device is not None else "cuda")
    x = pcm if torch.is_tensor(pcm) else torch.from_numpy(np.ascontiguousarray(pcm))
    if x.dtype not in (torch.int16, torch.float32):
        raise ValueError(f"PCM must be int16 or float32, got {x.dtype}")
    x = x.reshape(-1).to(device).contiguous()
    if channels < 1 or x.numel() % channels:
        raise ValueError(f"{x.numel()} samples do not divide into {channels} channels")
    n_frames = x.numel() // channels
    k, lo_hi, width, o, n = sinc_resample_kernel(int(rate), int(sr))
    if o == n:      # same rate: no filtering (as torchaudio / ffmpeg), the kernel only mixes channels and quantises
        k, lo_hi, width = np.ones((1, 1), np.float32), np.array([[0, 1]], np.int32), 0
    n_out = int(math.ceil(n * n_frames / o))
    out = torch.empty(n_out, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        d_k = torch.from_numpy(k).to(device)
        d_lh = torch.from_numpy(lo_hi).to(device)
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.mw_pcm_resample(x.data_ptr(), n_frames, int(channels), 0 if x.dtype == torch.int16 else 1, o, n,
                                       d_k.data_ptr(), d_lh.data_ptr(), k.shape[1], width, out.data_ptr(), n_out,
                                       int(bool(quantize_s16)), st), "mw_pcm_resample")
    return out


def load_audio_device(file: str, sr: int = SAMPLE_RATE, device=None) -> torch.Tensor:
    """16-bit PCM WAV (any rate, any channel count) -> mono float32 at `sr` Hz resident on the GPU, without ffmpeg and
    without a host round trip of the decoded waveform."""
    pcm, channels, rate = _read_wav_pcm16(file)
    return decode_pcm_device(pcm, channels, rate, sr, device)


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """whisperx.audio.pad_or_trim: zero-pad or cut `axis` to `length`."""
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = torch.nn.functional.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


@lru_cache(maxsize=None)
def mel_filters_np(n_mels: int) -> np.ndarray:
    """The rows of whisperx ``assets/mel_filters.npz`` (librosa slaney filterbank, sr 16 kHz, n_fft 400),
    regenerated because the asset file is not on disk.  float32 [n_mels, 201]."""
    if n_mels <= 0:
        raise ValueError("n_mels must be positive")
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    top = SAMPLE_RATE / 2.0
    top_mel = min_log_mel + np.log(top / min_log_hz) / logstep
    m = np.linspace(0.0, top_mel, n_mels + 2)
    hz = np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)
    bins = np.linspace(0.0, top, N_FREQ)
    up = (bins[None, :] - hz[:-2, None]) / (hz[1:-1] - hz[:-2])[:, None]
    down = (hz[2:, None] - bins[None, :]) / (hz[2:] - hz[1:-1])[:, None]
    tri = np.clip(np.minimum(up, down), 0.0, None)
    tri *= (2.0 / (hz[2:] - hz[:-2]))[:, None]
    return np.ascontiguousarray(tri.astype(np.float32))


def mel_filters(device, n_mels: int) -> torch.Tensor:
    """whisperx.audio.mel_filters(device, n_mels)."""
    return torch.from_numpy(mel_filters_np(n_mels)).to(device)


class LogMelPlan:
    """Owns an ``mw_logmel_plan`` (sparse filterbank + twiddle tables resident on one GPU)."""

    def __init__(self, n_mels: int, device: Union[int, torch.device, str] = 0, max_chunks: int = 4096,
                 filters: Optional[np.ndarray] = None):
        self.lib = _lib.load()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError("log-mel runs on a CUDA device only (no CPU fallback)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.n_mels = int(n_mels)
        self.max_chunks = int(max_chunks)
        f = mel_filters_np(n_mels) if filters is None else np.ascontiguousarray(filters, dtype=np.float32)
        if f.shape != (n_mels, N_FREQ):
            raise ValueError(f"filters must be [{n_mels}, {N_FREQ}]")
        handle = C.c_void_p()
        _lib.check(self.lib.mw_logmel_plan_create(self.n_mels, f.ctypes.data_as(_lib.c_f32p), self.max_chunks,
                                                  self.device.index, C.byref(handle)), "mw_logmel_plan_create")
        self.handle = handle

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self.lib.mw_logmel_plan_destroy(h)
            self.handle = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def chunks(self, audio: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor,
               out: Optional[torch.Tensor] = None, out_t: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio f32 [N] on device; offsets int64 [n], lengths int32 [n] on device ->
        f32 [n, n_mels, 3000] (each chunk zero-padded to 30 s, own global max)."""
        n = int(offsets.shape[0])
        if out is None:
            out = torch.empty((n, self.n_mels, N_FRAMES), dtype=torch.float32, device=self.device)
        assert audio.dtype == torch.float32 and audio.is_contiguous() and audio.device == self.device
        assert offsets.dtype == torch.int64 and lengths.dtype == torch.int32
        assert out.is_contiguous() and out.shape == (n, self.n_mels, N_FRAMES)
        if out_t is not None:
            assert out_t.dtype == _lib.storage_dtype() and out_t.is_contiguous()
            assert out_t.shape == (n, N_FRAMES + 2, self.n_mels)
        _lib.check(self.lib.mw_logmel(self.handle, audio.data_ptr(), audio.numel(), offsets.data_ptr(),
                                      lengths.data_ptr(), n, out.data_ptr(),
                                      out_t.data_ptr() if out_t is not None else None, self._stream()), "mw_logmel")
        return out

    def long(self, audio: torch.Tensor, padding: int = 0) -> torch.Tensor:
        """Un-chunked log_mel_spectrogram(audio, n_mels, padding): f32 [n_mels, (N+padding)//160]."""
        assert audio.dtype == torch.float32 and audio.is_contiguous() and audio.device == self.device
        n = audio.numel()
        out = torch.empty((self.n_mels, (n + padding) // HOP_LENGTH), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.mw_logmel_long(self.handle, audio.data_ptr() if n else None, n, int(padding),
                                           out.data_ptr(), self._stream()), "mw_logmel_long")
        return out


_PLANS = {}


def get_plan(n_mels: int, device) -> LogMelPlan:
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    if dev.type != "cuda":
        raise RuntimeError("manual_whisper_b200.log_mel_spectrogram needs a CUDA device; there is no CPU fallback")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (int(n_mels), idx)
    if key not in _PLANS:
        _PLANS[key] = LogMelPlan(n_mels, idx)
    return _PLANS[key]


def log_mel_spectrogram(audio: Union[str, np.ndarray, torch.Tensor], n_mels: int, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """Same signature as ``whisperx.audio.log_mel_spectrogram``; computed by the fused
    sm_100a kernels.  Returns a float32 CUDA tensor [n_mels, (T+padding)//160]."""
    if isinstance(audio, str):
        audio = load_audio(audio)
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    if device is None:
        device = audio.device if audio.is_cuda else torch.device("cuda", torch.cuda.current_device())
    audio = audio.to(device=device, dtype=torch.float32).contiguous()
    if audio.dim() != 1:
        raise ValueError("audio must be 1-D (mono)")
    return get_plan(n_mels, audio.device).long(audio, padding)
