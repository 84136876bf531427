"""manual_whisper_b200 — B200-native (sm_100a) implementation of the hot path behind
LuSicong22/manual-whisper's ``transcribe.py``: ``whisperx.load_model(...).transcribe(audio, batch_size)``.

Public surface = the whisperx names the reference uses (/root/reference/transcribe.py:7,107-113,117,123):
``load_model``, ``load_audio``, plus ``log_mel_spectrogram`` and the audio constants.  All arithmetic runs in
hand-written CUDA behind the C ABI of include/mw_b200.h (libmw_b200.so); importing this package never
imports the CPU oracle, and every call fails loudly if the CUDA library is missing.
"""
from .config import SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH, N_SAMPLES, N_FRAMES, model_dims, special_tokens
from .audio import load_audio, load_audio_device, decode_pcm_device, log_mel_spectrogram, pad_or_trim, mel_filters
from .asr import load_model, FasterWhisperPipeline, WhisperModel, TranscriptionOptions, get_prompt
from .vad import merge_chunks, InjectedVad, EnergyVad, GpuEnergyVad, synthetic_speech
from .tokenizer import Tokenizer
from .alignment import load_align_model, align, AlignModel, AlignEngine
from .w2v import W2vDims, random_init_w2v

__version__ = "0.1.0"
__all__ = [
    "load_audio_device", "decode_pcm_device",
    "load_align_model", "align", "AlignModel", "AlignEngine", "W2vDims", "random_init_w2v",
    "SAMPLE_RATE", "N_FFT", "HOP_LENGTH", "CHUNK_LENGTH", "N_SAMPLES", "N_FRAMES",
    "load_audio", "log_mel_spectrogram", "pad_or_trim", "mel_filters", "load_model", "FasterWhisperPipeline",
    "WhisperModel", "TranscriptionOptions", "get_prompt", "merge_chunks", "InjectedVad", "EnergyVad", "GpuEnergyVad",
    "synthetic_speech", "Tokenizer", "model_dims", "special_tokens",
]
