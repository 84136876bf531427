"""``import whisperx`` shim: put ``manual_whisper_b200/shim`` on PYTHONPATH and the reference's transcribe.py
runs unchanged on the B200 engine (after setting DEVICE="cuda" as /root/reference/README.md:101 instructs).

On the B200 engine: load_model, load_audio, log_mel_spectrogram, and load_align_model / align.
Off the path (out of scope): diarization - the entry points raise, which the reference tolerates
(transcribe.py:141-149 wraps them in try/except).
"""
from manual_whisper_b200 import (load_model, load_audio, log_mel_spectrogram, merge_chunks,  # noqa: F401
                                 load_align_model, align,
                                 SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH, N_SAMPLES, N_FRAMES)
from manual_whisper_b200 import asr, audio, vad, alignment  # noqa: F401


class DiarizationPipeline:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("whisperx shim: speaker diarization (pyannote) is outside the B200 hot path")


def assign_word_speakers(diarize_df, transcript_result, *args, **kwargs):
    return transcript_result
