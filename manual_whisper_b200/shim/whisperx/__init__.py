"""``import whisperx`` shim: put ``manual_whisper_b200/shim`` on PYTHONPATH and the reference's transcribe.py
runs unchanged on the B200 engine (after setting DEVICE="cuda" as /root/reference/README.md:101 instructs).

On the hot path (B200-native): load_model, load_audio, log_mel_spectrogram.
Off the hot path (out of scope): alignment and diarization.  If a real whisperx is installed
under another name they should be taken from there; here ``align`` returns the segments unaligned and the
diarization entry points raise, which the reference tolerates (transcribe.py:141-149 wraps them in try/except).
"""
import warnings

from manual_whisper_b200 import (load_model, load_audio, log_mel_spectrogram, merge_chunks,  # noqa: F401
                                 SAMPLE_RATE, N_FFT, HOP_LENGTH, CHUNK_LENGTH, N_SAMPLES, N_FRAMES)
from manual_whisper_b200 import asr, audio, vad  # noqa: F401


def load_align_model(language_code, device, model_name=None, model_dir=None):
    warnings.warn("whisperx shim: forced alignment (wav2vec2) is outside the B200 hot path; segments stay unaligned")
    return None, {"language": language_code, "dictionary": {}, "type": "none"}


def align(transcript, model, align_model_metadata, audio, device, interpolate_method="nearest",
          return_char_alignments=False, print_progress=False, combined_progress=False):
    segments = [dict(s) for s in transcript]
    return {"segments": segments, "word_segments": []}


class DiarizationPipeline:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("whisperx shim: speaker diarization (pyannote) is outside the B200 hot path")


def assign_word_speakers(diarize_df, transcript_result, *args, **kwargs):
    return transcript_result
