"""Whisper model dimensions and special-token ids for the hot path.

The reference never spells these out: it passes a size name to
``whisperx.load_model(MODEL_SIZE, ...)`` (/root/reference/transcribe.py:33,107-113)
and the upstream packages resolve it.  SURVEY.md Appendix C is the table
restated here.  Token ids are *parameters* of every kernel and of the oracle,
so a wrong id can only hurt realism, never parity.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import List, Optional

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000
N_AUDIO_CTX = 1500
N_TEXT_CTX = 448
N_FREQ = N_FFT // 2 + 1  # 201


@dataclass(frozen=True)
class ModelDims:
    name: str
    n_mels: int
    d_model: int
    n_heads: int
    enc_layers: int
    dec_layers: int
    ffn: int
    vocab: int
    n_audio_ctx: int = N_AUDIO_CTX
    n_text_ctx: int = N_TEXT_CTX

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    def asdict(self):
        return asdict(self)


_DIMS = {
    # name: (n_mels, d, heads, enc L, dec L, ffn, vocab)
    "tiny": (80, 384, 6, 4, 4, 1536, 51865),
    "base": (80, 512, 8, 6, 6, 2048, 51865),
    "small": (80, 768, 12, 12, 12, 3072, 51865),
    "medium": (80, 1024, 16, 24, 24, 4096, 51865),
    "large-v2": (80, 1280, 20, 32, 32, 5120, 51865),
    "large-v3": (128, 1280, 20, 32, 32, 5120, 51866),
}
_ALIASES = {"large": "large-v3", "large-v1": "large-v2"}


def model_dims(name: str) -> ModelDims:
    key = _ALIASES.get(name, name)
    if key.endswith(".en"):
        raise ValueError(f"'{name}': English-only checkpoints (vocab 51864, their own control-token ids) are not supported")
    if key not in _DIMS:
        raise ValueError(f"Invalid model size '{name}', expected one of: {', '.join(sorted(_DIMS))}")
    return ModelDims(key, *_DIMS[key])


def custom_dims(name, n_mels, d_model, n_heads, enc_layers, dec_layers, ffn, vocab,
                n_audio_ctx=N_AUDIO_CTX, n_text_ctx=N_TEXT_CTX) -> ModelDims:
    """Reduced-size architectures for fast parity tests (same code path, d_head must be 64)."""
    if d_model // n_heads != 64 or d_model % n_heads:
        raise ValueError("d_head must be 64 (every Whisper size has d_head=64)")
    return ModelDims(name, n_mels, d_model, n_heads, enc_layers, dec_layers, ffn, vocab, n_audio_ctx, n_text_ctx)


# Non-speech token list of the multilingual vocabulary (the ids CTranslate2 reads from the
# converted model's config.json "suppress_ids"; the same list is restated in-container at
# transformers/models/whisper/configuration_whisper.py:34-45).  large-v3 shifts ids >= 50259+99 by one.
_NON_SPEECH_MULTI = [
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647,
    7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
    18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863, 47425, 49870,
    50254,
]

LANGUAGES = [
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi",
    "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml",
    "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs",
    "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am",
    "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln",
    "ha", "ba", "jw", "su", "yue",
]


@dataclass(frozen=True)
class SpecialTokens:
    """Ids of the control tokens."""
    vocab: int
    eot: int
    sot: int
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int
    blank: int = 220
    n_langs: int = 99
    suppress_ids: List[int] = field(default_factory=list)

    def lang_id(self, language: str) -> int:
        if language not in LANGUAGES[: self.n_langs]:
            raise ValueError(f"'{language}' is not a valid language code")
        return self.sot + 1 + LANGUAGES.index(language)

    @property
    def suppress_ids_begin(self) -> List[int]:
        return [self.blank, self.eot]


def special_tokens(vocab: int) -> SpecialTokens:
    n_langs = 100 if vocab >= 51866 else 99
    sot = 50258
    translate = sot + 1 + n_langs
    st = dict(vocab=vocab, eot=50257, sot=sot, translate=translate, transcribe=translate + 1,
              sot_lm=translate + 2, sot_prev=translate + 3, no_speech=translate + 4,
              no_timestamps=translate + 5, timestamp_begin=translate + 6, n_langs=n_langs)
    # "suppress_tokens=[-1]" upstream expands to the non-speech list plus the control tokens that
    # must never be sampled: sot, translate, transcribe, sot_lm, sot_prev, no_speech.
    sup = sorted(set(_NON_SPEECH_MULTI + [sot, translate, translate + 1, translate + 2, translate + 3, translate + 4]))
    return SpecialTokens(suppress_ids=[t for t in sup if t < vocab], **st)


def scaled_tokens(vocab: int) -> SpecialTokens:
    """Control-token ids squeezed into a small test vocabulary (vocab >= 1024), keeping the same ordering
    eot < sot < langs < translate < ... < timestamp_begin so every logit rule is exercised."""
    if vocab >= 51865:
        return special_tokens(vocab)
    if vocab < 1024:
        raise ValueError("test vocab must be >= 1024")
    n_ts = min(1501, vocab // 2)
    ts_begin = vocab - n_ts
    n_langs = 4
    sot = ts_begin - 7 - n_langs
    translate = sot + 1 + n_langs
    sup = sorted({1, 2, 7, 8, 9, 10, 14, 25, sot, translate, translate + 1, translate + 2, translate + 3, translate + 4})
    return SpecialTokens(vocab=vocab, eot=sot - 1, sot=sot, translate=translate, transcribe=translate + 1,
                         sot_lm=translate + 2, sot_prev=translate + 3, no_speech=translate + 4,
                         no_timestamps=translate + 5, timestamp_begin=ts_begin, blank=220, n_langs=n_langs,
                         suppress_ids=sup)
