"""Token ids for prompts, and text decoding when a vocabulary is available.

Upstream: ``faster_whisper.tokenizer.Tokenizer`` over Hugging Face ``tokenizers`` (SURVEY.md §2.2
"Tokenizer"; out of scope for arithmetic, but the special-token ids are inputs of the hot path).  No
tokenizer.json exists offline, so text is unpinned: without a vocabulary file ``decode`` renders the ids
themselves and ``encode`` falls back to one id per UTF-8 byte (documented placeholder, used identically
by the oracle harness).  With ``tokenizer_file=<path to tokenizer.json>`` the real BPE is used.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

from .config import SpecialTokens, LANGUAGES


class Tokenizer:
    def __init__(self, tokens: SpecialTokens, multilingual: bool = True, task: Optional[str] = "transcribe",
                 language: Optional[str] = "en", tokenizer_file: Optional[str] = None, hf=None):
        self.tokens = tokens
        self.multilingual = multilingual
        self.task_name = task
        self.language_code = language
        self.hf = hf                      # an already loaded tokenizers.Tokenizer (the pipeline keeps one and re-uses it)
        if self.hf is None and tokenizer_file:
            import tokenizers
            self.hf = tokenizers.Tokenizer.from_file(tokenizer_file)
        if multilingual:
            if task not in ("transcribe", "translate"):
                raise ValueError(f"'{task}' is not a valid task (accepted tasks: transcribe, translate)")
            if language not in LANGUAGES[: tokens.n_langs]:
                raise ValueError(f"'{language}' is not a valid language code")
            self.task = tokens.transcribe if task == "transcribe" else tokens.translate
            self.language = tokens.lang_id(language)
        else:
            self.task = None
            self.language = None

    sot = property(lambda self: self.tokens.sot)
    sot_prev = property(lambda self: self.tokens.sot_prev)
    sot_lm = property(lambda self: self.tokens.sot_lm)
    eot = property(lambda self: self.tokens.eot)
    no_timestamps = property(lambda self: self.tokens.no_timestamps)
    no_speech = property(lambda self: self.tokens.no_speech)
    timestamp_begin = property(lambda self: self.tokens.timestamp_begin)
    transcribe = property(lambda self: self.tokens.transcribe)
    translate = property(lambda self: self.tokens.translate)

    @property
    def sot_sequence(self) -> List[int]:
        seq = [self.sot]
        if self.language is not None:
            seq.append(self.language)
        if self.task is not None:
            seq.append(self.task)
        return seq

    def encode(self, text: str) -> List[int]:
        if self.hf is not None:
            return self.hf.encode(text, add_special_tokens=False).ids
        if text.strip():
            import warnings
            warnings.warn("no tokenizer.json is loaded: text (initial_prompt / hotwords / prefix) is encoded one id per UTF-8 "
                          "byte, which is NOT the Whisper BPE; pass tokenizer_file= to load_model for real prompts", stacklevel=2)
        return [b for b in text.encode("utf-8")]     # placeholder: one id per byte (ids 0..255)

    def decode(self, ids: Sequence[int]) -> str:
        ids = [int(t) for t in ids if t < self.eot]
        if self.hf is not None:
            return self.hf.decode(ids)
        return " ".join(str(t) for t in ids)

    def decode_batch(self, batch: Sequence[Sequence[int]]) -> List[str]:
        return [self.decode(ids) for ids in batch]
