"""Host mirror of ``whisperx.asr``: ``load_model`` and ``FasterWhisperPipeline.transcribe``.

This is the drop-in boundary for the two reference lines
    model = whisperx.load_model(MODEL_SIZE, DEVICE, compute_type=..., language="zh",
                                asr_options={"initial_prompt": ...}, vad_options=...)   # transcribe.py:107-113
    result = model.transcribe(audio, batch_size=BATCH_SIZE, language="zh")              # transcribe.py:123
Same names, keyword arguments, return structure and error behaviour as whisperx 3.7.6 (SURVEY.md §8 a1/a2,
A.5, A.6); everything numeric runs in libmw_b200.so on the GPU(s) (no CTranslate2, no CPU fallback).

Data path per call: audio -> (VAD turns -> merge_chunks) -> one H2D copy of the waveform ->
per batch of B windows: fused log-mel kernels -> encoder -> batched decode -> ids (D2H) -> text.
With several GPUs (device_index=[...]) batches are dealt round-robin to one replica per GPU, each driven by
its own host thread; results are gathered on the host and restored to chunk order (no collective).
"""
from __future__ import annotations

import threading
import warnings
from dataclasses import dataclass, field, replace
from typing import Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from . import audio as _audio
from .audio import SAMPLE_RATE, N_SAMPLES, N_FRAMES, load_audio
from .config import ModelDims, SpecialTokens, model_dims, special_tokens, LANGUAGES
from .engine import Engine, GenerationResult
from .tokenizer import Tokenizer
from .vad import EnergyVad, GpuEnergyVad, merge_chunks
from .weights import random_init


@dataclass
class TranscriptionOptions:
    """faster_whisper.transcribe.TranscriptionOptions as whisperx fills it (SURVEY.md A.6 default list)."""
    beam_size: int = 5
    best_of: int = 5
    patience: float = 1
    length_penalty: float = 1
    repetition_penalty: float = 1
    no_repeat_ngram_size: int = 0
    temperatures: Sequence[float] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)
    compression_ratio_threshold: Optional[float] = 2.4
    log_prob_threshold: Optional[float] = -1.0
    no_speech_threshold: Optional[float] = 0.6
    condition_on_previous_text: bool = False
    prompt_reset_on_temperature: float = 0.5
    initial_prompt: Optional[str] = None
    prefix: Optional[str] = None
    suppress_blank: bool = True
    suppress_tokens: Optional[List[int]] = field(default_factory=lambda: [-1])
    without_timestamps: bool = True
    max_initial_timestamp: float = 0.0
    word_timestamps: bool = False
    prepend_punctuations: str = "\"'“¿([{-"
    append_punctuations: str = "\"'.。,，!！?？:：”)]}、"
    multilingual: bool = False
    suppress_numerals: bool = False
    max_new_tokens: Optional[int] = None
    clip_timestamps: Optional[str] = None
    hallucination_silence_threshold: Optional[float] = None
    hotwords: Optional[str] = None


DEFAULT_VAD_OPTIONS = {"chunk_size": 30, "vad_onset": 0.500, "vad_offset": 0.363}
_COMPUTE_TYPES = {"default", "auto", "int8", "int8_float16", "int8_bfloat16", "int8_float32", "int16", "float16",
                  "bfloat16", "float32"}


def get_prompt(tokenizer: Tokenizer, previous_tokens: List[int], without_timestamps: bool = False,
               prefix: Optional[str] = None, hotwords: Optional[str] = None, max_length: int = 448) -> List[int]:
    """faster_whisper.WhisperModel.get_prompt."""
    prompt: List[int] = []
    if previous_tokens or (hotwords and not prefix):
        prompt.append(tokenizer.sot_prev)
        if hotwords and not prefix:
            hot = tokenizer.encode(" " + hotwords.strip())
            if len(hot) >= max_length // 2:
                hot = hot[: max_length // 2 - 1]
            prompt.extend(hot)
        if previous_tokens:
            prompt.extend(previous_tokens[-(max_length // 2 - 1):])
    prompt.extend(tokenizer.sot_sequence)
    if without_timestamps:
        prompt.append(tokenizer.no_timestamps)
    if prefix:
        ptoks = tokenizer.encode(" " + prefix.strip())
        if len(ptoks) >= max_length // 2:
            ptoks = ptoks[: max_length // 2 - 1]
        if not without_timestamps:
            prompt.append(tokenizer.timestamp_begin)
        prompt.extend(ptoks)
    return prompt


class WhisperModel:
    """The role of whisperx.asr.WhisperModel(faster_whisper.WhisperModel): one engine replica + batched generate."""

    def __init__(self, dims: ModelDims, state_dict, device_index: int = 0, max_batch: int = 32, max_beam: int = 5,
                 tokens: Optional[SpecialTokens] = None, share_weights_with: Optional["WhisperModel"] = None):
        self.dims = dims
        self.tokens = tokens or special_tokens(dims.vocab)
        self.device = torch.device("cuda", device_index)
        packed = share_weights_with.engine.weights if share_weights_with is not None else None
        self.engine = Engine(dims, state_dict, self.device, max_batch=max_batch, max_beam=max_beam, packed=packed)
        # every replica drives its own stream so that several batches can be in flight on one GPU
        self.stream = torch.cuda.Stream(device=self.device) if share_weights_with is not None else None
        self.max_length = dims.n_text_ctx
        self.feat_kwargs = {"feature_size": dims.n_mels}
        self.is_multilingual = True
        self.plan = _audio.LogMelPlan(dims.n_mels, self.device, max_chunks=max_batch)
        self._feat_t = torch.empty((max_batch, N_FRAMES + 2, dims.n_mels), dtype=self.engine.h16, device=self.device)
        self._feat = torch.empty((max_batch, dims.n_mels, N_FRAMES), dtype=torch.float32, device=self.device)

    @property
    def model(self):
        return self.engine

    def encode(self, features: torch.Tensor) -> torch.Tensor:
        if features.dim() == 2:
            features = features.unsqueeze(0)
        return self.engine.encode(features.to(self.device, torch.float32).contiguous())

    def generate_segment_batched(self, features: torch.Tensor, tokenizer: Tokenizer, options: TranscriptionOptions,
                                 encoder_output=None, forced_eot_len: int = 0):
        """whisperx generate_segment_batched: one shared prompt, single deterministic pass.
        Returns (texts, token id lists)."""
        all_tokens: List[int] = []
        if options.initial_prompt is not None:
            all_tokens.extend(tokenizer.encode(" " + options.initial_prompt.strip()))
        prompt = get_prompt(tokenizer, all_tokens, without_timestamps=options.without_timestamps, prefix=options.prefix,
                            hotwords=options.hotwords, max_length=self.max_length)
        enc = encoder_output if encoder_output is not None else self.encode(features)
        results = self.engine.generate(enc, prompt, self.tokens, beam_size=options.beam_size, patience=options.patience,
                                       length_penalty=options.length_penalty, max_length=self.max_length,
                                       suppress_blank=options.suppress_blank, suppress_tokens=options.suppress_tokens,
                                       without_timestamps=options.without_timestamps, forced_eot_len=forced_eot_len)
        tokens_batch = [r.sequences_ids[0] for r in results]
        text = tokenizer.decode_batch([[t for t in tk if t < tokenizer.eot] for tk in tokens_batch])
        return text, tokens_batch

    # one batch of windows, device-resident audio -> ids
    def transcribe_windows(self, d_audio: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor,
                           tokenizer: Tokenizer, options: TranscriptionOptions, forced_eot_len: int = 0):
        n = offsets.shape[0]
        feat_t = self._feat_t[:n]
        self.plan.chunks(d_audio, offsets, lengths, out=self._feat[:n], out_t=feat_t)
        enc = self.engine.encode_time_major(feat_t)
        return self.generate_segment_batched(None, tokenizer, options, encoder_output=enc, forced_eot_len=forced_eot_len)


class FasterWhisperPipeline:
    """whisperx.asr.FasterWhisperPipeline: VAD -> windows -> batched ASR."""

    def __init__(self, model: Union[WhisperModel, List[WhisperModel]], vad, vad_params: dict, options: TranscriptionOptions,
                 tokenizer: Optional[Tokenizer] = None, device: Union[int, str, torch.device] = -1,
                 framework: str = "pt", language: Optional[str] = None, suppress_numerals: bool = False,
                 hf_tokenizer=None, **kwargs):
        self.hf_tokenizer = hf_tokenizer if hf_tokenizer is not None else getattr(tokenizer, "hf", None)
        self.replicas: List[WhisperModel] = list(model) if isinstance(model, (list, tuple)) else [model]
        self.model = self.replicas[0]
        self.tokenizer = tokenizer
        self.options = options
        self.preset_language = language
        self.suppress_numerals = suppress_numerals
        self._batch_size = kwargs.pop("batch_size", None)
        self.vad_model = vad
        self._vad_params = vad_params
        self.device = self.model.device
        self.last_stats: Dict = {}

    # -- HF-pipeline-shaped helpers kept for API parity
    def preprocess(self, audio):
        a = audio["inputs"]
        feats = _audio.log_mel_spectrogram(a, n_mels=self.model.dims.n_mels, padding=N_SAMPLES - a.shape[0], device=self.device)
        return {"inputs": feats}

    def detect_language(self, audio: np.ndarray) -> str:
        """First 30 s -> log-mel -> encoder -> softmax over the language ids at <sot>."""
        if audio.shape[0] < N_SAMPLES:
            print("Warning: audio is shorter than 30s, language detection may be inaccurate.")
        seg = torch.from_numpy(np.ascontiguousarray(audio[:N_SAMPLES], dtype=np.float32)).to(self.device)
        pad = N_SAMPLES - seg.shape[0]
        feats = _audio.get_plan(self.model.dims.n_mels, self.device).long(seg, padding=pad)
        enc = self.model.encode(feats)
        probs = self.model.engine.detect_language(enc, self.model.tokens)[0]
        k = int(np.argmax(probs))
        language = LANGUAGES[k]
        print(f"Detected language: {language} ({probs[k]:.2f}) in first 30s of audio...")
        return language

    def transcribe(self, audio: Union[str, np.ndarray], batch_size: Optional[int] = None, num_workers: int = 0,
                   language: Optional[str] = None, task: Optional[str] = None, chunk_size: int = 30,
                   print_progress: bool = False, combined_progress: bool = False, verbose: bool = False,
                   _forced_eot_len: int = 0) -> Dict:
        if isinstance(audio, str):
            audio = load_audio(audio)
        resident = None
        if torch.is_tensor(audio) and audio.is_cuda:
            # already decoded on the GPU (audio.load_audio_device / decode_pcm_device): no H2D copy at all
            d_audio = audio.to(torch.float32).reshape(-1).contiguous()
            resident = {"lo": 0, "audio": {dev: d_audio.to(dev) for dev in {rep.device for rep in self.replicas}}}
            audio = d_audio.cpu().numpy()
        elif torch.is_tensor(audio):
            audio = audio.numpy()
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        if audio.ndim != 1:
            raise ValueError("audio must be a mono 1-D float array at 16 kHz")

        # ---- VAD -> windows
        if resident is not None:
            waveform = resident["audio"][self.device].unsqueeze(0) if getattr(self.vad_model, "wants_device", False) \
                else torch.from_numpy(audio).unsqueeze(0)
        elif getattr(self.vad_model, "wants_device", False):
            # device-side VAD: one H2D copy of the whole waveform serves the VAD and every ASR batch
            resident = self.upload(audio, np.array([0], dtype=np.int64), np.array([len(audio)], dtype=np.int64))
            waveform = resident["audio"][self.device].unsqueeze(0)
        else:
            waveform = torch.from_numpy(audio).unsqueeze(0)
        if hasattr(self.vad_model, "device_windows"):
            # turns AND merge_chunks on the device (csrc/vad.cu): only the finished window table comes back
            vad_segments = self.vad_model.device_windows({"waveform": waveform, "sample_rate": SAMPLE_RATE}, chunk_size)
        else:
            vad_segments = self.vad_model({"waveform": waveform, "sample_rate": SAMPLE_RATE})
            vad_segments = merge_chunks(vad_segments, chunk_size, onset=self._vad_params["vad_onset"],
                                        offset=self._vad_params["vad_offset"])

        language, task = self._prepare_tokenizer(audio, language, task)
        segments = self.transcribe_windows_host(audio, vad_segments, batch_size=batch_size, language=language, task=task,
                                                print_progress=print_progress, combined_progress=combined_progress,
                                                verbose=verbose, _forced_eot_len=_forced_eot_len, _resident=resident)
        if self.preset_language is None:
            self.tokenizer = None
        return {"segments": segments, "language": language}

    def _prepare_tokenizer(self, audio, language, task):
        """language given => no detection; the tokenizer is rebuilt only when task/language differ."""
        if self.tokenizer is None:
            language = language or self.detect_language(audio)
            task = task or "transcribe"
            self.tokenizer = Tokenizer(self.model.tokens, self.model.is_multilingual, task=task, language=language,
                                       hf=self.hf_tokenizer)
        else:
            language = language or self.tokenizer.language_code
            task = task or self.tokenizer.task_name
            if task != self.tokenizer.task_name or language != self.tokenizer.language_code:
                self.tokenizer = Tokenizer(self.model.tokens, self.model.is_multilingual, task=task, language=language,
                                           hf=self.hf_tokenizer)
        return language, task

    def transcribe_windows_host(self, audio: np.ndarray, vad_segments: List[Dict], batch_size: Optional[int] = None,
                                language: Optional[str] = None, task: Optional[str] = None, print_progress: bool = False,
                                combined_progress: bool = False, verbose: bool = False, chunk_size: int = 30,
                                _forced_eot_len: int = 0, _resident: Optional[Dict] = None) -> List[Dict]:
        """The batched loop of ``transcribe`` over an explicit window list (what a rank of the sharded
        multi-process path runs on its share, manual_whisper_b200/distributed.py)."""
        if self.tokenizer is None or (language and language != self.tokenizer.language_code) or \
                (task and task != self.tokenizer.task_name):
            self._prepare_tokenizer(audio, language, task)
        options = self.options
        if self.suppress_numerals and self.tokenizer.hf is not None:
            previous = list(options.suppress_tokens or [])
            numeral = [i for i in range(self.tokenizer.eot)
                       if any(c in "0123456789%$£" for c in self.tokenizer.hf.decode([i]).removeprefix(" "))]
            options = replace(options, suppress_tokens=sorted(set(previous + numeral)))

        segments: List[Dict] = []
        batch_size = batch_size or self._batch_size
        if batch_size in (None, 0):
            batch_size = 1
        total = len(vad_segments)
        if total:
            offs = np.array([int(s["start"] * SAMPLE_RATE) for s in vad_segments], dtype=np.int64)
            ends = np.array([int(s["end"] * SAMPLE_RATE) for s in vad_segments], dtype=np.int64)
            offs = np.clip(offs, 0, len(audio))
            ends = np.clip(ends, offs, len(audio))
            lens = (ends - offs).astype(np.int64)
            if (lens > N_SAMPLES).any():
                raise ValueError("a VAD window is longer than 30 s; the VAD stage must bound turn duration (chunk_size)")
            if _resident is not None:
                results = self.run_device_batches(_resident, offs, lens.astype(np.int32), int(batch_size), options,
                                                  print_progress, combined_progress, _forced_eot_len)
            else:
                results = self._run_batches(audio, offs, lens.astype(np.int32), int(batch_size), options, print_progress,
                                            combined_progress, _forced_eot_len)
            for idx, (text, toks) in enumerate(results):
                if verbose:
                    print(f"Transcript: [{round(vad_segments[idx]['start'], 3)} --> {round(vad_segments[idx]['end'], 3)}] {text}")
                segments.append({"text": text, "start": round(vad_segments[idx]["start"], 3),
                                 "end": round(vad_segments[idx]["end"], 3), "tokens": toks})
        return segments

    # ---- batched execution over one or more replicas
    def upload(self, audio: np.ndarray, offs: np.ndarray, lens: np.ndarray) -> Dict:
        """One H2D copy per GPU of the span of the waveform the windows cover (pinned staging)."""
        lo = int(offs.min())
        hi = int((offs + lens).max())
        host = torch.from_numpy(audio[lo:hi])
        if not host.is_pinned():
            try:
                host = host.pin_memory()
            except RuntimeError:
                pass
        out = {}
        for dev in {rep.device for rep in self.replicas}:
            with torch.cuda.device(dev):
                out[dev] = host.to(dev, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
        return {"lo": lo, "audio": out}

    def run_device_batches(self, resident: Dict, offs, lens, batch_size, options=None, print_progress=False,
                           combined_progress=False, forced_eot_len=0, batch_order: Optional[Sequence[int]] = None):
        """Transcribes windows whose audio is already resident in HBM.  Batches are handed out dynamically to the
        replicas (one host thread + one stream each); results come back in window order."""
        options = options or self.options
        n = len(offs)
        # ceil(n / batch_size) batches as upstream, but of even size (28,27,27,27,27 instead of 32,32,32,32,8): rows are
        # independent, so results are unchanged, and no replica is left with a latency-bound stub batch at the end
        # (rounding the batch count up to a multiple of the replicas - 8 x 17 instead of 5 x 28 on 4 streams - was measured
        # slower: smaller batches re-read the decoder weights more often per window than the lone tail batch costs)
        n_b = max(1, -(-n // batch_size))
        edges = [round(i * n / n_b) for i in range(n_b + 1)]
        batches = [(edges[i], edges[i + 1]) for i in range(n_b) if edges[i + 1] > edges[i]]
        if batch_order is not None:
            batches = [batches[i] for i in batch_order]
        out: List = [None] * n
        tokenizer = self.tokenizer
        n_rep = max(1, min(len(self.replicas), len(batches)))
        errors: List[BaseException] = []
        state = {"next": 0, "done": 0}
        lock = threading.Lock()
        lo = resident["lo"]

        def worker(rep_idx: int):
            rep = self.replicas[rep_idx]
            try:
                with torch.cuda.device(rep.device):
                    if batch_size > rep.engine.max_batch:
                        raise ValueError(f"batch_size={batch_size} exceeds the replica's max_batch={rep.engine.max_batch}")
                    # the only replica working on its GPU for this job: build the decode step for latency (mw_set_solo)
                    rep.engine.set_solo(sum(1 for r in self.replicas[:n_rep] if r.device == rep.device) == 1)
                    stream = rep.stream if (rep.stream is not None and n_rep > 1) else torch.cuda.current_stream(rep.device)
                    d_audio = resident["audio"][rep.device]
                    with torch.cuda.stream(stream):
                        while True:
                            with lock:
                                i = state["next"]
                                state["next"] += 1
                            if i >= len(batches):
                                break
                            a, b = batches[i]
                            d_off = torch.from_numpy(offs[a:b] - lo).to(rep.device, non_blocking=True)
                            d_len = torch.from_numpy(np.ascontiguousarray(lens[a:b], dtype=np.int32)).to(rep.device, non_blocking=True)
                            texts, toks = rep.transcribe_windows(d_audio, d_off, d_len, tokenizer, options, forced_eot_len)
                            for k in range(b - a):
                                out[a + k] = (texts[k], toks[k])
                            with lock:
                                state["done"] += b - a
                                if print_progress:
                                    pct = state["done"] / n * 100
                                    print(f"Progress: {pct / 2 if combined_progress else pct:.2f}%...")
                        stream.synchronize()
            except BaseException as e:      # surfaced on the calling thread
                errors.append(e)

        if n_rep == 1:
            worker(0)
        else:
            threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_rep)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        self.last_stats = {"windows": n, "batches": len(batches), "replicas": n_rep}
        return out

    def _run_batches(self, audio, offs, lens, batch_size, options, print_progress, combined_progress, forced_eot_len):
        resident = self.upload(audio, offs, lens)
        return self.run_device_batches(resident, offs, lens, batch_size, options, print_progress, combined_progress,
                                       forced_eot_len)


def load_model(whisper_arch: str, device: str, device_index=0, compute_type: str = "float16",
               asr_options: Optional[dict] = None, language: Optional[str] = None, vad_model=None,
               vad_method: Optional[str] = "pyannote", vad_options: Optional[dict] = None,
               model: Optional[Union[WhisperModel, dict]] = None, task: str = "transcribe",
               download_root: Optional[str] = None, local_files_only: bool = False, threads: int = 4,
               *, max_batch: int = 32, streams_per_device: int = 2, init_scheme: str = "survey", init_seed: int = 1234,
               dims: Optional[ModelDims] = None, tokens: Optional[SpecialTokens] = None,
               tokenizer_file: Optional[str] = None) -> FasterWhisperPipeline:
    """``whisperx.load_model`` for the B200 engine.

    Arguments up to ``threads`` are whisperx's.  ``device`` must be "cuda": the shipped
    reference sets DEVICE="cpu" (/root/reference/transcribe.py:30) and tells GPU users to change that
    constant (/root/reference/README.md:101) — there is no CPU path here.  ``compute_type`` is accepted for
    signature parity; the engine always computes on fp16 storage (the reference's GPU compute type, /root/reference/transcribe_colab.ipynb:119) with fp32 accumulation.  ``model`` may be a ready
    WhisperModel, an HF-named state dict, or the path of a Hugging Face ``model.safetensors`` (also looked up under
    ``download_root``); with none of these, seeded random-init weights are used because no checkpoint can be downloaded
    offline (a warning says so).  Keyword-only arguments are additions;
    ``streams_per_device`` replicas per GPU share one copy of the weights and keep that many batches in flight.
    """
    if whisper_arch.endswith(".en"):
        # English-only checkpoints have their own vocabulary (51864 ids, <|endoftext|> 50256, no language/task tokens in the
        # sot sequence); running one with the multilingual control ids would be silently wrong, so it is refused
        raise ValueError(f"'{whisper_arch}': English-only checkpoints are not supported by this engine; the reference exposes "
                         "the multilingual sizes tiny / base / small / medium / large-v3 (/root/reference/README.md:85)")
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    if dev.type != "cuda":
        raise ValueError(f"unsupported device {device!r}: the B200-native engine has no CPU path; pass device='cuda'")
    if compute_type not in _COMPUTE_TYPES:
        raise ValueError(f"Requested compute type {compute_type!r} is not a valid compute type")
    if compute_type not in ("float16", "bfloat16", "default", "auto"):
        warnings.warn(f"compute_type={compute_type!r} requested; the sm_100a engine computes on 16-bit storage (float16) with fp32 accumulation")
    if vad_model is None and vad_method not in ("pyannote", "silero", "energy", "energy_gpu", None):
        raise ValueError(f"Invalid vad_method: {vad_method}")
    mdims = dims or model_dims(whisper_arch)
    toks = tokens or special_tokens(mdims.vocab)
    indices = list(device_index) if isinstance(device_index, (list, tuple)) else [int(device_index)]

    default_asr_options = TranscriptionOptions()
    if asr_options is not None:
        unknown = set(asr_options) - set(default_asr_options.__dataclass_fields__)
        if unknown:
            raise TypeError(f"TranscriptionOptions got unexpected option(s): {sorted(unknown)}")
        default_asr_options = replace(default_asr_options, **asr_options)
    suppress_numerals = default_asr_options.suppress_numerals
    max_beam = max(1, min(8, int(default_asr_options.beam_size)))

    if isinstance(model, WhisperModel):
        replicas = [model]
    else:
        ckpt = model if isinstance(model, str) else None
        if ckpt is None and download_root:
            import os
            for cand in (os.path.join(download_root, whisper_arch, "model.safetensors"), os.path.join(download_root, "model.safetensors")):
                if os.path.exists(cand):
                    ckpt = cand
                    break
        if isinstance(model, dict):
            sd = model
        elif ckpt is not None:
            # a Hugging Face Whisper checkpoint (keys of WhisperForConditionalGeneration.state_dict()), e.g. openai/whisper-large-v3
            from safetensors.torch import load_file
            sd = load_file(ckpt)
            if "model.encoder.embed_positions.weight" not in sd:
                raise ValueError(f"{ckpt} is not a Hugging Face Whisper checkpoint (missing model.encoder.* keys)")
        else:
            warnings.warn(f"no '{whisper_arch}' checkpoint is reachable offline: using seeded random-init weights "
                          f"(scheme={init_scheme!r}, seed={init_seed}); transcripts are token ids, not text")
            sd = random_init(mdims, seed=init_seed, scheme=init_scheme)
        replicas = []
        for i in indices:
            first = WhisperModel(mdims, sd, device_index=i, max_batch=max_batch, max_beam=max_beam, tokens=toks)
            first.stream = torch.cuda.Stream(device=first.device) if streams_per_device > 1 else None
            replicas.append(first)
            # extra replicas on the same GPU share the weights and add a workspace + stream each, so that the launch
            # gaps of one batch's decode steps are filled by another batch's kernels
            for _ in range(max(1, int(streams_per_device)) - 1):
                replicas.append(WhisperModel(mdims, None, device_index=i, max_batch=max_batch, max_beam=max_beam,
                                             tokens=toks, share_weights_with=first))

    tokenizer = None
    hf_tokenizer = None
    if tokenizer_file:
        import tokenizers
        hf_tokenizer = tokenizers.Tokenizer.from_file(tokenizer_file)
    if language is not None:
        tokenizer = Tokenizer(toks, True, task=task, language=language, hf=hf_tokenizer)
    else:
        print("No language specified, language will be first be detected for each audio file (increases inference time).")

    default_vad_options = dict(DEFAULT_VAD_OPTIONS)
    if vad_options is not None:
        default_vad_options.update(vad_options)
    if vad_model is None:
        if vad_method in ("pyannote", "silero"):
            warnings.warn(f"vad_method={vad_method!r}: the {vad_method} network's weights are not available offline; "
                          "using the built-in energy VAD with the same onset/offset/chunk_size knobs")
        cls = GpuEnergyVad if vad_method == "energy_gpu" else EnergyVad
        vad_model = cls(vad_onset=default_vad_options["vad_onset"], vad_offset=default_vad_options["vad_offset"],
                        chunk_size=default_vad_options["chunk_size"])
    return FasterWhisperPipeline(model=replicas, vad=vad_model, options=default_asr_options, tokenizer=tokenizer,
                                 language=language, suppress_numerals=suppress_numerals, vad_params=default_vad_options,
                                 hf_tokenizer=hf_tokenizer)
