"""Forced alignment on the GPU: the host mirror of ``whisperx.load_align_model`` / ``whisperx.align``
(/root/reference/transcribe.py:127-135).

    model_a, metadata = load_align_model(language_code="zh", device="cuda")
    result = align(result["segments"], model_a, metadata, audio, "cuda", return_char_alignments=False)

whisperx runs its wav2vec2-CTC model on one segment at a time and walks the CTC trellis in Python on the host; here all
segments of a batch go through one ragged launch list (csrc/w2v.cu: ``mw_w2v_emissions``) and the trellis + backtrack run on
the device (``mw_ctc_align``), one CTA per segment.  The character/word/sentence book-keeping around it follows
whisperx/alignment.py [UPSTREAM-MEMORY] and stays on the host.  There is no CPU fallback.

Deviations, all stated: (1) no checkpoint or vocabulary can be downloaded offline, so without ``model_dir`` /
``model=`` the weights are seeded random-init of the XLSR-53 architecture and the dictionary is the 32-symbol
wav2vec2 character set - characters outside it align as '*' wildcards, exactly as whisperx treats them; (2) sentence
splitting uses nltk's Punkt tokenizer when nltk is importable (as whisperx does) and otherwise keeps each segment as one
sentence; pass ``sentence_splitter=`` to override.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import warnings
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .config import SAMPLE_RATE
from .w2v import (W2vDims, random_init_w2v, effective_pos_conv_weight, torchaudio_to_hf, is_torchaudio_state_dict, base_dims,
                  BASE_LANGUAGES)

LANGUAGES_WITHOUT_SPACES = ["ja", "zh"]
WILDCARD = -1

# the character set of facebook/wav2vec2-base-960h / the XLSR fine-tunes' usual layout: <pad> (CTC blank) first
DEFAULT_DICTIONARY = {c: i for i, c in enumerate(
    ["<pad>", "<s>", "</s>", "<unk>", "|", "e", "t", "a", "o", "n", "i", "h", "s", "r", "d", "l", "u", "m", "w", "c", "f", "g",
     "y", "p", "b", "v", "k", "'", "x", "j", "q", "z"])}

# architecture used when no checkpoint is given (None = XLSR-53 large); tests shrink it
DEFAULT_ALIGN_DIMS: Optional[W2vDims] = None

_LAYER = ["ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2"]
_N_GLOBAL = 38      # enum mw_w2v_weight_id: MW_A_GLOBAL_COUNT


def pack_w2v_weights(sd: Dict[str, torch.Tensor], dims: W2vDims, device: torch.device) -> List[torch.Tensor]:
    """Hugging Face ``Wav2Vec2ForCTC`` state dict -> the engine's weight table (order of enum mw_w2v_weight_id, then the
    per-layer blocks of enum mw_enc_layer_weight_id).  wav2vec2-base checkpoints (no conv biases, GroupNorm after conv0 only)
    fill the slots the variant does not read with zeros / ones; positional-conv groups narrower than 64 channels get their
    weight rows padded to 64 per group."""
    zeros_c, ones_c = torch.zeros(dims.conv_dim), torch.ones(dims.conv_dim)
    def mat(t):
        return t.to(device=device, dtype=_lib.storage_dtype()).contiguous()

    def vec(t):
        return t.to(device=device, dtype=torch.float32).contiguous()

    out: List[torch.Tensor] = []
    fe = "wav2vec2.feature_extractor.conv_layers."
    for i in range(7):
        w = sd[f"{fe}{i}.conv.weight"]
        if i == 0:
            out.append(vec(w[:, 0, :]))                                            # f32 [C, 10]
        else:
            out.append(mat(w.permute(0, 2, 1).reshape(w.shape[0], -1)))             # [co][tap][ci]
        out += [vec(sd.get(f"{fe}{i}.conv.bias", zeros_c)), vec(sd.get(f"{fe}{i}.layer_norm.weight", ones_c)),
                vec(sd.get(f"{fe}{i}.layer_norm.bias", zeros_c))]
    fp = "wav2vec2.feature_projection."
    out += [vec(sd[fp + "layer_norm.weight"]), vec(sd[fp + "layer_norm.bias"]), mat(sd[fp + "projection.weight"]),
            vec(sd[fp + "projection.bias"])]
    G, gs, kp = dims.pos_groups, dims.d_model // dims.pos_groups, dims.pos_kernel
    wp = effective_pos_conv_weight(sd).view(G, gs, gs, kp).permute(0, 1, 3, 2)      # [g][out][tap][in]
    if gs != 64:                                                                    # 64 (zero-padded) output rows per group
        wp = torch.cat([wp, torch.zeros(G, 64 - gs, kp, gs)], dim=1)
    out += [mat(wp), vec(sd["wav2vec2.encoder.pos_conv_embed.conv.bias"])]
    out += [vec(sd["wav2vec2.encoder.layer_norm.weight"]), vec(sd["wav2vec2.encoder.layer_norm.bias"])]
    vp = (dims.vocab + 31) // 32 * 32
    lm_w = torch.zeros(vp, dims.d_model)
    lm_w[: dims.vocab] = sd["lm_head.weight"].float().cpu()
    lm_b = torch.zeros(vp)
    lm_b[: dims.vocab] = sd["lm_head.bias"].float().cpu()
    out += [mat(lm_w), vec(lm_b)]
    assert len(out) == _N_GLOBAL
    for l in range(dims.n_layers):
        p = f"wav2vec2.encoder.layers.{l}."
        a = p + "attention."
        out += [vec(sd[p + "layer_norm.weight"]), vec(sd[p + "layer_norm.bias"]),
                mat(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)),
                vec(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)),
                mat(sd[a + "out_proj.weight"]), vec(sd[a + "out_proj.bias"]),
                vec(sd[p + "final_layer_norm.weight"]), vec(sd[p + "final_layer_norm.bias"]),
                mat(sd[p + "feed_forward.intermediate_dense.weight"]), vec(sd[p + "feed_forward.intermediate_dense.bias"]),
                mat(sd[p + "feed_forward.output_dense.weight"]), vec(sd[p + "feed_forward.output_dense.bias"])]
    assert len(out) == _N_GLOBAL + dims.n_layers * len(_LAYER)
    return out


class AlignEngine:
    """Host wrapper of the opaque ``mw_w2v`` (include/mw_b200.h)."""

    def __init__(self, dims: W2vDims, sd: Dict[str, torch.Tensor], device_index: int = 0, max_batch: int = 16,
                 max_samples: int = 30 * SAMPLE_RATE):
        xlsr = dims.feat_norm == "layer" and dims.stable_layer_norm and dims.conv_bias
        base = dims.feat_norm == "group" and not dims.stable_layer_norm and not dims.conv_bias
        if not (xlsr or base):
            raise NotImplementedError("the CUDA alignment engine implements the two wav2vec2 families whisperx loads: layer-norm / "
                                      "stable-layer-norm / conv-bias (XLSR-53) and group-norm / post-layer-norm / no conv bias "
                                      f"(wav2vec2-base); got feat_norm={dims.feat_norm!r} stable_layer_norm={dims.stable_layer_norm} "
                                      f"conv_bias={dims.conv_bias}")
        if not torch.cuda.is_available():
            raise RuntimeError("manual_whisper_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.dims = dims
        self.device = torch.device("cuda", device_index)
        self.max_batch, self.max_samples = int(max_batch), int(max_samples)
        self.weights = pack_w2v_weights(sd, dims, self.device)
        cfg = _lib.W2vConfigC(dims.n_layers, dims.d_model, dims.n_heads, dims.ffn, dims.vocab, dims.conv_dim, dims.pos_kernel,
                              dims.pos_groups, self.max_batch, self.max_samples, device_index, 0 if xlsr else 1)
        ptrs = (C.c_void_p * len(self.weights))(*[w.data_ptr() for w in self.weights])
        table = _lib.WeightTableC(len(self.weights), ptrs)
        handle = C.c_void_p()
        _lib.check(self.lib.mw_w2v_create(C.byref(cfg), C.byref(table), C.byref(handle)), "mw_w2v_create")
        self.handle = handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.mw_w2v_destroy(h)

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.mw_w2v_workspace_bytes(self.handle))

    def frames(self, n_samples: int) -> int:
        return int(self.lib.mw_w2v_frames(int(max(n_samples, 400))))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def emissions(self, d_audio: torch.Tensor, offs: np.ndarray, lens: np.ndarray) -> Tuple[torch.Tensor, np.ndarray]:
        """log_softmax emissions of the windows d_audio[offs[c] : offs[c] + lens[c]] -> (f32 [n, T, vocab] on the device,
        valid frames per window); rows beyond a window's frames are zero."""
        n = len(offs)
        if not (0 < n <= self.max_batch):
            raise ValueError(f"{n} windows outside 1..max_batch={self.max_batch}")
        lens32 = np.ascontiguousarray(lens, dtype=np.int32)
        frames = np.array([self.frames(int(x)) for x in lens32], dtype=np.int32)
        T = int(frames.max())
        with torch.cuda.device(self.device):
            d_off = torch.from_numpy(np.ascontiguousarray(offs, dtype=np.int64)).to(self.device)
            d_len = torch.from_numpy(lens32).to(self.device)
            out = torch.zeros(n, T, self.dims.vocab, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.mw_w2v_emissions(self.handle, d_audio.data_ptr(), d_audio.numel(), d_off.data_ptr(),
                                                 d_len.data_ptr(), lens32.ctypes.data_as(_lib.c_i32p), n, out.data_ptr(),
                                                 T * self.dims.vocab, self._stream()), "mw_w2v_emissions")
        return out, frames

    def ctc_align(self, emissions: torch.Tensor, frames: np.ndarray, tokens: Sequence[Sequence[int]], blank: int):
        """whisperx get_trellis + backtrack on the device -> (frame_token i32 [n, T], frame_score f32 [n, T], ok bool [n])
        as numpy arrays: the token index each frame belongs to and the probability of the symbol emitted there."""
        n, T, V = emissions.shape
        max_tok = max(1, max(len(t) for t in tokens))
        tok = np.zeros((n, max_tok), dtype=np.int32)
        for i, t in enumerate(tokens):
            tok[i, : len(t)] = t
        n_tok = np.array([len(t) for t in tokens], dtype=np.int32)
        with torch.cuda.device(self.device):
            d_tok = torch.from_numpy(tok).to(self.device)
            d_ntok = torch.from_numpy(n_tok).to(self.device)
            d_frames = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.int32)).to(self.device)
            f_tok = torch.empty(n, T, dtype=torch.int32, device=self.device)
            f_sc = torch.empty(n, T, dtype=torch.float32, device=self.device)
            ok = torch.empty(n, dtype=torch.int32, device=self.device)
            ws = torch.empty(n * T * (max_tok + 1), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.mw_ctc_align(emissions.data_ptr(), T * V, V, d_frames.data_ptr(), d_tok.data_ptr(), max_tok,
                                             d_ntok.data_ptr(), n, int(blank), f_tok.data_ptr(), f_sc.data_ptr(), T,
                                             ok.data_ptr(), ws.data_ptr(), self._stream()), "mw_ctc_align")
            return f_tok.cpu().numpy(), f_sc.cpu().numpy(), ok.cpu().numpy().astype(bool)


class AlignModel:
    """What ``load_align_model`` returns in place of the torch wav2vec2 module."""

    def __init__(self, engine: AlignEngine, dictionary: Dict[str, int], language: str):
        self.engine, self.dictionary, self.language = engine, dictionary, language


def load_align_model(language_code: str, device: str, model_name: Optional[str] = None, model_dir: Optional[str] = None, *,
                     model: Optional[Union[dict, str]] = None, dictionary: Optional[Dict[str, int]] = None,
                     dims: Optional[W2vDims] = None, device_index: int = 0, max_batch: int = 16,
                     max_samples: int = 30 * SAMPLE_RATE, init_seed: int = 4321):
    """whisperx.load_align_model(language_code, device, model_name=None, model_dir=None) -> (model, metadata).

    A checkpoint is taken from ``model=`` (state dict or ``model.safetensors`` path) or ``model_dir`` (``model.safetensors``
    and ``vocab.json`` as saved by Hugging Face); with neither, seeded random-init weights are used (a warning says so)."""
    if not str(device).startswith("cuda"):
        raise ValueError(f"device={device!r}: this engine runs on CUDA (B200) only")
    if ":" in str(device):
        device_index = int(str(device).split(":")[1])
    ckpt = model if isinstance(model, str) else None
    if ckpt is None and model_dir and os.path.exists(os.path.join(model_dir, "model.safetensors")):
        ckpt = os.path.join(model_dir, "model.safetensors")
    if dictionary is None and model_dir and os.path.exists(os.path.join(model_dir, "vocab.json")):
        with open(os.path.join(model_dir, "vocab.json"), encoding="utf-8") as f:
            dictionary = {k.lower(): int(v) for k, v in json.load(f).items()}
    if dictionary is None:
        dictionary = dict(DEFAULT_DICTIONARY)
    if isinstance(model, dict):
        sd = torchaudio_to_hf(model) if is_torchaudio_state_dict(model) else model      # torchaudio bundles: whisperx's en/fr/de/es/it
    elif ckpt is not None:
        from safetensors.torch import load_file
        sd = load_file(ckpt)
        if "lm_head.weight" not in sd or "wav2vec2.feature_extractor.conv_layers.0.layer_norm.weight" not in sd:
            raise ValueError(f"{ckpt} is not a Hugging Face Wav2Vec2ForCTC checkpoint")
    else:
        sd = None
    if dims is None:
        if sd is not None:
            n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("wav2vec2.encoder.layers."))
            d = sd["lm_head.weight"].shape[1]
            fe = "wav2vec2.feature_extractor.conv_layers."
            layer_norm_family = fe + "1.layer_norm.weight" in sd          # wav2vec2-base normalises conv layer 0 only
            dims = W2vDims(name=model_name or "checkpoint", n_layers=n_layers, d_model=d, n_heads=d // 64,
                           ffn=sd["wav2vec2.encoder.layers.0.feed_forward.intermediate_dense.weight"].shape[0],
                           vocab=sd["lm_head.weight"].shape[0], feat_norm="layer" if layer_norm_family else "group",
                           stable_layer_norm=layer_norm_family, conv_bias=fe + "0.conv.bias" in sd)
        elif DEFAULT_ALIGN_DIMS is not None:
            dims = DEFAULT_ALIGN_DIMS
        elif language_code in BASE_LANGUAGES:
            # whisperx's DEFAULT_ALIGN_MODELS_TORCH: the torchaudio wav2vec2-base bundles (group-norm, post-LayerNorm family)
            dims = base_dims(vocab=max(dictionary.values()) + 1)
        else:
            dims = W2vDims(vocab=max(dictionary.values()) + 1)
    if sd is None:
        warnings.warn("no wav2vec2 checkpoint is available offline: using seeded random-init weights "
                      f"({dims.name}, seed {init_seed}); alignments are structurally valid but not meaningful")
        sd = random_init_w2v(dims, seed=init_seed)
    engine = AlignEngine(dims, sd, device_index=device_index, max_batch=max_batch, max_samples=max_samples)
    metadata = {"language": language_code, "dictionary": dictionary, "type": "b200"}
    return AlignModel(engine, dictionary, language_code), metadata


# ---------------------------------------------------------------------------------------------------------------------
def _default_sentence_splitter() -> Callable[[str], List[Tuple[int, int]]]:
    try:
        from nltk.tokenize.punkt import PunktParameters, PunktSentenceTokenizer     # as whisperx does
        params = PunktParameters()
        params.abbrev_types = {"dr", "vs", "mr", "mrs", "prof"}
        tok = PunktSentenceTokenizer(params)
        return lambda text: list(tok.span_tokenize(text))
    except Exception:
        return lambda text: [(0, len(text))]


def interpolate_nans(values: List[float], method: str = "nearest") -> List[float]:
    """whisperx.utils.interpolate_nans on a list: NaNs between known values take the nearest (ties: the earlier) known
    value, or the linear interpolation; leading/trailing NaNs are back/forward filled."""
    known = [i for i, v in enumerate(values) if not math.isnan(v)]
    if not known:
        return list(values)
    out = list(values)
    for i, v in enumerate(values):
        if not math.isnan(v):
            continue
        left = max((k for k in known if k < i), default=None)
        right = min((k for k in known if k > i), default=None)
        if left is None:
            out[i] = values[right]
        elif right is None:
            out[i] = values[left]
        elif method == "linear":
            out[i] = values[left] + (values[right] - values[left]) * (i - left) / (right - left)
        else:
            out[i] = values[left] if i - left <= right - i else values[right]
    return out


def _nanmin(xs: Iterable[Optional[float]]) -> float:
    v = [x for x in xs if x is not None]
    return min(v) if v else float("nan")


def _nanmax(xs: Iterable[Optional[float]]) -> float:
    v = [x for x in xs if x is not None]
    return max(v) if v else float("nan")


def preprocess_segment(text: str, dictionary: Dict[str, int], language: str):
    """Step 1 of whisperx.align: the characters that take part in the alignment and where they sit in the text.
    Returns (clean_char, clean_cdx, tokens) - characters outside the dictionary become '*' (token -1)."""
    num_leading = len(text) - len(text.lstrip())
    num_trailing = len(text) - len(text.rstrip())
    clean_char, clean_cdx = [], []
    for cdx, char in enumerate(text):
        char_ = char.lower()
        if language not in LANGUAGES_WITHOUT_SPACES:
            char_ = char_.replace(" ", "|")
        if cdx < num_leading or cdx > len(text) - num_trailing - 1:
            continue
        clean_char.append(char_ if char_ in dictionary else "*")
        clean_cdx.append(cdx)
    tokens = [dictionary.get(c, WILDCARD) for c in clean_char]
    return clean_char, clean_cdx, tokens


def chars_from_path(frame_token: np.ndarray, frame_score: np.ndarray, n_frames: int, n_tokens: int):
    """whisperx merge_repeats on the per-frame form: token j -> (first frame, last frame + 1, mean frame score)."""
    out = []
    t = 0
    while t < n_frames:
        j = int(frame_token[t])
        t2 = t
        while t2 < n_frames and int(frame_token[t2]) == j:
            t2 += 1
        out.append((j, t, t2, float(np.mean(frame_score[t:t2], dtype=np.float64))))
        t = t2
    return out if len(out) == n_tokens else None


def assemble_segment(segment: dict, text: str, clean_cdx: List[int], char_spans, ratio: float, language: str,
                     sentence_spans: List[Tuple[int, int]], interpolate_method: str, return_char_alignments: bool) -> List[dict]:
    """Steps after the backtrack in whisperx.align: per-character times -> words -> sentences (sub-segments)."""
    t1 = segment["start"]
    pos_of = {cdx: k for k, cdx in enumerate(clean_cdx)}
    chars = []
    word_idx = 0
    for cdx, char in enumerate(text):
        start = end = score = None
        if cdx in pos_of:
            _, f0, f1, sc = char_spans[pos_of[cdx]]
            start, end, score = round(f0 * ratio + t1, 3), round(f1 * ratio + t1, 3), round(sc, 3)
        chars.append({"char": char, "start": start, "end": end, "score": score, "word-idx": word_idx})
        if language in LANGUAGES_WITHOUT_SPACES:
            word_idx += 1
        elif cdx == len(text) - 1 or text[cdx + 1] == " ":
            word_idx += 1
    subs = []
    for sstart, send in sentence_spans:
        curr = [c for i, c in enumerate(chars) if sstart <= i <= send]          # inclusive on both ends, as upstream's .loc
        sentence = {"text": text[sstart:send], "start": _nanmin(c["start"] for c in curr),
                    "end": _nanmax(c["end"] for c in curr if c["char"] != " "), "words": []}
        seen = []
        for c in curr:
            if c["word-idx"] not in seen:
                seen.append(c["word-idx"])
        for w in seen:
            wc = [c for c in curr if c["word-idx"] == w]
            word_text = "".join(c["char"] for c in wc).strip()
            if not word_text:
                continue
            wc = [c for c in wc if c["char"] != " "]
            word = {"word": word_text}
            ws, we = _nanmin(c["start"] for c in wc), _nanmax(c["end"] for c in wc)
            scores = [c["score"] for c in wc if c["score"] is not None]
            if not math.isnan(ws):
                word["start"] = ws
            if not math.isnan(we):
                word["end"] = we
            if scores:
                word["score"] = round(sum(scores) / len(scores), 3)
            sentence["words"].append(word)
        if return_char_alignments:
            sentence["chars"] = [{k: v for k, v in (("char", c["char"]), ("start", c["start"]), ("end", c["end"]),
                                                    ("score", c["score"])) if v is not None} for c in curr]
        subs.append(sentence)
    starts = interpolate_nans([s["start"] for s in subs], interpolate_method)
    ends = interpolate_nans([s["end"] for s in subs], interpolate_method)
    for s, a, b in zip(subs, starts, ends):
        s["start"], s["end"] = a, b
    # sentences sharing both timestamps are concatenated; groups come out sorted by (start, end); NaN keys are dropped
    groups: Dict[Tuple[float, float], dict] = {}
    joiner = "" if language in LANGUAGES_WITHOUT_SPACES else " "
    for s in subs:
        if math.isnan(s["start"]) or math.isnan(s["end"]):
            continue
        key = (s["start"], s["end"])
        if key not in groups:
            groups[key] = {"start": s["start"], "end": s["end"], "text": s["text"], "words": list(s["words"])}
            if return_char_alignments:
                groups[key]["chars"] = list(s["chars"])
        else:
            groups[key]["text"] = groups[key]["text"] + joiner + s["text"]
            groups[key]["words"] += s["words"]
            if return_char_alignments:
                groups[key]["chars"] += s["chars"]
    return [groups[k] for k in sorted(groups)]


def align(transcript: Iterable[dict], model: AlignModel, align_model_metadata: dict, audio, device: str = "cuda",
          interpolate_method: str = "nearest", return_char_alignments: bool = False, print_progress: bool = False,
          combined_progress: bool = False, *, sentence_splitter: Optional[Callable[[str], List[Tuple[int, int]]]] = None,
          _keep_index: bool = False) -> dict:
    """whisperx.align(transcript, model, align_model_metadata, audio, device, ...) ->
    {"segments": [{"start","end","text","words":[{"word","start","end","score"}], ("chars")}], "word_segments": [...]}."""
    if isinstance(audio, str):
        from .audio import load_audio
        audio = load_audio(audio)
    eng = model.engine
    if torch.is_tensor(audio):
        d_audio = audio.to(device=eng.device, dtype=torch.float32).reshape(-1).contiguous()
    else:
        d_audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)).to(eng.device)
    n_audio = d_audio.numel()
    max_duration = n_audio / SAMPLE_RATE
    dictionary, language = align_model_metadata["dictionary"], align_model_metadata["language"]
    blank = 0
    for ch, code in dictionary.items():
        if ch in ("[pad]", "<pad>"):
            blank = code
    splitter = sentence_splitter or _default_sentence_splitter()
    transcript = list(transcript)
    prepared = []
    for seg in transcript:
        text = seg["text"]
        clean_char, clean_cdx, tokens = preprocess_segment(text, dictionary, language)
        prepared.append((clean_char, clean_cdx, tokens, splitter(text)))

    out_segments: List[Optional[List[dict]]] = [None] * len(transcript)
    todo = []
    for i, seg in enumerate(transcript):
        t1, t2, text = seg["start"], seg["end"], seg["text"]
        fallback = {"start": t1, "end": t2, "text": text, "words": []}
        if return_char_alignments:
            fallback["chars"] = []
        if len(prepared[i][0]) == 0:
            print(f'Failed to align segment ("{text}"): no characters in this segment found in model dictionary, resorting to original...')
            out_segments[i] = [fallback]
        elif t1 >= max_duration:
            print(f'Failed to align segment ("{text}"): original start time longer than audio duration, skipping...')
            out_segments[i] = [fallback]
        else:
            f1, f2 = int(t1 * SAMPLE_RATE), min(int(t2 * SAMPLE_RATE), n_audio)
            if f2 - f1 > eng.max_samples:
                raise ValueError(f"segment {i} spans {f2 - f1} samples, more than the align model's max_samples={eng.max_samples}")
            todo.append((i, f1, max(f2 - f1, 0)))
    done = 0
    for b0 in range(0, len(todo), eng.max_batch):
        part = todo[b0: b0 + eng.max_batch]
        offs = np.array([p[1] for p in part], dtype=np.int64)
        lens = np.array([p[2] for p in part], dtype=np.int32)
        em, frames = eng.emissions(d_audio, offs, lens)
        f_tok, f_sc, ok = eng.ctc_align(em, frames, [prepared[p[0]][2] for p in part], blank)
        for k, (i, f1, n_samp) in enumerate(part):
            seg = transcript[i]
            clean_char, clean_cdx, tokens, spans = prepared[i]
            T = int(frames[k])
            spans_c = chars_from_path(f_tok[k], f_sc[k], T, len(tokens)) if ok[k] else None
            if spans_c is None:
                print(f'Failed to align segment ("{seg["text"]}"): backtrack failed, resorting to original...')
                fb = {"start": seg["start"], "end": seg["end"], "text": seg["text"], "words": []}
                if return_char_alignments:
                    fb["chars"] = []
                out_segments[i] = [fb]
                continue
            # upstream: ratio = duration * waveform_segment.size(0) / (trellis.size(0) - 1), with a [1, n] waveform
            ratio = (seg["end"] - seg["start"]) / max(T - 1, 1)
            out_segments[i] = assemble_segment(seg, seg["text"], clean_cdx, spans_c, ratio, language, spans, interpolate_method,
                                               return_char_alignments)
        done += len(part)
        if print_progress:
            pct = done / max(len(todo), 1) * 100
            print(f"Progress: {50 + pct / 2 if combined_progress else pct:.2f}%...")
    if _keep_index:         # distributed.align_sharded: tag every output sub-segment with the input segment it came from
        for seg, group in zip(transcript, out_segments):
            for s in group or []:
                s["_idx"] = seg.get("_idx")
    segments = [s for group in out_segments for s in (group or [])]
    word_segments = [w for s in segments for w in s["words"]]
    return {"segments": segments, "word_segments": word_segments}
