"""Host wrapper of the opaque ``mw_model`` (include/mw_b200.h): weight packing, encode, generate.

Plays the role of ``ctranslate2.models.Whisper`` as held by ``faster_whisper.WhisperModel`` inside whisperx
(SURVEY.md §2.2 "Engine"; reached from /root/reference/transcribe.py:107-113,123).  PyTorch is used for
device memory and streams only; all arithmetic happens in libmw_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import ModelDims, SpecialTokens

# order must match enum mw_weight_id / mw_enc_layer_weight_id / mw_dec_layer_weight_id
_GLOBAL = ["conv1_w", "conv1_b", "conv2_w", "conv2_b", "enc_pos", "enc_ln_g", "enc_ln_b",
           "dec_emb", "dec_pos", "dec_ln_g", "dec_ln_b"]
_ENC = ["ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2"]
_DEC = ["ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "lnx_g", "lnx_b", "wxq", "bxq", "wxkv", "bxkv",
        "wxo", "bxo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2"]


def pack_weights(sd: Dict[str, torch.Tensor], dims: ModelDims, device: torch.device) -> List[torch.Tensor]:
    """HF-named state dict -> the engine's weight table (16-bit matrices in the library's storage type - fp16 unless built
    otherwise, include/mw_b200.h: mw_storage_dtype - fp32 vectors, engine layouts)."""
    d = dims.d_model
    h16 = _lib.storage_dtype()

    def mat(t):
        return t.to(device=device, dtype=h16).contiguous()

    def vec(t):
        return t.to(device=device, dtype=torch.float32).contiguous()

    def conv(w):  # [co, ci, 3] -> [co, tap, ci] flattened: K index = tap*ci_count + ci
        return mat(w.permute(0, 2, 1).reshape(w.shape[0], -1))

    zeros_d = torch.zeros(d, device=sd["model.encoder.conv1.bias"].device)
    out: List[torch.Tensor] = []
    e = "model.encoder."
    out += [conv(sd[e + "conv1.weight"]), vec(sd[e + "conv1.bias"]), conv(sd[e + "conv2.weight"]),
            vec(sd[e + "conv2.bias"]), vec(sd[e + "embed_positions.weight"][: dims.n_audio_ctx]),
            vec(sd[e + "layer_norm.weight"]), vec(sd[e + "layer_norm.bias"])]
    dd = "model.decoder."
    out += [mat(sd[dd + "embed_tokens.weight"]), vec(sd[dd + "embed_positions.weight"]),
            vec(sd[dd + "layer_norm.weight"]), vec(sd[dd + "layer_norm.bias"])]
    assert len(out) == len(_GLOBAL)

    def attn_qkv(p):
        w = torch.cat([sd[p + "q_proj.weight"], sd[p + "k_proj.weight"], sd[p + "v_proj.weight"]], 0)
        b = torch.cat([sd[p + "q_proj.bias"], zeros_d, sd[p + "v_proj.bias"]], 0)
        return mat(w), vec(b)

    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        wqkv, bqkv = attn_qkv(p + "self_attn.")
        out += [vec(sd[p + "self_attn_layer_norm.weight"]), vec(sd[p + "self_attn_layer_norm.bias"]), wqkv, bqkv,
                mat(sd[p + "self_attn.out_proj.weight"]), vec(sd[p + "self_attn.out_proj.bias"]),
                vec(sd[p + "final_layer_norm.weight"]), vec(sd[p + "final_layer_norm.bias"]),
                mat(sd[p + "fc1.weight"]), vec(sd[p + "fc1.bias"]), mat(sd[p + "fc2.weight"]), vec(sd[p + "fc2.bias"])]
    for i in range(dims.dec_layers):
        p = f"{dd}layers.{i}."
        wqkv, bqkv = attn_qkv(p + "self_attn.")
        x = p + "encoder_attn."
        wxkv = mat(torch.cat([sd[x + "k_proj.weight"], sd[x + "v_proj.weight"]], 0))
        bxkv = vec(torch.cat([zeros_d, sd[x + "v_proj.bias"]], 0))
        out += [vec(sd[p + "self_attn_layer_norm.weight"]), vec(sd[p + "self_attn_layer_norm.bias"]), wqkv, bqkv,
                mat(sd[p + "self_attn.out_proj.weight"]), vec(sd[p + "self_attn.out_proj.bias"]),
                vec(sd[p + "encoder_attn_layer_norm.weight"]), vec(sd[p + "encoder_attn_layer_norm.bias"]),
                mat(sd[x + "q_proj.weight"]), vec(sd[x + "q_proj.bias"]), wxkv, bxkv,
                mat(sd[x + "out_proj.weight"]), vec(sd[x + "out_proj.bias"]),
                vec(sd[p + "final_layer_norm.weight"]), vec(sd[p + "final_layer_norm.bias"]),
                mat(sd[p + "fc1.weight"]), vec(sd[p + "fc1.bias"]), mat(sd[p + "fc2.weight"]), vec(sd[p + "fc2.bias"])]
    assert len(out) == len(_GLOBAL) + dims.enc_layers * len(_ENC) + dims.dec_layers * len(_DEC)
    return out


def timestamps_enabled(prompt: Sequence[int], tokens: SpecialTokens, without_timestamps: Optional[bool] = None) -> bool:
    """Whether the timestamp logit rules apply to a generate call."""
    if without_timestamps is None:
        prompt = list(prompt)
        tail = prompt[prompt.index(tokens.sot):] if tokens.sot in prompt else prompt
        without_timestamps = tokens.no_timestamps in tail
    return not without_timestamps


@dataclass
class GenerationResult:
    """Mirror of ctranslate2.models.WhisperGenerationResult (the fields whisperx reads)."""
    sequences_ids: List[List[int]]
    scores: List[float]


class Engine:
    """One Whisper replica on one GPU."""

    def __init__(self, dims: ModelDims, state_dict: Optional[Dict[str, torch.Tensor]], device=0, max_batch: int = 32,
                 max_beam: int = 1, packed: Optional[List[torch.Tensor]] = None):
        """`packed` = the weight table of another Engine on the same device (weights are borrowed pointers, so
        several engines — each with its own workspace — can share one copy)."""
        self.lib = _lib.load()
        self.h16 = _lib.storage_dtype()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError("unsupported device %s: the B200 engine runs on CUDA devices only (no CPU fallback)" % device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.dims = dims
        self.max_batch = int(max_batch)
        self.max_beam = int(max_beam)
        with torch.cuda.device(self.device):
            if packed is not None:
                assert all(t.device == self.device for t in packed)
                self.weights = packed
            else:
                self.weights = pack_weights(state_dict, dims, self.device)
            torch.cuda.synchronize()
        ptrs = (C.c_void_p * len(self.weights))(*[t.data_ptr() for t in self.weights])
        table = _lib.WeightTableC(len(self.weights), C.cast(ptrs, C.POINTER(C.c_void_p)))
        cfg = _lib.ModelConfigC(dims.n_mels, dims.d_model, dims.n_heads, dims.enc_layers, dims.dec_layers, dims.ffn,
                                dims.vocab, dims.n_audio_ctx, dims.n_text_ctx, self.max_batch, self.max_beam,
                                self.device.index)
        handle = C.c_void_p()
        _lib.check(self.lib.mw_model_create(C.byref(cfg), C.byref(table), C.byref(handle)), "mw_model_create")
        self.handle = handle
        self._keep = (ptrs, table)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self.lib.mw_model_destroy(h)
            self.handle = None

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.mw_model_workspace_bytes(self.handle))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- S2
    def encode(self, features: torch.Tensor) -> torch.Tensor:
        """features f32 [B, n_mels, 2*n_audio_ctx] (CUDA) -> h16 [B, n_audio_ctx, d]."""
        assert features.is_cuda and features.dtype == torch.float32 and features.is_contiguous()
        B = features.shape[0]
        assert features.shape[1:] == (self.dims.n_mels, 2 * self.dims.n_audio_ctx), features.shape
        out = torch.empty((B, self.dims.n_audio_ctx, self.dims.d_model), dtype=self.h16, device=self.device)
        _lib.check(self.lib.mw_encode(self.handle, features.data_ptr(), B, out.data_ptr(), self._stream()), "mw_encode")
        return out

    def encode_time_major(self, features_t: torch.Tensor) -> torch.Tensor:
        """features h16 [B, 2*n_audio_ctx + 2, n_mels] as emitted by mw_logmel (zero edge rows)."""
        assert features_t.is_cuda and features_t.dtype == self.h16 and features_t.is_contiguous()
        B = features_t.shape[0]
        assert features_t.shape[1:] == (2 * self.dims.n_audio_ctx + 2, self.dims.n_mels), features_t.shape
        out = torch.empty((B, self.dims.n_audio_ctx, self.dims.d_model), dtype=self.h16, device=self.device)
        _lib.check(self.lib.mw_encode_t(self.handle, features_t.data_ptr(), B, out.data_ptr(), self._stream()),
                   "mw_encode_t")
        return out

    # ---- S3
    def set_solo(self, solo: bool) -> None:
        """mw_set_solo: True when this replica is the only one decoding on its GPU (step built for latency), False when
        several replicas decode concurrently (default).  Same ids either way."""
        _lib.check(self.lib.mw_set_solo(self.handle, 1 if solo else 0), "mw_set_solo")

    def generate(self, enc: torch.Tensor, prompt: Sequence[int], tokens: SpecialTokens, *, beam_size: int = 5,
                 patience: float = 1.0, length_penalty: float = 1.0, max_length: int = 448,
                 suppress_blank: bool = True, suppress_tokens: Optional[Sequence[int]] = (-1,),
                 max_initial_timestamp_index: int = 50, num_hypotheses: int = 1,
                 forced_eot_len: int = 0, without_timestamps: Optional[bool] = None) -> List[GenerationResult]:
        """`without_timestamps` (what the caller's options say) decides whether the timestamp rules apply; when it is not
        given, the rules are off iff <|notimestamps|> appears in the prompt after <|startoftranscript|> - NOT just as its last
        token, because faster-whisper's get_prompt appends `prefix` tokens after it."""
        assert enc.is_cuda and enc.dtype == self.h16 and enc.is_contiguous()
        B = enc.shape[0]
        prompt = [int(t) for t in prompt]
        if not prompt:
            raise ValueError("prompt must not be empty")
        with_ts = timestamps_enabled(prompt, tokens, without_timestamps)
        sup = set()
        for t in suppress_tokens or []:
            if t == -1:
                sup.update(tokens.suppress_ids)
            elif t >= 0:
                sup.add(int(t))
        if with_ts:
            sup.add(tokens.no_timestamps)
        sup = np.array(sorted(i for i in sup if i < tokens.vocab), dtype=np.int32)
        sup_begin = np.array(tokens.suppress_ids_begin if suppress_blank else [], dtype=np.int32)
        n_new = max(0, min(max_length // 2, max_length - len(prompt)))
        nh = max(1, min(num_hypotheses, beam_size))
        opt = _lib.GenOptionsC(int(beam_size), float(patience), float(length_penalty), int(max_length),
                               len(sup), sup.ctypes.data_as(_lib.c_i32p), len(sup_begin),
                               sup_begin.ctypes.data_as(_lib.c_i32p), tokens.eot, tokens.timestamp_begin,
                               tokens.no_timestamps, int(with_ts), int(max_initial_timestamp_index), nh,
                               int(forced_eot_len))
        ids = np.zeros((B, nh, max(n_new, 1)), dtype=np.int32)
        lens = np.zeros((B, nh), dtype=np.int32)
        scores = np.zeros((B, nh), dtype=np.float32)
        pr = np.array(prompt, dtype=np.int32)
        _lib.check(self.lib.mw_generate(self.handle, enc.data_ptr(), B, pr.ctypes.data_as(_lib.c_i32p), len(prompt),
                                        C.byref(opt), ids.ctypes.data_as(_lib.c_i32p), lens.ctypes.data_as(_lib.c_i32p),
                                        scores.ctypes.data_as(_lib.c_f32p), self._stream()), "mw_generate")
        out = []
        for b in range(B):
            out.append(GenerationResult([ids[b, h, : lens[b, h]].tolist() for h in range(nh)],
                                        [float(scores[b, h]) for h in range(nh)]))
        return out

    def decoder_logits(self, enc: torch.Tensor, tokens_in: np.ndarray) -> torch.Tensor:
        """Teacher-forced logits f32 [B, n, vocab] for tokens_in int32 [B, n] (parity tests)."""
        assert enc.is_cuda and enc.dtype == self.h16 and enc.is_contiguous()
        tk = np.ascontiguousarray(tokens_in, dtype=np.int32)
        B, n = tk.shape
        out = torch.empty((B, n, self.dims.vocab), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.mw_decoder_logits(self.handle, enc.data_ptr(), B, tk.ctypes.data_as(_lib.c_i32p), n,
                                              out.data_ptr(), self._stream()), "mw_decoder_logits")
        return out

    def bench_kernel(self, which: int, B: int, iters: int = 64) -> float:
        """Average ms per launch of one hot decode kernel (include/mw_b200.h: mw_bench_kernel)."""
        ms = C.c_float(0.0)
        _lib.check(self.lib.mw_bench_kernel(self.handle, int(which), int(B), int(iters), C.byref(ms), self._stream()),
                   "mw_bench_kernel")
        return float(ms.value)

    def bench_step(self, B: int, parts: int, iters: int = 20) -> float:
        """Average ms of one decode step restricted to some kernel classes (mw_bench_step)."""
        ms = C.c_float(0.0)
        _lib.check(self.lib.mw_bench_step(self.handle, int(B), int(parts), int(iters), C.byref(ms), self._stream()),
                   "mw_bench_step")
        return float(ms.value)

    def detect_language(self, enc: torch.Tensor, tokens: SpecialTokens) -> np.ndarray:
        B = enc.shape[0]
        probs = np.zeros((B, tokens.n_langs), dtype=np.float32)
        _lib.check(self.lib.mw_detect_language(self.handle, enc.data_ptr(), B, tokens.sot, tokens.sot + 1,
                                               tokens.n_langs, probs.ctypes.data_as(_lib.c_f32p), self._stream()),
                   "mw_detect_language")
        return probs
