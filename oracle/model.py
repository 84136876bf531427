"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the Whisper engine.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Restates what ``ctranslate2.models.Whisper.encode`` / ``.generate`` compute behind
``whisperx ... model.transcribe`` (/root/reference/transcribe.py:107-113,123).  CTranslate2 >= 4.5 /
faster-whisper >= 1.1.1 / whisperx 3.7.6 are un-vendored third-party dependencies and are absent from
this container, so this follows SURVEY.md Appendix A.7-A.8 and is pinned against the independent
in-container implementation transformers/models/whisper/modeling_whisper.py on the same weights
(tests/test_oracle_model.py).  PARITY PINNED BY: HF twin; the reference itself pins nothing.

``emulate=True`` rounds activations at exactly the points where the CUDA engine stores 16-bit values
(DESIGN.md "rounding points") - to fp16, the engine's storage type (or bf16 when MW_STORAGE_BF16=1 selects the A/B
build; "fp16" / "bf16" force one) - keeping fp32 accumulation; weights are exactly representable on both
sides already.  With it off this is the plain fp32 ground truth.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def engine_rounding() -> str:
    """Storage type of the library build under test: fp16 unless MW_STORAGE_BF16=1 built the bf16 variant."""
    import os
    return "bf16" if os.environ.get("MW_STORAGE_BF16") == "1" else "fp16"


def _r(x: torch.Tensor, on) -> torch.Tensor:
    """Round to the engine's storage type at a rounding point.  `on` is False (fp32 ground truth), True / "bf16"
    (what the engine stores) or "fp16" (analysis only: what an fp16-storing engine would do, scripts/parity_probe.py)."""
    if not on:
        return x
    return x.to(torch.float16 if on == "fp16" else torch.bfloat16).to(torch.float32)


class OracleWhisper:
    def __init__(self, dims, sd: Dict[str, torch.Tensor], emulate=False, int8: bool = False):
        """int8=True: every Linear (and the tied output projection) runs as a dynamically quantised int8 GEMM - int8 weights
        with per-output-row scales, activations quantised per call, int32 accumulation, fp32 out; convolutions, LayerNorm,
        softmax, GELU and residuals stay fp32.  That is the arithmetic of CTranslate2's compute_type="int8" on CPU (SURVEY.md
        A.9), the mode the reference ships (/root/reference/transcribe.py:30-32); used by bench.py's CPU legs only."""
        self.int8 = int8
        self._packed = {}
        self.dims = dims
        self.sd = {k: v.to(torch.float32) for k, v in sd.items()}
        self.emu = engine_rounding() if emulate is True else emulate
        self.scale = dims.d_head ** -0.5

    # ------------------------------------------------------------------ helpers
    def _ln(self, x, prefix):
        return F.layer_norm(x, (x.shape[-1],), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], 1e-5)

    def _qlinear(self, x, key, w, b):
        packed = self._packed.get(key)
        if packed is None:
            scales = (w.abs().amax(dim=1).clamp_min(1e-12) / 127.0).to(torch.float64)
            qw = torch.quantize_per_channel(w, scales, torch.zeros(w.shape[0], dtype=torch.int64), 0, torch.qint8)
            packed = self._packed[key] = torch.ops.quantized.linear_prepack(qw, b)
        shape = x.shape
        y = torch.ops.quantized.linear_dynamic(x.reshape(-1, shape[-1]).contiguous(), packed, reduce_range=False)
        return y.reshape(*shape[:-1], w.shape[0])

    def _lin(self, x, prefix, bias=True):
        w, b = self.sd[prefix + ".weight"], (self.sd.get(prefix + ".bias") if bias else None)
        if self.int8:
            return self._qlinear(x, prefix, w, b)
        return F.linear(x, w, b)

    def _heads(self, x):  # [B,T,d] -> [B,H,T,64]
        B, T, _ = x.shape
        return x.view(B, T, self.dims.n_heads, self.dims.d_head).transpose(1, 2)

    def _attn(self, q, k, v, mask=None, round_p=False):
        """softmax((q*scale) k^T) v, SURVEY.md A.7. q,k,v: [B,H,T,64]."""
        s = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        if mask is not None:
            s = s + mask
        p = torch.softmax(s, dim=-1)
        if round_p:
            # the encoder's flash kernel rounds the un-normalised exp() to bf16 before P.V
            m = s.max(dim=-1, keepdim=True).values
            e = _r(torch.exp(s - m), round_p)
            o = torch.matmul(e, v) / torch.exp(s - m).sum(dim=-1, keepdim=True)
        else:
            o = torch.matmul(p, v)
        B, H, T, D = o.shape
        return o.transpose(1, 2).reshape(B, T, H * D)

    # ------------------------------------------------------------------ encoder
    def encode(self, mel: torch.Tensor, return_layers: bool = False):
        """mel f32 [B, n_mels, 3000] -> [B, 1500, d]  (SURVEY.md A.7)."""
        emu = self.emu
        sd = self.sd
        x = _r(mel.to(torch.float32), emu)
        x = F.gelu(F.conv1d(x, sd["model.encoder.conv1.weight"], sd["model.encoder.conv1.bias"], padding=1))
        x = _r(x, emu)
        x = F.gelu(F.conv1d(x, sd["model.encoder.conv2.weight"], sd["model.encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1) + sd["model.encoder.embed_positions.weight"][: x.shape[2]]
        layers = []
        for i in range(self.dims.enc_layers):
            p = f"model.encoder.layers.{i}."
            h = _r(self._ln(x, p + "self_attn_layer_norm"), emu)
            q = _r(self._lin(h, p + "self_attn.q_proj"), emu)
            k = _r(self._lin(h, p + "self_attn.k_proj", bias=False), emu)
            v = _r(self._lin(h, p + "self_attn.v_proj"), emu)
            a = _r(self._attn(self._heads(q), self._heads(k), self._heads(v), round_p=emu), emu)
            x = x + self._lin(a, p + "self_attn.out_proj")
            h = _r(self._ln(x, p + "final_layer_norm"), emu)
            h = _r(F.gelu(self._lin(h, p + "fc1")), emu)
            x = x + self._lin(h, p + "fc2")
            if return_layers:
                layers.append(x.clone())
        out = _r(self._ln(x, "model.encoder.layer_norm"), emu)
        return (out, layers) if return_layers else out

    # ------------------------------------------------------------------ decoder
    def cross_kv(self, enc: torch.Tensor) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Per decoder layer K/V of the encoder output, projected once per chunk."""
        out = []
        for i in range(self.dims.dec_layers):
            p = f"model.decoder.layers.{i}.encoder_attn."
            k = _r(self._lin(enc, p + "k_proj", bias=False), self.emu)
            v = _r(self._lin(enc, p + "v_proj"), self.emu)
            out.append((self._heads(k), self._heads(v)))
        return out

    def new_cache(self):
        return [None] * self.dims.dec_layers

    def decode(self, tokens: torch.Tensor, pos0: int, cross, cache, cross_index: Optional[torch.Tensor] = None):
        """Run `tokens` [R, n] occupying positions pos0..pos0+n-1 through the decoder, appending to `cache`
        (list per layer of (K,V) [R,H,t,64]).  `cross_index` [R] maps a row to its chunk (beam search shares
        one chunk's cross K/V among its beams).  Returns fp32 logits [R, n, V]."""
        emu = self.emu
        sd = self.sd
        R, n = tokens.shape
        x = sd["model.decoder.embed_tokens.weight"][tokens] + sd["model.decoder.embed_positions.weight"][pos0: pos0 + n]
        mask = None
        if n > 1:
            mask = torch.full((n, pos0 + n), float("-inf")).triu(pos0 + 1)
        grouped = False
        if cross_index is not None and cross[0][0].shape[0] > 0 and R % cross[0][0].shape[0] == 0:
            Bc = cross[0][0].shape[0]
            grouped = bool(torch.equal(cross_index, torch.arange(Bc).repeat_interleave(R // Bc)))
        for i in range(self.dims.dec_layers):
            p = f"model.decoder.layers.{i}."
            h = _r(self._ln(x, p + "self_attn_layer_norm"), emu)
            q = self._heads(_r(self._lin(h, p + "self_attn.q_proj"), emu))
            k = self._heads(_r(self._lin(h, p + "self_attn.k_proj", bias=False), emu))
            v = self._heads(_r(self._lin(h, p + "self_attn.v_proj"), emu))
            if cache[i] is not None:
                k = torch.cat([cache[i][0], k], dim=2)
                v = torch.cat([cache[i][1], v], dim=2)
            cache[i] = (k, v)
            a = _r(self._attn(q, k, v, mask), emu)
            x = x + self._lin(a, p + "self_attn.out_proj")
            h = _r(self._ln(x, p + "encoder_attn_layer_norm"), emu)
            q = self._heads(_r(self._lin(h, p + "encoder_attn.q_proj"), emu))
            ck, cv = cross[i]
            if cross_index is not None and grouped:
                # beams of one chunk are consecutive rows: attend as [B, H, k*n, 64] against the chunk's K/V instead of
                # copying K/V per beam (same dot products; it only spares the oracle gigabytes of copies per step)
                Bc, kb = ck.shape[0], R // ck.shape[0]
                qg = q.reshape(Bc, kb, q.shape[1], n, q.shape[3]).permute(0, 2, 1, 3, 4).reshape(Bc, q.shape[1], kb * n, q.shape[3])
                a = self._attn(qg, ck, cv).reshape(Bc, kb, n, -1).reshape(R, n, -1)
                a = _r(a, emu)
            else:
                if cross_index is not None:
                    ck, cv = ck[cross_index], cv[cross_index]
                a = _r(self._attn(q, ck, cv), emu)
            x = x + self._lin(a, p + "encoder_attn.out_proj")
            h = _r(self._ln(x, p + "final_layer_norm"), emu)
            h = _r(F.gelu(self._lin(h, p + "fc1")), emu)
            x = x + self._lin(h, p + "fc2")
        h = _r(self._ln(x, "model.decoder.layer_norm"), emu)
        if self.int8:
            return self._qlinear(h, "proj_out", sd["model.decoder.embed_tokens.weight"], None)
        return F.linear(h, sd["model.decoder.embed_tokens.weight"])

    @staticmethod
    def reorder_cache(cache, parent: torch.Tensor):
        return [None if c is None else (c[0][parent], c[1][parent]) for c in cache]
