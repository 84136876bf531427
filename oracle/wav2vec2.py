"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the wav2vec2-CTC acoustic model that
``whisperx.align`` runs on every segment (/root/reference/transcribe.py:127-135: ``whisperx.load_align_model`` +
``whisperx.align``; SURVEY.md §8f row 3, "forced alignment").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

The reference's alignment model for its language ("zh", transcribe.py:110) is a Hugging Face ``Wav2Vec2ForCTC`` of the
XLSR-53 family (whisperx DEFAULT_ALIGN_MODELS_HF: jonatasgrosman/wav2vec2-large-xlsr-53-chinese-zh-cn [UPSTREAM-MEMORY]):
``feat_extract_norm="layer"``, ``do_stable_layer_norm=True``, conv bias on.  whisperx feeds the RAW waveform slice (no
processor normalisation) and takes ``log_softmax(model(x).logits)``.  The checkpoint cannot be downloaded offline, so
weights are random-init of that architecture (HF key names).  PARITY PINNED BY: transformers'
modeling_wav2vec2.py, importable here, on the same weights (tests/test_oracle_align.py).

``emulate=True`` rounds activations where the CUDA engine stores 16-bit values (GEMM operands; fp16, or bf16 for the
MW_STORAGE_BF16=1 build - oracle/model.py: engine_rounding), keeping fp32 accumulation.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


from .model import _r, engine_rounding


def pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Effective weight of the weight-normalised positional conv (norm over every dim but the tap axis, dim=2)."""
    p = "wav2vec2.encoder.pos_conv_embed.conv."
    if p + "weight" in sd:
        return sd[p + "weight"].float()
    if p + "parametrizations.weight.original0" in sd:
        g, v = sd[p + "parametrizations.weight.original0"].float(), sd[p + "parametrizations.weight.original1"].float()
    else:
        g, v = sd[p + "weight_g"].float(), sd[p + "weight_v"].float()
    return v * (g / v.norm(p=2, dim=(0, 1), keepdim=True))


class OracleWav2Vec2:
    """logits = lm_head(encoder(feature_projection(conv_stack(waveform)))), one waveform at a time like whisperx.align."""

    def __init__(self, dims, sd: Dict[str, torch.Tensor], emulate=False):
        self.dims = dims
        self.sd = {k: v.to(torch.float32) for k, v in sd.items()}
        self.emu = engine_rounding() if emulate is True else emulate
        self.w_pos = pos_conv_weight(self.sd)
        if self.emu:
            self.w_pos = _r(self.w_pos, self.emu)

    def _ln(self, x, prefix, eps=1e-5):
        return F.layer_norm(x, (x.shape[-1],), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"], eps)

    def _lin(self, x, prefix):
        return F.linear(_r(x, self.emu), self.sd[prefix + ".weight"], self.sd[prefix + ".bias"])

    def features(self, wave: torch.Tensor) -> torch.Tensor:
        """[n_samples] -> [T, conv_dim]: 7 x (conv1d, LayerNorm over channels, GELU)  (modeling_wav2vec2.py:275-300)."""
        d = self.dims
        group = getattr(d, "feat_norm", "layer") == "group"
        h = wave.view(1, 1, -1).float()
        for i, (k, s) in enumerate(zip(d.conv_kernel, d.conv_stride)):
            p = f"wav2vec2.feature_extractor.conv_layers.{i}"
            if i > 0:
                h = _r(h, self.emu)          # the engine stores every conv layer's GELU output as bf16
            h = F.conv1d(h, self.sd[p + ".conv.weight"], self.sd.get(p + ".conv.bias"), stride=s)
            if not group:
                h = self._ln(h.transpose(1, 2), p + ".layer_norm").transpose(1, 2)
            elif i == 0:     # wav2vec2-base: GroupNorm(C groups) = per-channel statistics over time, first layer only (:302-323)
                h = F.group_norm(h, h.shape[1], self.sd[p + ".layer_norm.weight"], self.sd[p + ".layer_norm.bias"], 1e-5)
            h = F.gelu(h)
        return h[0].transpose(0, 1)

    def hidden(self, wave: torch.Tensor) -> torch.Tensor:
        d = self.dims
        feat = self.features(wave)                                               # [T, 512]
        x = self._lin(self._ln(feat, "wav2vec2.feature_projection.layer_norm"), "wav2vec2.feature_projection.projection")
        # positional conv embedding (modeling_wav2vec2.py:326-368): grouped conv k=128 pad 64, drop last, GELU, added
        xp = _r(x, self.emu).transpose(0, 1).unsqueeze(0)
        pos = F.conv1d(xp, self.w_pos, self.sd["wav2vec2.encoder.pos_conv_embed.conv.bias"], padding=d.pos_kernel // 2,
                       groups=d.pos_groups)
        if d.pos_kernel % 2 == 0:
            pos = pos[:, :, :-1]
        x = x + F.gelu(pos)[0].transpose(0, 1)
        T = x.shape[0]
        H, dh = d.n_heads, d.d_model // d.n_heads
        stable = getattr(d, "stable_layer_norm", True)
        if not stable:                                                           # post-LN encoder (:658-727): LN right after pos
            x = self._ln(x, "wav2vec2.encoder.layer_norm")
        for l in range(d.n_layers):                                             # stable-layer-norm (pre-LN) layers, :612-655
            p = f"wav2vec2.encoder.layers.{l}"
            y = self._ln(x, p + ".layer_norm") if stable else x
            q = _r(self._lin(y, p + ".attention.q_proj"), self.emu).view(T, H, dh).transpose(0, 1)
            k = _r(self._lin(y, p + ".attention.k_proj"), self.emu).view(T, H, dh).transpose(0, 1)
            v = _r(self._lin(y, p + ".attention.v_proj"), self.emu).view(T, H, dh).transpose(0, 1)
            s = torch.matmul(q, k.transpose(1, 2)) * dh ** -0.5
            if self.emu:   # the flash kernel rounds the un-normalised exp() to bf16 before P.V, sums unrounded
                m = s.max(dim=-1, keepdim=True).values
                e = torch.exp(s - m)
                o = torch.matmul(_r(e, self.emu), v) / e.sum(dim=-1, keepdim=True)
            else:
                o = torch.matmul(torch.softmax(s, dim=-1), v)
            o = o.transpose(0, 1).reshape(T, d.d_model)
            x = x + self._lin(o, p + ".attention.out_proj")
            if stable:
                y = self._ln(x, p + ".final_layer_norm")
                y = F.gelu(self._lin(y, p + ".feed_forward.intermediate_dense"))
                x = x + self._lin(y, p + ".feed_forward.output_dense")
            else:                                                                # :592-609: LN after each residual add
                x = self._ln(x, p + ".layer_norm")
                y = F.gelu(self._lin(x, p + ".feed_forward.intermediate_dense"))
                x = self._ln(x + self._lin(y, p + ".feed_forward.output_dense"), p + ".final_layer_norm")
        return self._ln(x, "wav2vec2.encoder.layer_norm") if stable else x

    def logits(self, wave: torch.Tensor) -> torch.Tensor:
        return self._lin(self.hidden(wave), "lm_head")

    def emissions(self, wave: torch.Tensor) -> torch.Tensor:
        """whisperx.align: ``torch.log_softmax(model(waveform_segment).logits, dim=-1)[0]`` -> [T, vocab]."""
        return torch.log_softmax(self.logits(wave), dim=-1)
