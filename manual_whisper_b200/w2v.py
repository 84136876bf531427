"""wav2vec2-CTC alignment model: dimensions, Hugging Face checkpoint key names and seeded random-init weights.

The reference aligns with ``whisperx.load_align_model(language_code, device)`` (/root/reference/transcribe.py:127-129); for
"zh" that is a Hugging Face ``Wav2Vec2ForCTC`` of the XLSR-53 family (feat_extract_norm="layer", do_stable_layer_norm=True);
for en/fr/de/es/it whisperx loads torchaudio's wav2vec2-base bundles (group-norm extractor, post-LayerNorm encoder).  The CUDA
engine implements both families (csrc/w2v.cu, variant 0 / 1).  No checkpoint exists offline, so `random_init_w2v` provides
seeded weights of either architecture for tests and benchmarks; `torchaudio_to_hf` renames a torchaudio state dict.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch


@dataclass(frozen=True)
class W2vDims:
    name: str = "w2v-large-xlsr"
    n_layers: int = 24
    d_model: int = 1024
    n_heads: int = 16
    ffn: int = 4096
    vocab: int = 32
    conv_dim: int = 512
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    pos_kernel: int = 128
    pos_groups: int = 16
    # "layer" + stable_layer_norm=True + conv bias is the XLSR-53 family; "group" + False + no conv bias is wav2vec2-base
    # (whisperx's torchaudio models for en/fr/de/es/it); the CUDA engine implements exactly these two combinations
    feat_norm: str = "layer"
    stable_layer_norm: bool = True
    conv_bias: bool = True

    def frames(self, n_samples: int) -> int:
        t = int(n_samples)
        for k, s in zip(self.conv_kernel, self.conv_stride):
            t = (t - k) // s + 1 if t >= k else 0
        return max(t, 0)



def effective_pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """The positional conv's weight [d, d/groups, taps] with its weight-norm parametrisation folded in (norm over every
    dim but the tap axis: modeling_wav2vec2.py:337-355); accepts the plain, ``parametrizations`` and ``weight_g/v`` forms."""
    p = "wav2vec2.encoder.pos_conv_embed.conv."
    if p + "weight" in sd:
        return sd[p + "weight"].float().cpu()
    if p + "parametrizations.weight.original0" in sd:
        g, v = sd[p + "parametrizations.weight.original0"], sd[p + "parametrizations.weight.original1"]
    else:
        g, v = sd[p + "weight_g"], sd[p + "weight_v"]
    g, v = g.float().cpu(), v.float().cpu()
    return v * (g / v.norm(p=2, dim=(0, 1), keepdim=True))


def torchaudio_to_hf(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """State dict of a ``torchaudio.models.Wav2Vec2Model`` (what whisperx's English / French / German / Spanish / Italian
    alignment models are: ``torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H.get_model()`` etc.) -> the Hugging Face
    ``Wav2Vec2ForCTC`` key names the rest of this package uses.  Same tensors, renamed: the two implementations share the
    architecture (tests/test_oracle_align.py pins the mapping against torchaudio's forward pass)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith("feature_extractor."):
            nk = "wav2vec2." + k
        elif k.startswith("encoder.feature_projection."):
            nk = "wav2vec2." + k[len("encoder."):]
        elif k.startswith("encoder.transformer."):
            nk = "wav2vec2.encoder." + k[len("encoder.transformer."):]
        elif k.startswith("aux."):
            nk = "lm_head." + k[len("aux."):]
        else:
            raise ValueError(f"unexpected torchaudio wav2vec2 key {k!r}")
        out[nk] = v
    return out


def is_torchaudio_state_dict(sd) -> bool:
    return any(k.startswith("encoder.transformer.") for k in sd) and not any(k.startswith("wav2vec2.") for k in sd)


# wav2vec2-base: facebook/wav2vec2-base-960h = torchaudio WAV2VEC2_ASR_BASE_960H, the family whisperx aligns en/fr/de/es/it with
BASE_LANGUAGES = ("en", "fr", "de", "es", "it")


def base_dims(vocab: int = 32, name: str = "w2v-base") -> "W2vDims":
    return W2vDims(name=name, n_layers=12, d_model=768, n_heads=12, ffn=3072, vocab=vocab, conv_dim=512, pos_kernel=128,
                   pos_groups=16, feat_norm="group", stable_layer_norm=False, conv_bias=False)


def random_init_w2v(dims: W2vDims, seed: int = 0, std: float = 0.05) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights with the Hugging Face ``Wav2Vec2ForCTC`` key names; every matrix is representable in the engine's 16-bit storage (weights.round_shared) so
    the CUDA engine and this oracle hold identical values.  The positional conv is stored as its effective weight."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        from .weights import round_shared
        return round_shared(torch.randn(*shape, generator=g) * s)

    sd: Dict[str, torch.Tensor] = {}
    c_in = 1
    for i, k in enumerate(dims.conv_kernel):
        p = f"wav2vec2.feature_extractor.conv_layers.{i}"
        sd[p + ".conv.weight"] = rnd(dims.conv_dim, c_in, k, s=(2.0 / (c_in * k)) ** 0.5)
        if dims.conv_bias:
            sd[p + ".conv.bias"] = rnd(dims.conv_dim, s=0.02)
        if dims.feat_norm == "layer" or i == 0:
            sd[p + ".layer_norm.weight"] = 1.0 + rnd(dims.conv_dim, s=0.05)
            sd[p + ".layer_norm.bias"] = rnd(dims.conv_dim, s=0.05)
        c_in = dims.conv_dim
    sd["wav2vec2.feature_projection.layer_norm.weight"] = 1.0 + rnd(dims.conv_dim, s=0.05)
    sd["wav2vec2.feature_projection.layer_norm.bias"] = rnd(dims.conv_dim, s=0.05)
    sd["wav2vec2.feature_projection.projection.weight"] = rnd(dims.d_model, dims.conv_dim)
    sd["wav2vec2.feature_projection.projection.bias"] = rnd(dims.d_model, s=0.02)
    gs = dims.d_model // dims.pos_groups
    sd["wav2vec2.encoder.pos_conv_embed.conv.weight"] = rnd(dims.d_model, gs, dims.pos_kernel, s=(1.0 / (gs * dims.pos_kernel)) ** 0.5)
    sd["wav2vec2.encoder.pos_conv_embed.conv.bias"] = rnd(dims.d_model, s=0.02)
    for l in range(dims.n_layers):
        p = f"wav2vec2.encoder.layers.{l}"
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{p}.attention.{nm}.weight"] = rnd(dims.d_model, dims.d_model)
            sd[f"{p}.attention.{nm}.bias"] = rnd(dims.d_model, s=0.02)
        for nm in ("layer_norm", "final_layer_norm"):
            sd[f"{p}.{nm}.weight"] = 1.0 + rnd(dims.d_model, s=0.05)
            sd[f"{p}.{nm}.bias"] = rnd(dims.d_model, s=0.05)
        sd[p + ".feed_forward.intermediate_dense.weight"] = rnd(dims.ffn, dims.d_model)
        sd[p + ".feed_forward.intermediate_dense.bias"] = rnd(dims.ffn, s=0.02)
        sd[p + ".feed_forward.output_dense.weight"] = rnd(dims.d_model, dims.ffn, s=std / 2)
        sd[p + ".feed_forward.output_dense.bias"] = rnd(dims.d_model, s=0.02)
    sd["wav2vec2.encoder.layer_norm.weight"] = 1.0 + rnd(dims.d_model, s=0.05)
    sd["wav2vec2.encoder.layer_norm.bias"] = rnd(dims.d_model, s=0.05)
    sd["lm_head.weight"] = rnd(dims.vocab, dims.d_model, s=0.2)
    sd["lm_head.bias"] = rnd(dims.vocab, s=0.02)
    return sd
