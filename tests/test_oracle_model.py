"""Pins the encoder/decoder oracle against the independent Hugging Face Whisper on the same weights."""
import numpy as np
import pytest
import torch

from manual_whisper_b200.config import model_dims, custom_dims, special_tokens
from manual_whisper_b200.weights import random_init, to_hf, sinusoids
from oracle.model import OracleWhisper


def hf_model(dims, sd):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    cfg = WhisperConfig(vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model,
                        encoder_layers=dims.enc_layers, decoder_layers=dims.dec_layers,
                        encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
                        encoder_ffn_dim=dims.ffn, decoder_ffn_dim=dims.ffn, max_source_positions=dims.n_audio_ctx,
                        max_target_positions=dims.n_text_ctx)
    m = WhisperForConditionalGeneration(cfg).eval()
    res = m.load_state_dict(to_hf(sd), strict=False)
    assert not res.unexpected_keys and not res.missing_keys
    return m


@pytest.fixture(scope="module")
def tiny():
    dims = model_dims("tiny")
    sd = random_init(dims, seed=1234, scheme="lively")
    return dims, sd, OracleWhisper(dims, sd), hf_model(dims, sd)


def test_state_dict_covers_every_hf_parameter(tiny):
    dims, sd, _, hf = tiny
    assert set(to_hf(sd)) == set(hf.state_dict())
    for v in sd.values():
        if v.dtype == torch.float32 and v.numel() > 10 and v is not sd["model.encoder.embed_positions.weight"]:
            assert torch.equal(v, v.half().float())           # exactly representable in the engine's fp16 storage


def test_sinusoids_match_hf(tiny):
    dims, sd, _, hf = tiny
    assert torch.allclose(sinusoids(1500, dims.d_model), hf.model.encoder.embed_positions.weight, atol=1e-6)


def test_encoder_matches_hf(tiny):
    dims, sd, orc, hf = tiny
    g = torch.Generator().manual_seed(0)
    mel = (torch.randn(2, dims.n_mels, 3000, generator=g) * 0.5).clamp(-1.5, 1.5)
    with torch.no_grad():
        a = orc.encode(mel)
        b = hf.model.encoder(mel).last_hidden_state
    assert a.shape == (2, 1500, dims.d_model)
    assert (a - b).abs().max().item() < 1e-4


def test_decoder_logits_match_hf_and_stepwise_equals_batched(tiny):
    dims, sd, orc, hf = tiny
    g = torch.Generator().manual_seed(1)
    mel = (torch.randn(1, dims.n_mels, 3000, generator=g) * 0.5).clamp(-1.5, 1.5)
    toks = torch.randint(0, dims.vocab, (1, 9), generator=g)
    with torch.no_grad():
        enc = orc.encode(mel)
        cross = orc.cross_kv(enc)
        batched = orc.decode(toks, 0, cross, orc.new_cache())
        ref = hf(input_features=mel, decoder_input_ids=toks).logits
        cache = orc.new_cache()
        step = torch.cat([orc.decode(toks[:, i:i + 1], i, cross, cache) for i in range(toks.shape[1])], 1)
    scale = ref.abs().max().item()
    assert (batched - ref).abs().max().item() < 1e-4 * max(1.0, scale)
    assert (step - batched).abs().max().item() < 1e-4 * max(1.0, scale)


def test_bf16_emulation_stays_within_bf16_tolerance(tiny):
    dims, sd, orc, _ = tiny
    g = torch.Generator().manual_seed(2)
    mel = (torch.randn(1, dims.n_mels, 3000, generator=g) * 0.5).clamp(-1.5, 1.5)
    emu = OracleWhisper(dims, sd, emulate=True)
    with torch.no_grad():
        a, b = orc.encode(mel), emu.encode(mel)
    rel = ((a - b).norm() / a.norm()).item()
    assert 0 < rel < 2e-2


def test_small_model_matches_committed_golden(golden_small):
    from manual_whisper_b200.config import scaled_tokens
    dims = custom_dims("golden-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=100, n_text_ctx=32)
    sd = random_init(dims, seed=11, scheme="lively")
    orc = OracleWhisper(dims, sd)
    mel = torch.from_numpy(golden_small["mel"])
    with torch.no_grad():
        enc, layers = orc.encode(mel, return_layers=True)
        logits = orc.decode(torch.from_numpy(golden_small["tokens"]), 0, orc.cross_kv(enc), orc.new_cache())
    np.testing.assert_allclose(enc.reshape(-1)[::17].numpy(), golden_small["enc_sub"], atol=2e-5)
    np.testing.assert_allclose([l.double().sum().item() for l in layers], golden_small["layer_sums"], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(logits.reshape(-1)[::97].numpy(), golden_small["logits_sub"], atol=5e-5)
