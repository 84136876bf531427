"""The forced-alignment oracle (oracle/wav2vec2.py, oracle/align.py): the acoustic model is pinned to transformers'
Wav2Vec2ForCTC on the same weights; the CTC trellis/backtrack to brute-force enumeration (no second implementation of
that objective exists offline)."""
import itertools

import numpy as np
import pytest
import torch

from manual_whisper_b200.w2v import W2vDims, random_init_w2v
from oracle.wav2vec2 import OracleWav2Vec2, pos_conv_weight
from oracle import align as OA

SMALL = W2vDims(name="w2v-test", n_layers=2, d_model=128, n_heads=2, ffn=256, vocab=40, conv_dim=64, pos_kernel=16, pos_groups=4)


def _hf_twin(dims, sd):
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    cfg = Wav2Vec2Config(vocab_size=dims.vocab, hidden_size=dims.d_model, num_hidden_layers=dims.n_layers,
                         num_attention_heads=dims.n_heads, intermediate_size=dims.ffn, feat_extract_norm=dims.feat_norm,
                         do_stable_layer_norm=dims.stable_layer_norm, conv_bias=dims.conv_bias, conv_dim=[dims.conv_dim] * 7,
                         conv_kernel=list(dims.conv_kernel), conv_stride=list(dims.conv_stride),
                         num_conv_pos_embeddings=dims.pos_kernel, num_conv_pos_embedding_groups=dims.pos_groups,
                         hidden_act="gelu", feat_extract_activation="gelu", layer_norm_eps=1e-5)
    model = Wav2Vec2ForCTC(cfg).eval()
    own = model.state_dict()
    new = {}
    w = sd["wav2vec2.encoder.pos_conv_embed.conv.weight"]
    for k in own:
        if k.endswith("parametrizations.weight.original1") or k.endswith("weight_v"):
            new[k] = w
        elif k.endswith("parametrizations.weight.original0") or k.endswith("weight_g"):
            new[k] = w.norm(p=2, dim=(0, 1), keepdim=True)        # g = |v|  ->  effective weight = v
        elif k in sd:
            new[k] = sd[k]
        elif k == "wav2vec2.masked_spec_embed":
            new[k] = own[k]
        else:
            raise KeyError(k)
    model.load_state_dict(new)
    return model


def test_frames_formula():
    d = W2vDims()
    assert d.frames(480000) == 1499 and d.frames(16000) == 49 and d.frames(400) == 1 and d.frames(399) == 0


def test_wav2vec2_oracle_matches_hf():
    sd = random_init_w2v(SMALL, seed=3)
    hf = _hf_twin(SMALL, sd)
    ora = OracleWav2Vec2(SMALL, sd)
    g = torch.Generator().manual_seed(0)
    for n in (400, 3217, 16000):
        wave = torch.randn(n, generator=g) * 0.1
        with torch.no_grad():
            want = hf(wave[None]).logits[0]
            got = ora.logits(wave)
        assert got.shape == want.shape == (SMALL.frames(n), SMALL.vocab)
        assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())
        with torch.no_grad():
            assert torch.allclose(ora.emissions(wave), torch.log_softmax(want, -1), atol=3e-4)


def test_group_norm_base_variant_oracle_matches_hf():
    """wav2vec2-base layout (GroupNorm on conv 0, no conv bias, post-LN encoder), the oracle of tests/test_gpu_align.py's
    base-variant test, pinned to transformers' Wav2Vec2ForCTC."""
    from dataclasses import replace
    base = replace(SMALL, name="w2v-base-test", feat_norm="group", stable_layer_norm=False, conv_bias=False)
    sd = random_init_w2v(base, seed=4)
    hf = _hf_twin(base, sd)
    ora = OracleWav2Vec2(base, sd)
    wave = torch.randn(5000, generator=torch.Generator().manual_seed(1)) * 0.1
    with torch.no_grad():
        want, got = hf(wave[None]).logits[0], ora.logits(wave)
    assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())


def test_torchaudio_state_dict_maps_onto_the_oracle():
    """whisperx aligns en/fr/de/es/it with torchaudio bundles (WAV2VEC2_ASR_BASE_960H ...): a torchaudio Wav2Vec2Model state
    dict renamed by torchaudio_to_hf gives torchaudio's own logits through the oracle (group-norm / post-LN family)."""
    import torchaudio
    from dataclasses import replace
    from manual_whisper_b200.w2v import torchaudio_to_hf, is_torchaudio_state_dict
    torch.manual_seed(5)
    m = torchaudio.models.wav2vec2_model(
        extractor_mode="group_norm", extractor_conv_layer_config=[(64, 10, 5)] + [(64, 3, 2)] * 4 + [(64, 2, 2)] * 2,
        extractor_conv_bias=False, encoder_embed_dim=128, encoder_projection_dropout=0.0, encoder_pos_conv_kernel=16,
        encoder_pos_conv_groups=4, encoder_num_layers=2, encoder_num_heads=2, encoder_attention_dropout=0.0,
        encoder_ff_interm_features=256, encoder_ff_interm_dropout=0.0, encoder_dropout=0.0, encoder_layer_norm_first=False,
        encoder_layer_drop=0.0, aux_num_out=40).eval()
    sd = m.state_dict()
    assert is_torchaudio_state_dict(sd)
    hf_sd = torchaudio_to_hf(sd)
    assert not is_torchaudio_state_dict(hf_sd) and "lm_head.weight" in hf_sd
    dims = replace(SMALL, name="ta", feat_norm="group", stable_layer_norm=False, conv_bias=False)
    wave = torch.randn(6000, generator=torch.Generator().manual_seed(2)) * 0.1
    with torch.no_grad():
        want = m(wave[None])[0][0]
        got = OracleWav2Vec2(dims, hf_sd).logits(wave)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 2e-4 * max(1.0, want.abs().max().item())


def test_pos_conv_weight_norm_forms():
    sd = random_init_w2v(SMALL, seed=1)
    w = sd["wav2vec2.encoder.pos_conv_embed.conv.weight"]
    p = "wav2vec2.encoder.pos_conv_embed.conv."
    g = torch.rand(1, 1, SMALL.pos_kernel) + 0.5
    want = w * (g / w.norm(p=2, dim=(0, 1), keepdim=True))
    for names in (("parametrizations.weight.original0", "parametrizations.weight.original1"), ("weight_g", "weight_v")):
        assert torch.allclose(pos_conv_weight({p + names[0]: g, p + names[1]: w}), want)


@pytest.mark.parametrize("seed", range(6))
def test_trellis_and_backtrack_against_enumeration(seed):
    rng = np.random.default_rng(seed)
    T, V, N = int(rng.integers(5, 10)), 6, int(rng.integers(2, 5))
    em = np.log(rng.dirichlet(np.ones(V), size=T)).astype(np.float32)
    tokens = [int(x) for x in rng.integers(1, V, size=N)]
    if seed % 2:
        tokens[-1] = OA.WILDCARD
    tr = OA.get_trellis(em, tokens, blank=0)
    # objective of a path: token j is entered by emitting tokens[j] at its change frame; every other frame before the last
    # scores blank - except that column 0 starts counting blanks at frame 1 (the published cumsum starts at emission[1])
    best = -np.inf
    for cs in itertools.combinations(range(T - 1), N - 1):
        s = 0.0
        for t in range(T - 1):
            if t in cs:
                s += OA._token_emission(em[t], tokens[cs.index(t) + 1], 0)
            elif t < cs[0]:
                s += em[t + 1, 0]                                  # column 0: trellis[t, 0] = sum(em[1..t, blank])
            else:
                s += em[t, 0]
        if s > best:
            best, arg = s, cs
    assert abs(tr[T - 1, N - 1] - best) < 1e-4
    path = OA.backtrack(tr, em, tokens, 0)
    assert path is not None and len(path) == T
    ft = OA.frame_tokens(path, T)
    assert ft[0] == 0 and ft[-1] == N - 1 and np.all(np.diff(ft) >= 0) and np.all(np.diff(ft) <= 1)
    changes = tuple(int(t) for t in range(T - 1) if ft[t + 1] != ft[t])
    assert changes == arg
    segs = OA.merge_repeats(path, "abcdefgh"[:N])
    assert [s.label for s in segs] == list("abcdefgh"[:N]) and segs[0].start == 0 and segs[-1].end == T
    assert all(a.end == b.start for a, b in zip(segs, segs[1:]))


def test_backtrack_reports_unalignable():
    em = np.log(np.full((3, 4), 0.25, dtype=np.float32))
    tokens = [1, 2, 3, 1, 2]
    assert OA.backtrack(OA.get_trellis(em, tokens), em, tokens) is None


def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows_golden.npz"))


def test_oracles_against_committed_golden_vectors():
    """Fixtures written by scripts/make_golden_next_rows.py from transformers / torchaudio (and the oracle DP itself)."""
    from oracle.resample import resample
    g = _golden()
    sd = random_init_w2v(SMALL, seed=3)
    with torch.no_grad():
        got = OracleWav2Vec2(SMALL, sd).logits(torch.from_numpy(g["w2v_wave"]))
    assert np.abs(got.numpy() - g["w2v_logits"]).max() < 2e-4 * max(1.0, np.abs(g["w2v_logits"]).max())
    for rate in (44100, 48000):
        assert np.abs(resample(g[f"resample_in_{rate}"], rate, 16000) - g[f"resample_out_{rate}"]).max() < 2e-5
    em, toks = g["ctc_emission"], [int(t) for t in g["ctc_tokens"]]
    tr = OA.get_trellis(em, toks, 0)
    path = OA.backtrack(tr, em, toks, 0)
    assert abs(float(tr[-1, -1]) - float(g["ctc_final_score"])) < 1e-5
    assert np.array_equal(OA.frame_tokens(path, em.shape[0]), g["ctc_frame_tokens"])
    assert np.allclose([p.score for p in path], g["ctc_frame_scores"], rtol=1e-6)
