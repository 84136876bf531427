"""Forced alignment on the GPU through the C ABI: wav2vec2-CTC emissions against the oracle pinned to
transformers' Wav2Vec2ForCTC, the CTC trellis/backtrack kernel against the oracle DP, and the whisperx.align mirror."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from manual_whisper_b200.w2v import W2vDims, random_init_w2v

SMALL = W2vDims(name="w2v-test", n_layers=2, d_model=128, n_heads=2, ffn=256, vocab=40, conv_dim=128, pos_kernel=16, pos_groups=2)


def _audio(n, seed):
    g = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.1 * g.standard_normal(n) + 0.2 * np.sin(2 * np.pi * 220 * t) * (1 + np.sin(2 * np.pi * 3 * t))).astype(np.float32)


@pytest.fixture(scope="module")
def small():
    from manual_whisper_b200.alignment import AlignEngine
    sd = random_init_w2v(SMALL, seed=7)
    eng = AlignEngine(SMALL, sd, device_index=0, max_batch=4, max_samples=40000)
    return sd, eng


def _oracle_emissions(dims, sd, wave, emulate):
    from oracle.wav2vec2 import OracleWav2Vec2
    with torch.no_grad():
        w = torch.from_numpy(wave)
        if len(w) < 400:
            w = torch.nn.functional.pad(w, (0, 400 - len(w)))
        return OracleWav2Vec2(dims, sd, emulate=emulate).emissions(w)


def test_emissions_match_oracle_on_ragged_batch(small):
    sd, eng = small
    audio = _audio(70000, 0)
    offs = np.array([0, 16000, 30000, 69000], dtype=np.int64)
    lens = np.array([16000, 5003, 40000, 250], dtype=np.int32)              # the last one is shorter than 400 samples
    em, frames = eng.emissions(torch.from_numpy(audio).cuda(), offs, lens)
    assert list(frames) == [SMALL.frames(16000), SMALL.frames(5003), SMALL.frames(40000), 1]
    assert em.shape == (4, int(frames.max()), SMALL.vocab)
    em = em.cpu()
    for c in range(4):
        T = int(frames[c])
        wave = audio[offs[c]: offs[c] + lens[c]]
        ref = _oracle_emissions(SMALL, sd, wave, emulate=True)
        f32 = _oracle_emissions(SMALL, sd, wave, emulate=False)
        assert ref.shape == (T, SMALL.vocab)
        got = em[c, :T]
        assert torch.allclose(got.exp().sum(-1), torch.ones(T), atol=1e-4)          # rows are log-probabilities
        spread = (f32.max() - f32.min()).item()
        assert (got - ref).abs().max().item() <= 0.02 * spread, (c, (got - ref).abs().max().item(), spread)
        assert (got - f32).abs().max().item() <= 0.05 * spread
        assert torch.all(em[c, T:] == 0)                                             # rows beyond the window stay untouched


BASE_SMALL = None


def _base_dims(**kw):
    from dataclasses import replace
    return replace(W2vDims(**kw), feat_norm="group", stable_layer_norm=False, conv_bias=False)


@pytest.mark.parametrize("which", ["small48", "base768"])
def test_wav2vec2_base_variant_matches_oracle(which):
    """wav2vec2-base family (what whisperx loads for en/fr/de/es/it): GroupNorm after conv0 over each window's OWN frames, no
    conv biases, 48-channel positional-conv groups, post-LayerNorm encoder.  Oracle pinned to transformers' Wav2Vec2ForCTC
    (tests/test_oracle_align.py); ragged batch incl. a window shorter than 400 samples; a window in a batch equals its solo run."""
    from manual_whisper_b200.alignment import AlignEngine
    if which == "small48":
        dims = _base_dims(name="w2v-base-test", n_layers=2, d_model=384, n_heads=6, ffn=512, vocab=40, conv_dim=128, pos_kernel=16, pos_groups=8)
        lens = np.array([16000, 5003, 40000, 250], dtype=np.int32)
    else:       # facebook/wav2vec2-base-960h: 12 layers, d 768, 12 heads, ffn 3072, 16 positional groups of 48 channels
        dims = _base_dims(name="w2v-base", n_layers=12, d_model=768, n_heads=12, ffn=3072, vocab=32, conv_dim=512, pos_kernel=128, pos_groups=16)
        lens = np.array([24000, 9001], dtype=np.int32)
    sd = random_init_w2v(dims, seed=13)
    eng = AlignEngine(dims, sd, device_index=0, max_batch=4, max_samples=40000)
    audio = _audio(70000, 3)
    offs = np.array([0, 16000, 30000, 69000][: len(lens)], dtype=np.int64)
    em, frames = eng.emissions(torch.from_numpy(audio).cuda(), offs, lens)
    em = em.cpu()
    for c in range(len(lens)):
        T = int(frames[c])
        wave = audio[offs[c]: offs[c] + lens[c]]
        ref = _oracle_emissions(dims, sd, wave, emulate=True)
        f32 = _oracle_emissions(dims, sd, wave, emulate=False)
        assert ref.shape == (T, dims.vocab)
        got = em[c, :T]
        assert torch.allclose(got.exp().sum(-1), torch.ones(T), atol=1e-4)
        spread = (f32.max() - f32.min()).item()
        assert (got - ref).abs().max().item() <= 0.02 * spread, (which, c, (got - ref).abs().max().item(), spread)
        assert (got - f32).abs().max().item() <= 0.05 * spread
    solo, f1 = eng.emissions(torch.from_numpy(audio).cuda(), offs[1:2], lens[1:2])
    assert int(f1[0]) == int(frames[1])
    assert (solo[0, : int(f1[0])].cpu() - em[1, : int(frames[1])]).abs().max().item() < 1e-4      # GroupNorm statistics are per window


def test_batched_window_equals_solo_run(small):
    sd, eng = small
    audio = _audio(60000, 1)
    d_audio = torch.from_numpy(audio).cuda()
    em_b, fr_b = eng.emissions(d_audio, np.array([1000, 20000], dtype=np.int64), np.array([9000, 40000], dtype=np.int32))
    em_s, fr_s = eng.emissions(d_audio, np.array([1000], dtype=np.int64), np.array([9000], dtype=np.int32))
    T = int(fr_s[0])
    assert fr_b[0] == T
    assert (em_b[0, :T] - em_s[0, :T]).abs().max().item() <= 1e-5


def test_ctc_align_kernel_matches_oracle_dp(small):
    from oracle import align as OA
    _, eng = small
    rng = np.random.default_rng(3)
    n, T, V = 5, 37, SMALL.vocab
    frames = np.array([37, 20, 9, 3, 30], dtype=np.int32)
    em = np.log(rng.dirichlet(np.ones(V) * 0.3, size=(n, T))).astype(np.float32)
    tokens = [list(rng.integers(1, V, size=11)), list(rng.integers(1, V, size=20)), [5, -1, 7, -1], [1, 2, 3, 4, 5], [9]]
    tokens = [[int(x) for x in t] for t in tokens]
    ft, fs, ok = eng.ctc_align(torch.from_numpy(em).cuda(), frames, tokens, blank=0)
    assert list(ok) == [True, True, True, False, True]                     # window 3: 5 tokens in 3 frames
    for c in range(n):
        if not ok[c]:
            continue
        Tc = int(frames[c])
        tr = OA.get_trellis(em[c, :Tc], tokens[c], 0)
        path = OA.backtrack(tr, em[c, :Tc], tokens[c], 0)
        want = OA.frame_tokens(path, Tc)
        assert np.array_equal(ft[c, :Tc], want), c
        assert np.allclose(fs[c, :Tc], [p.score for p in path], rtol=1e-5, atol=1e-7)


def test_align_mirror_end_to_end(small):
    import manual_whisper_b200 as mw
    from oracle import align as OA
    sd, eng = small
    audio = _audio(16000 * 12, 5)
    dictionary = dict(mw.alignment.DEFAULT_DICTIONARY)
    model_a, meta = mw.load_align_model("en", "cuda", model=sd, dims=SMALL, dictionary=dictionary, max_batch=2, max_samples=16000 * 6)
    segs = [{"text": " hello world", "start": 0.5, "end": 3.25}, {"text": "it's ok!", "start": 3.5, "end": 8.0},
            {"text": " ?? ", "start": 8.0, "end": 9.0}, {"text": "late", "start": 20.0, "end": 21.0}]
    res = mw.align(segs, model_a, meta, audio, "cuda", return_char_alignments=True)
    assert set(res) == {"segments", "word_segments"} and len(res["segments"]) == 4
    assert res["word_segments"] == [w for s in res["segments"] for w in s["words"]]
    s0, s1, s2, s3 = res["segments"]
    assert [w["word"] for w in s0["words"]] == ["hello", "world"] and [w["word"] for w in s1["words"]] == ["it's", "ok!"]
    assert s3["words"] == [] and s3["start"] == 20.0                        # starts after the audio ends: passed through
    assert [w["word"] for w in s2["words"]] == ["??"]                        # wildcards still align
    for seg, src in ((s0, segs[0]), (s1, segs[1])):
        times = [(w["start"], w["end"]) for w in seg["words"]]
        assert all(a <= b for a, b in times) and times == sorted(times)
        # upstream's ratio is duration / (frames - 1), so the last character may end one frame (20 ms) past the segment
        assert src["start"] <= times[0][0] and times[-1][1] <= src["end"] + 0.05
        assert all(0.0 <= w["score"] <= 1.0 for w in seg["words"])
        assert "".join(c["char"] for c in seg["chars"]) == src["text"]
    # the same path as the oracle DP run on the engine's own emissions
    f1, f2 = int(0.5 * 16000), int(3.25 * 16000)
    em, frames = model_a.engine.emissions(torch.from_numpy(audio).cuda(), np.array([f1]), np.array([f2 - f1], dtype=np.int32))
    chars, cdx, toks = mw.alignment.preprocess_segment(segs[0]["text"], dictionary, "en")
    e = em[0, : int(frames[0])].cpu().numpy()
    path = OA.backtrack(OA.get_trellis(e, toks, 0), e, toks, 0)
    merged = OA.merge_repeats(path, "".join(chars))
    ratio = (3.25 - 0.5) / (int(frames[0]) - 1)
    want_first = round(merged[0].start * ratio + 0.5, 3)
    assert s0["chars"][1]["start"] == want_first and s0["words"][0]["start"] == want_first


@pytest.mark.parametrize("vocab", [32, 3503])
def test_full_size_model_matches_oracle(vocab):
    """The XLSR-53 large architecture (24 layers, d 1024, conv_dim 512, 128-tap positional conv in 16 groups)."""
    from manual_whisper_b200.alignment import AlignEngine
    dims = W2vDims(vocab=vocab, n_layers=24 if vocab == 32 else 2)
    sd = random_init_w2v(dims, seed=11, std=0.02)
    eng = AlignEngine(dims, sd, device_index=0, max_batch=2, max_samples=16000 * 5)
    audio = _audio(16000 * 8, 9)
    offs, lens = np.array([0, 48000], dtype=np.int64), np.array([16000 * 3 + 77, 16000 * 5], dtype=np.int32)
    em, frames = eng.emissions(torch.from_numpy(audio).cuda(), offs, lens)
    em = em.cpu()
    for c in range(2):
        T = int(frames[c])
        ref = _oracle_emissions(dims, sd, audio[offs[c]: offs[c] + lens[c]], emulate=True)
        spread = (ref.max() - ref.min()).item()
        err = (em[c, :T] - ref).abs().max().item()
        assert err <= 0.03 * spread, (c, err, spread)


def test_ctc_kernel_against_golden_vector(small):
    import os
    _, eng = small
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_rows_golden.npz"))
    em = g["ctc_emission"]
    pad = np.full((em.shape[0], SMALL.vocab), -30.0, dtype=np.float32)      # the engine's vocabulary is wider: unused symbols
    pad[:, : em.shape[1]] = em
    ft, fs, ok = eng.ctc_align(torch.from_numpy(pad[None]).cuda(), np.array([em.shape[0]], dtype=np.int32),
                               [[int(t) for t in g["ctc_tokens"]]], blank=0)
    assert ok[0] and np.array_equal(ft[0], g["ctc_frame_tokens"])
    assert np.allclose(fs[0], g["ctc_frame_scores"], rtol=1e-5, atol=1e-7)
