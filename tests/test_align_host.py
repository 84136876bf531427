"""Host-side book-keeping of the forced-alignment mirror (manual_whisper_b200/alignment.py): no GPU needed."""
import math

import numpy as np
import pytest
import torch

from manual_whisper_b200 import _lib

from manual_whisper_b200 import alignment as AL
from manual_whisper_b200.w2v import W2vDims, random_init_w2v, effective_pos_conv_weight


def test_preprocess_segment_marks_unknown_characters_as_wildcards():
    d = AL.DEFAULT_DICTIONARY
    chars, cdx, toks = AL.preprocess_segment("  Hi, you ", d, "en")
    assert chars == ["h", "i", "*", "|", "y", "o", "u"] and cdx == [2, 3, 4, 5, 6, 7, 8]
    assert toks == [d["h"], d["i"], AL.WILDCARD, d["|"], d["y"], d["o"], d["u"]]
    chars, cdx, toks = AL.preprocess_segment("你好 a", d, "zh")            # no '|' substitution for zh / ja
    assert chars == ["*", "*", "*", "a"] and toks[:3] == [AL.WILDCARD] * 3
    assert AL.preprocess_segment("   ", d, "en") == ([], [], [])


def test_interpolate_nans():
    nan = float("nan")
    assert AL.interpolate_nans([nan, 1.0, nan, nan, 4.0, nan]) == [1.0, 1.0, 1.0, 4.0, 4.0, 4.0]
    assert AL.interpolate_nans([0.0, nan, 2.0], "linear") == [0.0, 1.0, 2.0]
    assert all(math.isnan(v) for v in AL.interpolate_nans([nan, nan]))


def test_chars_from_path():
    ft = np.array([0, 0, 1, 2, 2, 2, 9, 9], dtype=np.int32)
    fs = np.array([0.5, 0.7, 0.2, 0.1, 0.2, 0.3, 0, 0], dtype=np.float32)
    spans = AL.chars_from_path(ft, fs, 6, 3)
    assert [(j, a, b) for j, a, b, _ in spans] == [(0, 0, 2), (1, 2, 3), (2, 3, 6)]
    assert spans[0][3] == pytest.approx(0.6) and spans[2][3] == pytest.approx(0.2)
    assert AL.chars_from_path(ft, fs, 6, 4) is None


def test_assemble_segment_words_and_sentences():
    seg = {"start": 10.0, "end": 12.0, "text": " ab c"}
    d = AL.DEFAULT_DICTIONARY
    chars, cdx, toks = AL.preprocess_segment(seg["text"], d, "en")       # a b | c
    assert chars == ["a", "b", "|", "c"] and cdx == [1, 2, 3, 4]
    spans = [(0, 0, 2, 0.9), (1, 2, 4, 0.8), (2, 4, 5, 0.5), (3, 5, 10, 0.7)]
    out = AL.assemble_segment(seg, seg["text"], cdx, spans, 0.2, "en", [(0, len(seg["text"]))], "nearest", True)
    assert len(out) == 1
    s = out[0]
    assert s["text"] == " ab c" and s["start"] == 10.0 and s["end"] == 12.0
    assert s["words"] == [{"word": "ab", "start": 10.0, "end": 10.8, "score": 0.85},
                          {"word": "c", "start": 11.0, "end": 12.0, "score": 0.7}]
    assert s["chars"][0] == {"char": " "} and s["chars"][1] == {"char": "a", "start": 10.0, "end": 10.4, "score": 0.9}
    # languages without spaces: every character is its own word
    seg = {"start": 0.0, "end": 1.0, "text": "ab"}
    out = AL.assemble_segment(seg, "ab", [0, 1], [(0, 0, 3, 0.5), (1, 3, 5, 0.25)], 0.25, "zh", [(0, 2)], "nearest", False)
    assert [w["word"] for w in out[0]["words"]] == ["a", "b"] and "chars" not in out[0]
    # two sentences with identical timestamps are merged, different ones come out ordered by start
    seg = {"start": 0.0, "end": 4.0, "text": "a. b."}
    cdx = [0, 1, 2, 3, 4]
    spans = [(0, 0, 1, 1.0), (1, 1, 2, 1.0), (2, 2, 3, 1.0), (3, 3, 4, 1.0), (4, 4, 5, 1.0)]
    out = AL.assemble_segment(seg, seg["text"], cdx, spans, 1.0, "en", [(0, 2), (3, 5)], "nearest", False)
    assert [s["text"] for s in out] == ["a.", "b."] and out[0]["start"] == 0.0 and out[1]["end"] == 5.0


def test_pack_w2v_weights_layouts():
    dims = W2vDims(name="t", n_layers=1, d_model=128, n_heads=2, ffn=256, vocab=40, conv_dim=128, pos_kernel=4, pos_groups=2)
    sd = random_init_w2v(dims, seed=0)
    packed = AL.pack_w2v_weights(sd, dims, torch.device("cpu"))
    assert len(packed) == 38 + 12
    assert packed[0].shape == (128, 10) and packed[0].dtype == torch.float32
    w1 = sd["wav2vec2.feature_extractor.conv_layers.1.conv.weight"]
    assert packed[4].shape == (128, 3 * 128) and packed[4].dtype == _lib.storage_dtype()
    assert torch.equal(packed[4].float()[5, 2 * 128 + 7], w1[5, 7, 2])            # [co][tap][ci]
    wp = effective_pos_conv_weight(sd)                                          # [d, 64, taps]
    assert packed[32].shape == (2, 64, 4, 64)
    assert torch.equal(packed[32].float()[1, 3, 2, 9], wp[64 + 3, 9, 2])        # [g][out][tap][in]
    assert packed[36].shape == (64, 128) and torch.all(packed[36][40:] == 0)      # lm_head rows padded to 64
    assert packed[38 + 2].shape == (3 * 128, 128) and packed[38 + 3].shape == (3 * 128,)
    k_bias = sd["wav2vec2.encoder.layers.0.attention.k_proj.bias"]
    assert torch.equal(packed[38 + 3][128:256], k_bias)                           # the key bias is real here


def test_load_align_model_needs_cuda():
    with pytest.raises(ValueError, match="CUDA"):
        AL.load_align_model("zh", "cpu")
    if not torch.cuda.is_available():
        with pytest.warns(UserWarning, match="random-init"), pytest.raises(RuntimeError, match="no CPU fallback"):
            AL.load_align_model("zh", "cuda", dims=W2vDims(name="t", n_layers=1, d_model=128, n_heads=2, ffn=256, vocab=32,
                                                           conv_dim=128, pos_kernel=4, pos_groups=2))


class _OracleEngine:
    """Stands in for AlignEngine on the CPU: oracle emissions + oracle DP behind the same three calls align() makes."""

    def __init__(self, dims, sd, max_batch=2, max_samples=16000 * 8):
        from oracle.wav2vec2 import OracleWav2Vec2
        self.model = OracleWav2Vec2(dims, sd)
        self.dims, self.device, self.max_batch, self.max_samples = dims, torch.device("cpu"), max_batch, max_samples
        self.calls = []

    def emissions(self, d_audio, offs, lens):
        frames = np.array([self.dims.frames(max(int(n), 400)) for n in lens], dtype=np.int32)
        out = torch.zeros(len(offs), int(frames.max()), self.dims.vocab)
        self.calls.append(len(offs))
        for c, (o, n) in enumerate(zip(offs, lens)):
            w = d_audio[int(o): int(o) + int(n)]
            if len(w) < 400:
                w = torch.nn.functional.pad(w, (0, 400 - len(w)))
            with torch.no_grad():
                out[c, : frames[c]] = self.model.emissions(w)
        return out, frames

    def ctc_align(self, em, frames, tokens, blank):
        from oracle import align as OA
        n, T, _ = em.shape
        ft, fs, ok = np.zeros((n, T), np.int32), np.zeros((n, T), np.float32), np.zeros(n, bool)
        for c in range(n):
            e = em[c, : frames[c]].numpy()
            if not (0 < len(tokens[c]) <= frames[c]):
                continue
            path = OA.backtrack(OA.get_trellis(e, tokens[c], blank), e, tokens[c], blank)
            if path is None:
                continue
            ok[c] = True
            ft[c, : frames[c]] = OA.frame_tokens(path, int(frames[c]))
            fs[c, : frames[c]] = [p.score for p in path]
        return ft, fs, ok


def test_align_control_flow_on_cpu_with_an_oracle_engine(capsys):
    from oracle import align as OA
    dims = W2vDims(name="t", n_layers=1, d_model=128, n_heads=2, ffn=256, vocab=32, conv_dim=64, pos_kernel=8, pos_groups=2)
    sd = random_init_w2v(dims, seed=2)
    eng = _OracleEngine(dims, sd)
    model = AL.AlignModel(eng, dict(AL.DEFAULT_DICTIONARY), "en")
    meta = {"language": "en", "dictionary": dict(AL.DEFAULT_DICTIONARY), "type": "b200"}
    rng = np.random.default_rng(0)
    audio = (rng.standard_normal(16000 * 10) * 0.1).astype(np.float32)
    segs = [{"text": " one two", "start": 0.0, "end": 2.0}, {"text": "three", "start": 2.5, "end": 4.0},
            {"text": "   ", "start": 4.0, "end": 4.5}, {"text": "four five six", "start": 5.0, "end": 9.5},
            {"text": "x" * 40, "start": 9.6, "end": 9.7},                    # more characters than frames: not alignable
            {"text": "gone", "start": 11.0, "end": 12.0}]                     # starts after the audio ends
    res = AL.align(segs, model, meta, audio, "cpu", return_char_alignments=True, sentence_splitter=lambda t: [(0, len(t))])
    assert eng.calls == [2, 2]                                                 # 4 alignable segments in batches of max_batch
    out = res["segments"]
    assert [s["text"] for s in out] == [s["text"] for s in segs]
    assert [[w["word"] for w in s["words"]] for s in out] == [["one", "two"], ["three"], [], ["four", "five", "six"], [], []]
    assert res["word_segments"] == [w for s in out for w in s["words"]]
    printed = capsys.readouterr().out
    assert printed.count("Failed to align segment") == 3
    # the first segment, recomputed by hand from the oracle pieces
    chars, cdx, toks = AL.preprocess_segment(segs[0]["text"], meta["dictionary"], "en")
    with torch.no_grad():
        e = eng.model.emissions(torch.from_numpy(audio[: 32000])).numpy()
    merged = OA.merge_repeats(OA.backtrack(OA.get_trellis(e, toks, 0), e, toks, 0), "".join(chars))
    ratio = 2.0 / (e.shape[0] - 1)
    w0 = out[0]["words"][0]
    assert w0["start"] == round(merged[0].start * ratio, 3) and w0["end"] == round(merged[2].end * ratio, 3)
    assert w0["score"] == round(sum(round(m.score, 3) for m in merged[:3]) / 3, 3)
    assert out[0]["chars"][1] == {"char": "o", "start": w0["start"], "end": round(merged[0].end * ratio, 3), "score": round(merged[0].score, 3)}
    assert out[0]["start"] == w0["start"] and out[0]["end"] == out[0]["words"][-1]["end"]


def test_engine_accepts_both_whisperx_families_and_rejects_hybrids():
    """XLSR (layer norm, stable LN, conv bias) and wav2vec2-base (group norm, post-LN, no conv bias) are built; anything in
    between is refused before any device work."""
    from dataclasses import replace
    hybrid = replace(W2vDims(), feat_norm="group", stable_layer_norm=True, conv_bias=False)
    with pytest.raises(NotImplementedError, match="two wav2vec2 families"):
        AL.AlignEngine(hybrid, {}, device_index=0)


def test_pack_base_variant_pads_missing_slots_and_narrow_pos_groups():
    from dataclasses import replace
    dims = replace(W2vDims(name="b", n_layers=1, d_model=384, n_heads=6, ffn=256, vocab=40, conv_dim=128, pos_kernel=8, pos_groups=8),
                   feat_norm="group", stable_layer_norm=False, conv_bias=False)
    sd = random_init_w2v(dims, seed=0)
    assert "wav2vec2.feature_extractor.conv_layers.0.conv.bias" not in sd
    assert "wav2vec2.feature_extractor.conv_layers.1.layer_norm.weight" not in sd
    packed = AL.pack_w2v_weights(sd, dims, torch.device("cpu"))
    assert len(packed) == 38 + 12
    assert torch.all(packed[1] == 0) and torch.all(packed[5] == 0)               # absent conv biases
    assert torch.all(packed[6] == 1) and torch.all(packed[7] == 0)               # conv 1 has no norm in this variant
    assert packed[32].shape == (8, 64, 8, 48) and torch.all(packed[32][:, 48:] == 0)     # 48-channel groups, rows padded to 64
    wp = effective_pos_conv_weight(sd)
    assert torch.equal(packed[32].float()[3, 5, 2, 9], wp[3 * 48 + 5, 9, 2])
