"""Host-side book-keeping of the forced-alignment mirror (manual_whisper_b200/alignment.py): no GPU needed."""
import math

import numpy as np
import pytest
import torch

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
    assert packed[4].shape == (128, 3 * 128) and packed[4].dtype == torch.bfloat16
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
