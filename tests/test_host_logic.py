"""Host-side mirror of the whisperx interface: merge_chunks, prompts, options, tokenizer, error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
import manual_whisper_b200 as mw
from manual_whisper_b200.asr import TranscriptionOptions, get_prompt
from manual_whisper_b200.config import model_dims, special_tokens, scaled_tokens, custom_dims
from manual_whisper_b200.tokenizer import Tokenizer
from oracle.vad import merge_chunks as oracle_merge


def test_constants_match_whisperx_audio():
    assert (mw.SAMPLE_RATE, mw.N_FFT, mw.HOP_LENGTH, mw.CHUNK_LENGTH, mw.N_SAMPLES, mw.N_FRAMES) == (16000, 400, 160, 30, 480000, 3000)


def test_merge_chunks_golden_and_oracle_agree():
    cases = json.load(open(os.path.join(GOLDEN, "merge_chunks.json")))
    for name, c in cases.items():
        segs = [tuple(s) for s in c["segments"]]
        got = mw.merge_chunks(segs, c["chunk_size"])
        exp = [{"start": e["start"], "end": e["end"], "segments": [tuple(x) for x in e["segments"]]} for e in c["expected"]]
        assert got == exp, name
        assert oracle_merge(segs, c["chunk_size"]) == exp, name


def test_merge_chunks_shapes():
    assert mw.merge_chunks([], 30) == []
    w = mw.merge_chunks([(0.0, 42.5)], 30)
    assert len(w) == 1 and w[0]["end"] - w[0]["start"] == 42.5            # a single long turn stays whole
    w = mw.merge_chunks([(0, 10), (10.5, 30.0), (30.5, 31)], 30)
    assert [(x["start"], x["end"]) for x in w] == [(0.0, 30.0), (30.5, 31.0)]  # exactly 30 s still merges
    class Seg:
        def __init__(s, a, b): s.start, s.end = a, b
    assert mw.merge_chunks([Seg(1, 2), {"start": 3, "end": 4}], 30)[0]["segments"] == [(1.0, 2.0), (3.0, 4.0)]
    with pytest.raises(ValueError):
        mw.merge_chunks([(0, 1)], 0)


def test_synthetic_speech_is_deterministic_and_windows_are_bounded():
    a1, t1 = mw.synthetic_speech(300.0, seed=1)
    a2, t2 = mw.synthetic_speech(300.0, seed=1)
    assert np.array_equal(a1, a2) and t1 == t2 and a1.dtype == np.float32 and len(a1) == 300 * 16000
    wins = mw.merge_chunks(t1, 30)
    assert all(w["end"] - w["start"] <= 30.0 for w in wins) and len(wins) >= 10


def test_energy_vad_recovers_generator_turns():
    a, turns = mw.synthetic_speech(120.0, seed=4)
    segs = mw.EnergyVad()({"waveform": torch.from_numpy(a)[None], "sample_rate": 16000})
    assert abs(len(segs) - len(turns)) <= 2
    got = np.array([(s.start, s.end) for s in segs][: len(turns)])
    if len(segs) == len(turns):
        assert np.abs(got - np.array(turns)).max() < 0.1


def test_get_prompt_layouts():
    tok = special_tokens(51866)
    t = Tokenizer(tok, True, task="transcribe", language="zh")
    assert get_prompt(t, [], without_timestamps=True) == [50258, 50260, 50360, 50364]
    assert get_prompt(t, [], without_timestamps=False) == [50258, 50260, 50360]
    prev = list(range(1000, 1300))
    p = get_prompt(t, prev, without_timestamps=True)
    assert p[0] == tok.sot_prev and p[1:224] == prev[-223:] and p[224:] == [50258, 50260, 50360, 50364] and len(p) == 228
    p = get_prompt(t, [], without_timestamps=False, prefix="ab")
    assert p[:4] == [50258, 50260, 50360, tok.timestamp_begin] and p[4:] == t.encode(" ab")
    assert get_prompt(t, [], True, hotwords="x")[0] == tok.sot_prev


def test_tokenizer_validation():
    tok = special_tokens(51865)
    with pytest.raises(ValueError):
        Tokenizer(tok, True, task="summarize", language="en")
    with pytest.raises(ValueError):
        Tokenizer(tok, True, task="transcribe", language="xx")
    t = Tokenizer(tok, True, task="translate", language="en")
    assert t.sot_sequence == [50258, 50259, 50358]
    assert t.decode([5, 6, 50257, 50300]) == "5 6"


def test_load_model_argument_errors():
    with pytest.raises(ValueError, match="no CPU path"):
        mw.load_model("tiny", "cpu", compute_type="int8", language="zh")        # the reference's shipped DEVICE
    with pytest.raises(ValueError, match="Invalid model size"):
        mw.load_model("huge", "cuda")
    with pytest.raises(ValueError, match="compute type"):
        mw.load_model("tiny", "cuda", compute_type="fp7")
    with pytest.raises(ValueError, match="vad_method"):
        mw.load_model("tiny", "cuda", compute_type="float16", vad_method="webrtc")
    with pytest.raises(TypeError, match="unexpected option"):
        mw.load_model("tiny", "cuda", compute_type="float16", asr_options={"beam": 3})


def test_default_options_match_whisperx_defaults():
    o = TranscriptionOptions()
    assert (o.beam_size, o.patience, o.length_penalty, o.without_timestamps, o.suppress_blank, o.suppress_tokens,
            o.condition_on_previous_text) == (5, 1, 1, True, True, [-1], False)


def test_model_dims_table():
    d = model_dims("large-v3")
    assert (d.n_mels, d.d_model, d.n_heads, d.enc_layers, d.dec_layers, d.ffn, d.vocab, d.d_head) == (128, 1280, 20, 32, 32, 5120, 51866, 64)
    assert model_dims("tiny").vocab == 51865 and model_dims("large").name == "large-v3"
    with pytest.raises(ValueError):
        custom_dims("bad", 80, 100, 2, 1, 1, 256, 2048)
    tk = scaled_tokens(2048)
    assert tk.eot < tk.sot < tk.translate < tk.no_timestamps < tk.timestamp_begin < tk.vocab


def test_load_audio_wav_roundtrip(tmp_path):
    import wave
    pcm = (np.sin(np.arange(16000) / 20.0) * 12000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    a = mw.load_audio(p)
    assert a.dtype == np.float32 and np.array_equal(a, pcm.astype(np.float32) / 32768.0)
    with pytest.raises(RuntimeError, match="Failed to load audio"):
        mw.load_audio(str(tmp_path / "missing.wav"))


def test_shim_exposes_the_names_transcribe_py_uses():
    import importlib, sys
    shim = os.path.join(os.path.dirname(mw.__file__), "shim")
    sys.path.insert(0, shim)
    try:
        wx = importlib.import_module("whisperx")
        for name in ("load_model", "load_audio", "load_align_model", "align", "DiarizationPipeline", "assign_word_speakers"):
            assert hasattr(wx, name)
        assert wx.align is mw.align and wx.load_align_model is mw.load_align_model
        with pytest.raises(RuntimeError):
            wx.DiarizationPipeline(use_auth_token=None, device="cuda")
    finally:
        sys.path.remove(shim)
        sys.modules.pop("whisperx", None)


def test_load_model_rejects_non_whisper_checkpoint(tmp_path):
    from safetensors.torch import save_file
    bad = str(tmp_path / "model.safetensors")
    save_file({"foo": torch.zeros(2)}, bad)
    with pytest.raises(ValueError, match="not a Hugging Face Whisper checkpoint"):
        mw.load_model("tiny", "cuda", compute_type="float16", language="en", model=bad)


def test_sinc_resample_kernel_matches_oracle_and_wav_rate_error(tmp_path):
    import wave
    from manual_whisper_b200.audio import sinc_resample_kernel
    from oracle.resample import sinc_kernel
    for orig in (48000, 44100, 8000):
        k, lo_hi, width, o, n = sinc_resample_kernel(orig, 16000)
        k2, w2, o2, n2 = sinc_kernel(orig, 16000)
        assert (width, o, n) == (w2, o2, n2) and np.array_equal(k, k2)
        for i in range(n):
            assert np.all(k[i, : lo_hi[i, 0]] == 0) and np.all(k[i, lo_hi[i, 1]:] == 0)
    p = str(tmp_path / "hi.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(np.zeros(9600, np.int16).tobytes())
    if not torch.cuda.is_available() and not __import__("shutil").which("ffmpeg"):
        with pytest.raises(RuntimeError, match="GPU decoder"):
            mw.load_audio(p)


def test_english_only_checkpoints_are_refused():
    import manual_whisper_b200 as mw
    from manual_whisper_b200.config import model_dims
    with pytest.raises(ValueError, match="English-only"):
        model_dims("small.en")
    with pytest.raises(ValueError, match="English-only"):
        mw.load_model("tiny.en", "cuda")


def test_timestamp_rules_follow_the_option_not_the_last_prompt_token():
    from manual_whisper_b200.engine import timestamps_enabled
    from manual_whisper_b200.config import special_tokens
    tok = special_tokens(51866)
    base = [tok.sot, tok.sot + 1, tok.transcribe]
    assert timestamps_enabled(base, tok) is True
    assert timestamps_enabled(base + [tok.no_timestamps], tok) is False
    prefixed = base + [tok.no_timestamps, 400, 401, 402]                      # get_prompt(prefix=...) appends after <|notimestamps|>
    assert timestamps_enabled(prefixed, tok) is False
    assert timestamps_enabled(prefixed, tok, without_timestamps=False) is True     # an explicit option wins
    assert timestamps_enabled([tok.sot_prev, tok.no_timestamps] + base, tok) is True   # only the part after <sot> counts


def test_pipeline_keeps_the_loaded_vocabulary_for_every_tokenizer(tmp_path):
    """tokenizer_file is loaded once and re-used whenever the pipeline rebuilds its Tokenizer (language=None or a call with
    another language); without one, encoding text warns that it is not the Whisper BPE."""
    import tokenizers
    from tokenizers import models, pre_tokenizers
    from manual_whisper_b200.asr import FasterWhisperPipeline, TranscriptionOptions
    from manual_whisper_b200.tokenizer import Tokenizer
    from manual_whisper_b200.config import special_tokens
    t = tokenizers.Tokenizer(models.WordLevel({"hello": 5, "world": 6, "[UNK]": 0}, unk_token="[UNK]"))
    t.pre_tokenizer = pre_tokenizers.Whitespace()
    path = str(tmp_path / "tokenizer.json")
    t.save(path)
    tok = special_tokens(51865)

    class _M:
        tokens = tok
        is_multilingual = True
        device = "cpu"
    hf = tokenizers.Tokenizer.from_file(path)
    pipe = FasterWhisperPipeline(model=_M(), vad=None, vad_params={}, options=TranscriptionOptions(), tokenizer=None, hf_tokenizer=hf)
    pipe._prepare_tokenizer(None, "en", None)
    assert pipe.tokenizer.hf is hf and pipe.tokenizer.encode("hello world") == [5, 6]
    pipe._prepare_tokenizer(None, "zh", None)                                    # language change rebuilds the Tokenizer
    assert pipe.tokenizer.language_code == "zh" and pipe.tokenizer.hf is hf
    with pytest.warns(UserWarning, match="NOT the Whisper BPE"):
        assert Tokenizer(tok, True, language="en").encode("hi") == [104, 105]
