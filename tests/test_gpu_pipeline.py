"""The public call surface on the GPU: load_model(...).transcribe(audio, batch_size) against the oracle pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import manual_whisper_b200 as mw
    from manual_whisper_b200.config import custom_dims, scaled_tokens
    from manual_whisper_b200.weights import random_init
    dims = custom_dims("pipe-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=1500, n_text_ctx=24)
    tok = scaled_tokens(2048)
    sd = random_init(dims, seed=5, scheme="lively")
    audio, turns = mw.synthetic_speech(200.0, seed=2)
    pipe = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1},
                         vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=4)
    return mw, dims, tok, sd, audio, turns, pipe


def test_transcribe_structure_order_and_batch_tail(setup):
    mw, dims, tok, sd, audio, turns, pipe = setup
    wins = mw.merge_chunks(turns, 30)
    assert len(wins) % 4 != 0 or len(wins) > 4
    r4 = pipe.transcribe(audio, batch_size=4, language="en")
    r3 = pipe.transcribe(audio, batch_size=3)
    r1 = pipe.transcribe(audio)                       # batch_size None -> 1
    assert r4["language"] == "en" and len(r4["segments"]) == len(wins)
    for seg, w in zip(r4["segments"], wins):
        assert set(seg) >= {"text", "start", "end"} and seg["start"] == round(w["start"], 3) and seg["end"] == round(w["end"], 3)
        assert isinstance(seg["text"], str) and len(seg["tokens"]) <= 12
    # batching must not change a window's ids (per-chunk max, no cross-chunk state)
    assert [s["tokens"] for s in r4["segments"]] == [s["tokens"] for s in r3["segments"]] == [s["tokens"] for s in r1["segments"]]


def test_transcribe_matches_oracle_pipeline():
    """The public call against the oracle pipeline (oracle log-mel -> oracle encoder -> oracle greedy search), every window's
    ids identical to the plain fp32 oracle: "peaked" weights (manual_whisper_b200/weights.py) give margins of nats, so no
    near-tie exemption is needed or made."""
    import manual_whisper_b200 as mw
    from manual_whisper_b200.config import custom_dims, scaled_tokens
    from manual_whisper_b200.weights import random_init
    from oracle.logmel import log_mel_chunks
    from oracle.model import OracleWhisper
    from oracle.generate import generate, GenOptions
    dims = custom_dims("pipe-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=1500, n_text_ctx=24)
    tok = scaled_tokens(2048)
    sd = random_init(dims, seed=5, scheme="peaked", emb_std=0.3)       # oracle min margin 0.046 nat, median 1.2
    audio, turns = mw.synthetic_speech(200.0, seed=2)
    pipe = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1},
                         vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=4)
    res = pipe.transcribe(audio, batch_size=4)
    wins = mw.merge_chunks(turns, 30)
    offs = [int(w["start"] * 16000) for w in wins]
    lens = [int(w["end"] * 16000) - o for w, o in zip(wins, offs)]
    orc = OracleWhisper(dims, sd)
    prompt = [tok.sot, tok.lang_id("en"), tok.transcribe, tok.no_timestamps]
    with torch.no_grad():
        mel = log_mel_chunks(audio, offs, lens, 80)
        ref, trace = generate(orc, orc.encode(mel), prompt, tok, GenOptions(beam_size=1, max_length=dims.n_text_ctx), return_trace=True)
    margins = torch.stack([t.topk(2, dim=-1).values[:, 0] - t.topk(2, dim=-1).values[:, 1] for t in trace])
    print(f"oracle top-2 margins: min {margins.min().item():.3f} median {margins.median().item():.3f} nats")
    assert [s["tokens"] for s in res["segments"]] == [r.sequences_ids[0] for r in ref]


def test_empty_vad_and_short_audio(setup, capsys):
    mw, dims, tok, sd, audio, turns, pipe = setup
    saved = pipe.vad_model
    try:
        pipe.vad_model = mw.InjectedVad([])
        out = pipe.transcribe(audio[:16000], batch_size=4)
        assert out == {"segments": [], "language": "en"}
        assert "No active speech found in audio" in capsys.readouterr().out
        pipe.vad_model = mw.InjectedVad([(0.0, 0.5)])
        out = pipe.transcribe(audio[:8000], batch_size=4)
        assert len(out["segments"]) == 1 and out["segments"][0]["end"] == 0.5
        pipe.vad_model = mw.InjectedVad([(0.0, 31.0)])
        with pytest.raises(ValueError, match="longer than 30 s"):
            pipe.transcribe(audio, batch_size=4)
    finally:
        pipe.vad_model = saved


def test_default_energy_vad_and_language_detection(setup):
    mw, dims, tok, sd, audio, turns, pipe = setup
    with pytest.warns(UserWarning):
        p2 = mw.load_model("tiny", "cuda", compute_type="float16", asr_options={"beam_size": 1}, model=sd, dims=dims,
                           tokens=tok, max_batch=4, vad_options={"vad_onset": 0.5, "vad_offset": 0.363})
    out = p2.transcribe(audio[: 16000 * 60], batch_size=4)
    assert out["language"] in mw.config.LANGUAGES[: tok.n_langs] and len(out["segments"]) >= 1
    assert p2.tokenizer is None                      # created with language=None -> reset after the call


def test_sharded_helper_single_process(setup):
    from manual_whisper_b200.distributed import transcribe_sharded
    mw, dims, tok, sd, audio, turns, pipe = setup
    full = pipe.transcribe(audio, batch_size=4)
    sh = transcribe_sharded(pipe, audio, 4, rank=0, world=1, language="en")
    assert [s["tokens"] for s in sh["segments"]] == [s["tokens"] for s in full["segments"]]


def test_batches_in_flight_do_not_change_results(setup):
    """Two shared-weight replicas on two streams (the default) give the ids of a single replica."""
    mw, dims, tok, sd, audio, turns, pipe = setup
    one = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1},
                        vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=4, streams_per_device=1)
    three = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1},
                          vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=4, streams_per_device=3)
    assert len(one.replicas) == 1 and len(three.replicas) == 3
    assert three.replicas[1].engine.weights is three.replicas[0].engine.weights        # one copy of the weights
    a = one.transcribe(audio, batch_size=2)
    b = three.transcribe(audio, batch_size=2)
    assert [s["tokens"] for s in a["segments"]] == [s["tokens"] for s in b["segments"]]
    assert three.last_stats["replicas"] == 3


def test_beam_search_on_concurrent_replicas_matches_single_stream(setup):
    """Beam 5 (the whisperx default) with 6 batches in flight: the step counter is advanced by a kernel none of whose
    CTAs reads it, so late-scheduled CTAs of the re-parenting kernel cannot see the next step's position."""
    mw, dims, tok, sd, audio, turns, pipe = setup
    kw = dict(compute_type="float16", language="en", asr_options={"beam_size": 5, "without_timestamps": False},
              vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=2)
    one = mw.load_model("tiny", "cuda", streams_per_device=1, **kw)
    six = mw.load_model("tiny", "cuda", streams_per_device=6, **kw)
    a = one.transcribe(audio, batch_size=2)
    for _ in range(2):
        b = six.transcribe(audio, batch_size=2)
        assert [s["tokens"] for s in a["segments"]] == [s["tokens"] for s in b["segments"]]
    assert six.last_stats["replicas"] >= 4


def test_device_vad_windows_equal_the_host_twin_exactly():
    """mw_vad_windows (score, Binarize hysteresis, gap fill, merge_chunks on the device) against the host twin fed with the
    SAME device frame energies: turns and windows bit-equal, incl. a 1-hour recording, a turn longer than 30 s that must be
    cut at max_duration, silence, and clips shorter than a frame."""
    import manual_whisper_b200 as mw
    from manual_whisper_b200.vad import merge_chunks
    import ctypes as C
    from manual_whisper_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    hour, _ = mw.synthetic_speech(3600.0, seed=1)
    long_turn = (0.1 * rng.standard_normal(16000 * 90)).astype(np.float32)            # 75 s of continuous "speech" ...
    long_turn[: 16000 * 15] *= 1e-3                                                   # ... after 15 s of background
    cases = {"hour": hour, "200s": mw.synthetic_speech(200.0, seed=2)[0], "long_turn": long_turn,
             "silence": np.zeros(16000 * 5, np.float32), "tiny": np.zeros(100, np.float32)}
    for name, a in cases.items():
        vad = mw.GpuEnergyVad()
        d = torch.from_numpy(a).cuda()
        n = d.numel() // vad.frame
        rms = torch.empty(max(n, 1), dtype=torch.float32, device="cuda")
        if n:
            _lib.check(lib.mw_frame_rms(d.data_ptr(), d.numel(), vad.frame, rms.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rms")
        want_turns = mw.EnergyVad().turns_from_rms(rms[:n].cpu().numpy())
        got_turns, got_wins = vad._run({"waveform": d[None], "sample_rate": 16000}, 30.0)
        assert got_turns.tolist() == want_turns, name
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            want_wins = merge_chunks(want_turns, 30.0)
        assert got_wins.tolist() == [[w["start"], w["end"]] for w in want_wins], name
        if name == "hour":
            assert 100 <= len(got_wins) <= 160 and max(b - a_ for a_, b in got_wins.tolist()) <= 30.0 + 1e-9
        if name == "long_turn":
            assert len(got_turns) >= 3 and max(b - a_ for a_, b in got_turns.tolist()) <= 30.0 + 1e-9
        if name in ("silence", "tiny"):
            assert len(got_wins) == 0 or name == "silence"
        print(f"[vad {name}] {len(got_turns)} turns -> {len(got_wins)} windows")


def test_device_side_vad_front_end(setup):
    """the VAD front end on the GPU gives the same turns as the host VAD and the same transcript."""
    mw, dims, tok, sd, audio, turns, pipe = setup
    a = audio[: 16000 * 90]
    host = mw.EnergyVad()({"waveform": torch.from_numpy(a)[None], "sample_rate": 16000})
    dev = mw.GpuEnergyVad()({"waveform": torch.from_numpy(a).cuda()[None], "sample_rate": 16000})
    assert [(round(s.start, 2), round(s.end, 2)) for s in host] == [(round(s.start, 2), round(s.end, 2)) for s in dev]
    p_host = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1}, model=sd,
                           dims=dims, tokens=tok, max_batch=4, vad_method="energy", streams_per_device=1)
    p_dev = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1}, model=sd,
                          dims=dims, tokens=tok, max_batch=4, vad_method="energy_gpu", streams_per_device=1)
    r1, r2 = p_host.transcribe(a, batch_size=4), p_dev.transcribe(a, batch_size=4)
    assert [(s["start"], s["end"], s["tokens"]) for s in r1["segments"]] == [(s["start"], s["end"], s["tokens"]) for s in r2["segments"]]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_in_process_multi_gpu_replicas(setup):
    """device_index=[0, 1]: one replica set per GPU in one process; batches are dealt out dynamically, order restored."""
    mw, dims, tok, sd, audio, turns, pipe = setup
    two = mw.load_model("tiny", "cuda", device_index=[0, 1], compute_type="float16", language="en", asr_options={"beam_size": 1},
                        vad_model=mw.InjectedVad(turns), model=sd, dims=dims, tokens=tok, max_batch=4, streams_per_device=2)
    assert sorted({r.device.index for r in two.replicas}) == [0, 1] and len(two.replicas) == 4
    a = pipe.transcribe(audio, batch_size=2)
    b = two.transcribe(audio, batch_size=2)
    assert [s["tokens"] for s in a["segments"]] == [s["tokens"] for s in b["segments"]]
    assert [(s["start"], s["end"]) for s in a["segments"]] == [(s["start"], s["end"]) for s in b["segments"]]


def test_load_model_from_safetensors_checkpoint(setup, tmp_path):
    """A Hugging Face ``model.safetensors`` path (or one under download_root) loads to the same engine as its state dict."""
    from safetensors.torch import save_file
    mw, dims, tok, sd, audio, turns, pipe = setup
    path = tmp_path / "tiny" / "model.safetensors"
    path.parent.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(path))
    kw = dict(compute_type="float16", language="en", asr_options={"beam_size": 1}, vad_model=mw.InjectedVad(turns), dims=dims,
              tokens=tok, max_batch=4)
    by_path = mw.load_model("tiny", "cuda", model=str(path), **kw)
    by_root = mw.load_model("tiny", "cuda", download_root=str(tmp_path), **kw)
    want = [s["tokens"] for s in pipe.transcribe(audio, batch_size=4)["segments"]]
    assert [s["tokens"] for s in by_path.transcribe(audio, batch_size=4)["segments"]] == want
    assert [s["tokens"] for s in by_root.transcribe(audio, batch_size=4)["segments"]] == want


def test_reference_flow_through_the_whisperx_shim(setup, monkeypatch):
    """/root/reference/transcribe.py:107-131 verbatim in shape: load_model -> transcribe -> load_align_model -> align, with
    `import whisperx` resolving to the shim."""
    import importlib, os, sys
    mw, dims, tok, sd, audio, turns, pipe = setup
    from manual_whisper_b200 import alignment
    from manual_whisper_b200.w2v import W2vDims
    monkeypatch.setattr(alignment, "DEFAULT_ALIGN_DIMS",
                        W2vDims(name="w2v-test", n_layers=2, d_model=128, n_heads=2, ffn=256, vocab=32, conv_dim=128, pos_kernel=16, pos_groups=2))
    shim = os.path.join(os.path.dirname(mw.__file__), "shim")
    sys.path.insert(0, shim)
    try:
        whisperx = importlib.import_module("whisperx")
        result = pipe.transcribe(audio, batch_size=4, language="en")
        with pytest.warns(UserWarning, match="random-init"):
            model_a, metadata = whisperx.load_align_model(language_code="en", device="cuda")
        aligned = whisperx.align(result["segments"], model_a, metadata, audio, "cuda", return_char_alignments=False)
    finally:
        sys.path.remove(shim)
        sys.modules.pop("whisperx", None)
    assert set(aligned) == {"segments", "word_segments"} and len(aligned["segments"]) >= len(result["segments"])
    for seg in aligned["segments"]:
        assert {"start", "end", "text", "words"} <= set(seg)
        assert all("word" in w for w in seg["words"])
    assert sum(len(s["words"]) for s in aligned["segments"]) == len(aligned["word_segments"]) > 0
