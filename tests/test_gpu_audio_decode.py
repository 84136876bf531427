"""Audio decode on the GPU: mw_pcm_resample against oracle/resample.py (pinned to torchaudio)."""
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rate,channels", [(48000, 1), (48000, 2), (44100, 2), (8000, 1), (22050, 3), (16000, 2)])
def test_decode_pcm16_matches_oracle(rate, channels):
    import manual_whisper_b200 as mw
    from oracle.resample import decode_pcm16
    rng = np.random.default_rng(rate + channels)
    n = rate // 2 + 13
    t = np.arange(n) / rate
    sig = 0.4 * np.sin(2 * np.pi * 440 * t)[:, None] + 0.1 * rng.standard_normal((n, channels))
    pcm = np.clip(np.rint(sig * 32768), -32768, 32767).astype(np.int16).reshape(-1)
    for quant in (False, True):
        got = mw.decode_pcm_device(pcm, channels, rate, quantize_s16=quant).cpu().numpy()
        want = decode_pcm16(pcm, channels, rate, 16000, quantize=quant)
        assert got.shape == want.shape
        if quant:       # both on the int16 grid: a rounding tie may flip one step, nothing more
            steps = np.abs(got - want) * 32768
            assert steps.max() <= 1.0 and (steps > 0).mean() < 5e-3     # 2e-6 * 32768 = 0.07 LSB of float noise
        else:
            assert np.abs(got - want).max() < 2e-6


def test_float_pcm_and_identity_rate():
    import manual_whisper_b200 as mw
    from oracle.resample import resample
    x = (np.random.default_rng(0).standard_normal(32000) * 0.2).astype(np.float32)
    got = mw.decode_pcm_device(x, 1, 32000, quantize_s16=False).cpu().numpy()
    assert np.abs(got - resample(x, 32000, 16000)).max() < 2e-6
    same = mw.decode_pcm_device(x, 1, 16000, quantize_s16=False).cpu().numpy()
    assert np.array_equal(same, x)


def test_load_audio_wav_at_other_rates(tmp_path):
    import shutil
    import manual_whisper_b200 as mw
    from oracle.resample import decode_pcm16
    if shutil.which("ffmpeg"):
        pytest.skip("ffmpeg present: load_audio uses it, as the reference does")
    rng = np.random.default_rng(1)
    pcm = rng.integers(-12000, 12000, size=2 * 44100, dtype=np.int16)
    p = str(tmp_path / "stereo44k.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    a = mw.load_audio(p)
    d = mw.load_audio_device(p)
    want = decode_pcm16(pcm, 2, 44100)
    assert a.dtype == np.float32 and a.shape == want.shape == (16000,) and d.is_cuda
    assert np.array_equal(a, d.cpu().numpy()) and np.abs(a - want).max() <= 1.0 / 32768


def test_transcribe_accepts_device_resident_audio():
    import manual_whisper_b200 as mw
    from manual_whisper_b200.config import custom_dims, scaled_tokens
    from manual_whisper_b200.weights import random_init
    dims = custom_dims("pipe-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=1500, n_text_ctx=16)
    audio, turns = mw.synthetic_speech(40.0, seed=4)
    pipe = mw.load_model("tiny", "cuda", compute_type="float16", language="en", asr_options={"beam_size": 1},
                         vad_model=mw.InjectedVad(turns), model=random_init(dims, seed=5, scheme="lively"), dims=dims,
                         tokens=scaled_tokens(2048), max_batch=4)
    a = pipe.transcribe(audio, batch_size=4)
    b = pipe.transcribe(torch.from_numpy(audio).cuda(), batch_size=4)
    assert [s["tokens"] for s in a["segments"]] == [s["tokens"] for s in b["segments"]] and len(a["segments"]) > 0
