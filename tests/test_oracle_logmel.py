"""Pins the log-mel oracle: against the Hugging Face twin (importable here), the committed golden vectors,
a float64 direct-DFT restatement, and checks the CUDA kernel's own stage code through the CPU emulation."""
import subprocess

import numpy as np
import pytest
import torch

from conftest import audio_case
from oracle.logmel import log_mel_spectrogram, log_mel_float64, mel_filters, log_mel_chunks

SUB = 37
CASES = ("noise", "sweep", "zeros", "impulse0", "impulseN", "speechlike")


@pytest.mark.parametrize("n_mels", [80, 128])
def test_filters_match_hf_and_product(n_mels):
    from transformers.audio_utils import mel_filter_bank
    from manual_whisper_b200.audio import mel_filters_np
    hf = mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
    assert np.array_equal(mel_filters(n_mels), hf)
    assert np.array_equal(mel_filters_np(n_mels), hf)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("n_mels", [80, 128])
def test_oracle_matches_golden(name, n_mels, golden_logmel):
    got = log_mel_spectrogram(audio_case(name), n_mels).numpy()
    assert got.shape == (n_mels, 3000)
    np.testing.assert_allclose(got.reshape(-1)[::SUB], golden_logmel[f"{name}_{n_mels}_sub"], atol=0, rtol=0)
    s = golden_logmel[f"{name}_{n_mels}_sum"]
    assert abs(got.astype(np.float64).sum() - s[0]) <= 1e-6 * max(1.0, abs(s[0]))
    assert got.max() == np.float32(s[1]) and got.min() == np.float32(s[2])


@pytest.mark.parametrize("n", [1, 399, 400, 16000, 479999])
def test_ragged_lengths_match_golden(n, golden_logmel):
    a = audio_case("noise")[:n]
    got = log_mel_spectrogram(a, 128, padding=480000 - n).numpy()
    np.testing.assert_array_equal(got.reshape(-1)[::SUB], golden_logmel[f"len{n}_128_sub"])


def test_oracle_equals_hf_torch_path_live():
    from transformers import WhisperFeatureExtractor
    a = audio_case("speechlike")
    fe = WhisperFeatureExtractor(feature_size=128)
    assert np.array_equal(fe._torch_extract_fbank_features(a), log_mel_spectrogram(a, 128).numpy())


def test_all_zero_hits_the_clamp():
    out = log_mel_spectrogram(np.zeros(480000, np.float32), 80).numpy()
    assert np.all(out == np.float32(-1.5))        # (log10(1e-10) + 4) / 4


def test_range_and_float64_agreement():
    a = audio_case("speechlike")
    out = log_mel_spectrogram(a, 128).numpy()
    assert out.max() - out.min() <= 2.0 + 1e-6
    ref64 = log_mel_float64(a[:48000], 128)
    got = log_mel_spectrogram(a[:48000], 128).numpy()
    assert np.abs(got - ref64).max() < 5e-5


def test_chunk_semantics_each_chunk_has_its_own_max():
    a = np.concatenate([audio_case("noise")[:100000], 1e-3 * audio_case("noise")[:100000]])
    out = log_mel_chunks(a, [0, 100000], [100000, 100000], 80)
    assert out.shape == (2, 80, 3000)
    assert abs(out[0].max() - out[1].max()) > 0.5


@pytest.mark.parametrize("name,n_mels,n,padding", [("noise", 128, 480000, 0), ("sweep", 80, 480000, 0), ("zeros", 128, 480000, 0),
                                                   ("noise", 128, 1, 479999), ("noise", 80, 399, 479601), ("noise", 128, 33333, 0),
                                                   ("impulse0", 80, 480000, 0), ("impulseN", 128, 480000, 0)])
def test_cuda_stage_code_via_cpu_emulation(name, n_mels, n, padding, logmel_emu, tmp_path):
    """Runs csrc/logmel_core.cuh (the exact per-thread stage functions the kernel executes) on the CPU."""
    a = audio_case(name)[:n]
    (tmp_path / "a.f32").write_bytes(a.tobytes())
    (tmp_path / "f.f32").write_bytes(mel_filters(n_mels).tobytes())
    subprocess.check_call([logmel_emu, str(tmp_path / "a.f32"), str(n), str(n + padding), str(n_mels), str(tmp_path / "f.f32"),
                           str(tmp_path / "o.f32")])
    got = np.fromfile(tmp_path / "o.f32", np.float32).reshape(n_mels, (n + padding) // 160)
    ref = log_mel_spectrogram(a, n_mels, padding).numpy()
    assert np.abs(got - ref).max() < 1e-4


@pytest.mark.parametrize("name,n,padding,n_mels", [("noise", 48000, 0, 128), ("speechlike", 32000, 16000, 80), ("noise", 399, 9601, 128),
                                                   ("zeros", 16000, 0, 80), ("impulse0", 16000, 0, 128)])
def test_plain_c_oracle_agrees_with_torch_oracle(name, n, padding, n_mels, tmp_path):
    """oracle/logmel_ref.c: double-precision direct DFT + its own slaney filterbank, no torch involved."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "oracle_logmel_ref"
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", str(exe), os.path.join(root, "oracle", "logmel_ref.c"), "-lm"])
    a = audio_case(name)[:n]
    (tmp_path / "a.f32").write_bytes(a.tobytes())
    subprocess.check_call([str(exe), str(tmp_path / "a.f32"), str(n), str(padding), str(n_mels), str(tmp_path / "o.f32")])
    got = np.fromfile(tmp_path / "o.f32", np.float32).reshape(n_mels, (n + padding) // 160)
    assert np.abs(got - log_mel_spectrogram(a, n_mels, padding).numpy()).max() < 1e-4
