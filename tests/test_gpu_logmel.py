"""GPU parity of the fused log-mel kernels (C ABI: mw_logmel / mw_logmel_long) against the oracle and the
golden vectors; tolerance 1e-4 abs (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from manual_whisper_b200 import _lib

from conftest import audio_case

pytestmark = pytest.mark.gpu
TOL = 1e-4
SUB = 37


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["noise", "sweep", "zeros", "impulse0", "impulseN", "speechlike"])
@pytest.mark.parametrize("n_mels", [80, 128])
def test_matches_golden_and_oracle(name, n_mels, golden_logmel, dev):
    import manual_whisper_b200 as mw
    from oracle.logmel import log_mel_spectrogram as oracle
    a = audio_case(name)
    got = mw.log_mel_spectrogram(a, n_mels, device=dev)
    assert got.is_cuda and got.dtype == torch.float32 and got.shape == (n_mels, 3000)
    got = got.cpu().numpy()
    assert np.abs(got.reshape(-1)[::SUB] - golden_logmel[f"{name}_{n_mels}_sub"]).max() < TOL
    assert np.abs(got - oracle(a, n_mels).numpy()).max() < TOL
    if name == "zeros":
        assert np.all(got == np.float32(-1.5))


@pytest.mark.parametrize("n", [1, 399, 400, 16000, 479999, 480000])
def test_ragged_lengths_padded_to_30s(n, dev):
    import manual_whisper_b200 as mw
    from oracle.logmel import log_mel_spectrogram as oracle
    a = audio_case("noise")[:n]
    got = mw.log_mel_spectrogram(a, 128, padding=480000 - n, device=dev).cpu().numpy()
    assert got.shape == (128, 3000)
    assert np.abs(got - oracle(a, 128, padding=480000 - n).numpy()).max() < TOL


@pytest.mark.parametrize("n", [201, 16000 + 77, 160000, 16000 * 300])
def test_unchunked_lengths(n, dev):
    import manual_whisper_b200 as mw
    from oracle.logmel import log_mel_spectrogram as oracle
    rng = np.random.default_rng(n)
    a = (0.1 * rng.standard_normal(n)).astype(np.float32)
    got = mw.log_mel_spectrogram(a, 80, device=dev).cpu().numpy()
    assert got.shape == (80, n // 160)
    assert np.abs(got - oracle(a, 80).numpy()).max() < TOL


def test_too_short_input_raises_like_torch_stft(dev):
    import manual_whisper_b200 as mw
    with pytest.raises(ValueError, match="too short"):
        mw.log_mel_spectrogram(np.zeros(100, np.float32), 80, device=dev)
    with pytest.raises(RuntimeError, match="CUDA"):
        mw.log_mel_spectrogram(np.zeros(16000, np.float32), 80, device="cpu")


def test_chunked_api_matches_oracle_and_time_major_copy(dev):
    from manual_whisper_b200.audio import LogMelPlan
    from oracle.logmel import log_mel_chunks
    rng = np.random.default_rng(0)
    N = 16000 * 200
    audio = (0.1 * rng.standard_normal(N)).astype(np.float32)
    audio[1000000:1300000] *= 1e-3          # a quiet chunk: per-chunk max must differ
    offs = np.array([0, 123457, 1000001, 2000000, N - 5, 500000], dtype=np.int64)
    lens = np.array([480000, 333333, 17, 480000, 5, 0], dtype=np.int32)
    plan = LogMelPlan(128, 0, max_chunks=8)
    out_t = torch.full((6, 3002, 128), 7.0, dtype=_lib.storage_dtype(), device=dev)
    got = plan.chunks(torch.from_numpy(audio).to(dev), torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev), out_t=out_t)
    ref = log_mel_chunks(audio, offs, lens, 128)
    assert (got.cpu() - ref).abs().max().item() < TOL
    assert torch.equal(out_t[:, 1:3001].float().cpu(), got.cpu().transpose(1, 2).to(_lib.storage_dtype()).float())
    assert out_t[:, [0, 3001]].float().abs().max().item() == 0.0
    assert torch.all(got[5] == -1.5)          # empty chunk = silence
    with pytest.raises(ValueError, match="max_chunks"):
        plan.chunks(torch.zeros(10, device=dev), torch.zeros(9, dtype=torch.int64, device=dev),
                    torch.zeros(9, dtype=torch.int32, device=dev))


def test_gain_property_at_full_size(dev):
    """Size-independent property: scaling the waveform by c shifts every un-floored value by 2*log10(c)/4."""
    import manual_whisper_b200 as mw
    rng = np.random.default_rng(5)
    a = (0.05 * rng.standard_normal(16000 * 600)).astype(np.float32)       # 10 minutes
    x = mw.log_mel_spectrogram(a, 128, device=dev)
    y = mw.log_mel_spectrogram(a * 4.0, 128, device=dev)
    assert x.shape == (128, 60000)
    assert (y - x - 2 * np.log10(4.0) / 4).abs().max().item() < 1e-5
    assert (x.max() - x.min()).item() <= 2.0 + 1e-6


def test_chunk_equals_unchunked_on_the_same_slice(dev):
    import manual_whisper_b200 as mw
    from manual_whisper_b200.audio import get_plan
    a = audio_case("speechlike")
    d = torch.from_numpy(a).to(dev)
    plan = get_plan(80, dev)
    c = plan.chunks(d, torch.tensor([4321], device=dev), torch.tensor([200000], dtype=torch.int32, device=dev))[0]
    u = mw.log_mel_spectrogram(a[4321:204321], 80, padding=280000, device=dev)
    assert torch.equal(c, u)


def test_unchunked_clamp_touches_only_the_tiles_that_need_it(dev):
    """mw_logmel_long stores scaled values and revisits only tiles holding a value below max - 8: a clip with loud noise, a
    digital-silence gap (every tile of it clamps) and a faint stretch (no tile clamps) must equal the oracle everywhere, and
    calling again on a longer clip (the per-tile buffer grows) and on a shorter one must too."""
    import manual_whisper_b200 as mw
    from oracle.logmel import log_mel_spectrogram as oracle
    rng = np.random.default_rng(9)
    for secs in (40, 95, 12):
        a = (0.1 * rng.standard_normal(16000 * secs)).astype(np.float32)
        a[16000 * 5: 16000 * 8] = 0.0              # digital silence: -10 -> clamped to max - 8
        a[16000 * 9: 16000 * 11] *= 1e-2           # 40 dB down: stays above the clamp
        got = mw.log_mel_spectrogram(a, 128, device=dev).cpu()
        ref = oracle(a, 128)
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < TOL
        lo = float(ref.max()) - 2.0
        assert abs(float(got.min()) - lo) < TOL and float((got[:, 503:797] - lo).abs().max()) < TOL    # frames wholly inside the gap
