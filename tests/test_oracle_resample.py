"""oracle/resample.py pinned to torchaudio.functional.resample (importable here)."""
import numpy as np
import pytest
import torch

from oracle import resample as R


@pytest.mark.parametrize("orig,new", [(48000, 16000), (44100, 16000), (8000, 16000), (22050, 16000), (32000, 16000)])
def test_kernel_and_output_match_torchaudio(orig, new):
    import math
    import torchaudio.functional as F
    g = math.gcd(orig, new)
    want_k, want_w = F.functional._get_sinc_resample_kernel(orig, new, g)
    k, w, o, n = R.sinc_kernel(orig, new)
    assert w == want_w and k.shape == tuple(want_k[:, 0].shape)
    assert np.array_equal(k, want_k[:, 0].numpy())
    rng = np.random.default_rng(orig)
    x = (rng.standard_normal(orig // 3 + 17) * 0.3).astype(np.float32)
    want = F.resample(torch.from_numpy(x), orig, new).numpy()
    got = R.resample(x, orig, new)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5          # torchaudio's own fp32 conv1d sits up to 9e-6 from the float64 result
    padded = np.concatenate([np.zeros(w), x.astype(np.float64), np.zeros(w + o)])
    ref = (np.lib.stride_tricks.sliding_window_view(padded, k.shape[1])[::o] @ k.T.astype(np.float64)).reshape(-1)[: len(want)]
    assert np.abs(got - ref).max() < 1e-6


def test_decode_pcm16_downmix_quantise_and_identity():
    rng = np.random.default_rng(0)
    pcm = rng.integers(-20000, 20000, size=2 * 4800, dtype=np.int16)
    y = R.decode_pcm16(pcm, 2, 48000)
    assert y.shape == (1600,) and y.dtype == np.float32
    assert np.array_equal(y * 32768.0, np.rint(y * 32768.0))                 # on the int16 grid
    mono = rng.integers(-30000, 30000, size=1000, dtype=np.int16)
    assert np.array_equal(R.decode_pcm16(mono, 1, 16000), mono.astype(np.float32) / 32768.0)
