"""Generates tests/golden/next_rows_golden.npz: fixtures for the SURVEY.md §8f rows built after the hot path (run once in
the dev container, CPU only).

Sources of truth:
  * wav2vec2-CTC logits: transformers' Wav2Vec2ForCTC (modeling_wav2vec2.py) on seeded weights (manual_whisper_b200.w2v:
    random_init_w2v, seed 3, the SMALL architecture of tests/test_oracle_align.py) and a seeded waveform;
  * resampling: torchaudio.functional.resample on seeded noise, 44.1 kHz and 48 kHz -> 16 kHz;
  * CTC trellis/backtrack: the oracle itself (regression pin; its optimum is checked against brute force in
    tests/test_oracle_align.py) - the reference holds no fixture for it.
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torchaudio.functional as F
from manual_whisper_b200.w2v import random_init_w2v
from oracle import align as OA
from test_oracle_align import SMALL, _hf_twin

out = {}
sd = random_init_w2v(SMALL, seed=3)
hf = _hf_twin(SMALL, sd)
g = torch.Generator().manual_seed(123)
wave = torch.randn(3217, generator=g) * 0.1
with torch.no_grad():
    out["w2v_wave"] = wave.numpy()
    out["w2v_logits"] = hf(wave[None]).logits[0].numpy()
rng = np.random.default_rng(7)
for rate in (44100, 48000):
    x = (rng.standard_normal(rate // 4 + 5) * 0.3).astype(np.float32)
    out[f"resample_in_{rate}"] = x
    out[f"resample_out_{rate}"] = F.resample(torch.from_numpy(x), rate, 16000).numpy()
em = np.log(rng.dirichlet(np.ones(12) * 0.4, size=40)).astype(np.float32)
tokens = np.array([3, 7, -1, 2, 9, 9, 1], dtype=np.int32)
tr = OA.get_trellis(em, list(tokens), 0)
path = OA.backtrack(tr, em, list(tokens), 0)
out["ctc_emission"], out["ctc_tokens"] = em, tokens
out["ctc_final_score"] = np.float32(tr[-1, -1])
out["ctc_frame_tokens"] = OA.frame_tokens(path, 40)
out["ctc_frame_scores"] = np.array([p.score for p in path], dtype=np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "next_rows_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
