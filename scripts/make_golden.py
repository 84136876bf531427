"""Generates the committed fixtures under tests/golden/ (run once in the dev container, CPU only).

Sources of truth:
  * log-mel: the Hugging Face twin of whisperx.audio.log_mel_spectrogram
    (transformers/models/whisper/feature_extraction_whisper.py:135-163, `_torch_extract_fbank_features`),
    which IS importable here; the reference's own whisperx is not.
  * timestamp rules: transformers.generation.logits_process.WhisperTimeStampLogitsProcessor.
  * encoder / decoder / search: the oracle itself on seeded weights (regression pins; seeds recorded).
  * merge_chunks: hand-built cases from SURVEY.md §8c(4).
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import WhisperFeatureExtractor
from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor
from manual_whisper_b200.config import custom_dims, scaled_tokens
from manual_whisper_b200.weights import random_init
from oracle.model import OracleWhisper
from oracle.generate import generate, GenOptions
from oracle.vad import merge_chunks

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)


def audio_case(name, n=480000):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    t = np.arange(n) / 16000.0
    if name == "noise":
        return (0.1 * rng.standard_normal(n)).astype(np.float32)
    if name == "sweep":
        return (0.5 * np.sin(2 * np.pi * (100.0 + 120.0 * t) * t)).astype(np.float32)
    if name == "zeros":
        return np.zeros(n, np.float32)
    if name == "impulse0":
        a = np.zeros(n, np.float32); a[0] = 1.0; return a
    if name == "impulseN":
        a = np.zeros(n, np.float32); a[-1] = 1.0; return a
    if name == "speechlike":
        am = 0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t)
        return (0.1 * am * rng.standard_normal(n) + 0.05 * np.sin(2 * np.pi * 220 * t)).astype(np.float32)
    raise KeyError(name)


def hf_logmel(audio, n_mels, padding):
    fe = WhisperFeatureExtractor(feature_size=n_mels)
    a = np.concatenate([audio, np.zeros(padding, np.float32)]) if padding else audio
    return fe._torch_extract_fbank_features(a)          # [n_mels, frames]


out = {}
SUB = 37   # keep every 37th element of the flattened output (fixtures stay small)
for name in ("noise", "sweep", "zeros", "impulse0", "impulseN", "speechlike"):
    for nm in (80, 128):
        ref = hf_logmel(audio_case(name), nm, 0)
        out[f"{name}_{nm}_sub"] = ref.reshape(-1)[::SUB].astype(np.float32)
        out[f"{name}_{nm}_sum"] = np.array([ref.astype(np.float64).sum(), ref.max(), ref.min()])
for n in (1, 399, 400, 16000, 479999):
    a = audio_case("noise")[:n]
    ref = hf_logmel(a, 128, 480000 - n)
    out[f"len{n}_128_sub"] = ref.reshape(-1)[::SUB].astype(np.float32)
    out[f"len{n}_128_sum"] = np.array([ref.astype(np.float64).sum(), ref.max(), ref.min()])
np.savez_compressed(os.path.join(G, "logmel_golden.npz"), **out)

# ---- merge_chunks
cases = {
    "empty": [],
    "single_short": [(1.0, 4.0)],
    "single_long": [(0.0, 42.5)],
    "exact_30": [(0.0, 10.0), (10.5, 30.0), (30.5, 31.0)],
    "just_over": [(0.0, 10.0), (10.5, 30.001), (30.5, 31.0)],
    "many": [(0.5, 7.0), (8.0, 19.0), (19.4, 28.0), (29.0, 41.0), (42.0, 55.0), (56.0, 71.9), (72.0, 73.0)],
    "gap_start": [(12.0, 20.0), (21.0, 41.0), (50.0, 55.0)],
}
json.dump({k: {"segments": v, "chunk_size": 30, "expected": merge_chunks(v, 30)} for k, v in cases.items()},
          open(os.path.join(G, "merge_chunks.json"), "w"), indent=1)

# ---- timestamp rules: masks from the HF processor on hand-built histories
class _Cfg: pass
tok = scaled_tokens(2048)
gc = _Cfg(); gc.eos_token_id = tok.eot; gc.no_timestamps_token_id = tok.no_timestamps; gc.max_initial_timestamp_index = 50
tb = tok.timestamp_begin
hist = {"first_step": [], "one_text": [300], "lone_ts": [tb + 5], "ts_pair": [tb + 5, tb + 5], "text_after_ts": [tb + 2, 300, 301],
        "closed_then_open": [tb + 1, 300, tb + 9, tb + 9, 400, tb + 20], "two_text": [300, 301]}
rules = {}
g = torch.Generator().manual_seed(5)
for name, h in hist.items():
    P = 3
    proc = WhisperTimeStampLogitsProcessor(gc, begin_index=P)
    ids = torch.tensor([[tok.sot, tok.sot + 1, tok.transcribe] + h])
    scores = torch.randn(1, tok.vocab, generator=g)
    if name in ("lone_ts",):
        pass
    masked = proc(ids, scores.clone())
    rules[name] = {"history": h, "seed_scores": scores[0].tolist(), "masked_is_inf": torch.isinf(masked[0]).nonzero().flatten().tolist()}
# a case where timestamp mass dominates
scores = torch.full((1, tok.vocab), -5.0); scores[0, tb:] = 0.0; scores[0, 300] = 1.0
proc = WhisperTimeStampLogitsProcessor(gc, begin_index=3)
masked = proc(torch.tensor([[tok.sot, tok.sot + 1, tok.transcribe, 300, 301]]), scores.clone())
rules["ts_mass"] = {"history": [300, 301], "seed_scores": scores[0].tolist(), "masked_is_inf": torch.isinf(masked[0]).nonzero().flatten().tolist()}
json.dump({"vocab": tok.vocab, "timestamp_begin": tb, "eot": tok.eot, "no_timestamps": tok.no_timestamps, "cases": rules},
          open(os.path.join(G, "timestamp_rules.json"), "w"))

# ---- small model regression pins (oracle on seeded weights)
dims = custom_dims("golden-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=100, n_text_ctx=32)
sd = random_init(dims, seed=11, scheme="lively")
orc = OracleWhisper(dims, sd)
gm = torch.Generator().manual_seed(3)
mel = (torch.randn(2, 80, 200, generator=gm) * 0.5).clamp(-1.5, 1.5)
with torch.no_grad():
    enc, layers = orc.encode(mel, return_layers=True)
    toks = torch.randint(0, 2048, (2, 6), generator=gm)
    logits = orc.decode(toks, 0, orc.cross_kv(enc), orc.new_cache())
    pr = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
    greedy = generate(orc, enc, pr, tok, GenOptions(beam_size=1, max_length=32))
    beam = generate(orc, enc, pr, tok, GenOptions(beam_size=5, max_length=32))
    beam_ts = generate(orc, enc, pr[:-1], tok, GenOptions(beam_size=5, max_length=32))
np.savez_compressed(os.path.join(G, "small_model.npz"), mel=mel.numpy(), enc_sub=enc.reshape(-1)[::17].numpy(),
                    layer_sums=np.array([l.double().sum().item() for l in layers]), tokens=toks.numpy(),
                    logits_sub=logits.reshape(-1)[::97].numpy(),
                    greedy=np.array([r.sequences_ids[0] for r in greedy]), beam=np.array([r.sequences_ids[0] for r in beam]),
                    beam_scores=np.array([r.scores[0] for r in beam]),
                    beam_ts=np.array([r.sequences_ids[0] + [-1] * (16 - len(r.sequences_ids[0])) for r in beam_ts]))
print("golden written:", sorted(os.listdir(G)), {f: os.path.getsize(os.path.join(G, f)) for f in os.listdir(G)})
