"""Small large-v3 run for ncu launch lists: one batch of B windows, a short decode."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.weights import _keys
from manual_whisper_b200.engine import Engine
from manual_whisper_b200 import audio as A
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda:0")
dims = model_dims("large-v3"); tok = special_tokens(dims.vocab)
g = torch.Generator(device=dev); g.manual_seed(1)
sd = {}
for name, shape, kind in _keys(dims):
    if kind == "g": sd[name] = torch.ones(shape, device=dev)
    elif kind == "beta": sd[name] = torch.zeros(shape, device=dev)
    else: sd[name] = (torch.randn(shape, device=dev, generator=g) * 0.02)
sd["model.encoder.embed_positions.weight"] = torch.zeros(1500, dims.d_model, device=dev)
eng = Engine(dims, sd, 0, max_batch=B); del sd
plan = A.LogMelPlan(128, 0, max_chunks=B)
audio = torch.randn(B * 480000, device=dev) * 0.1
offs = torch.arange(B, dtype=torch.int64, device=dev) * 480000
lens = torch.full((B,), 480000, dtype=torch.int32, device=dev)
feat = torch.empty(B, 128, 3000, device=dev); feat_t = torch.empty(B, 3002, 128, device=dev, dtype=torch.bfloat16)
prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    plan.chunks(audio, offs, lens, out=feat, out_t=feat_t)
    enc = eng.encode_time_major(feat_t)
    out = eng.generate(enc, prompt, tok, beam_size=1, max_length=max_len)
    torch.cuda.synchronize(); print("iter", it, time.time() - t0, flush=True)
