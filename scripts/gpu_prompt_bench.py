"""large-v3, B=32: cost of a long prompt (the reference passes an initial_prompt, transcribe.py:40,111)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.engine import Engine
from bench import device_weights
B = 32
dev = torch.device("cuda:0"); dims = model_dims("large-v3"); tok = special_tokens(dims.vocab)
eng = Engine(dims, device_weights(dims, dev, 1234), 0, max_batch=B)
enc = eng.encode(torch.randn(B, 128, 3000, device=dev) * 0.5)
short = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
long = [tok.sot_prev] + np.random.default_rng(0).integers(1000, 40000, size=80).tolist() + short
res = {}
for name, p in (("prompt4", short), ("prompt85", long)):
    eng.generate(enc, p, tok, beam_size=1, max_length=len(p) + 8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = eng.generate(enc, p, tok, beam_size=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = {"seconds": dt, "new_tokens": len(out[0].sequences_ids[0])}
print(json.dumps(res))
