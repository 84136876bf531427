"""Short beam-search run (default small, B=16, beam 5, timestamps) for an ncu launch list.  args: MODEL MAXLEN B"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.engine import Engine
from bench import device_weights
name = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda:0"); dims = model_dims(name); tok = special_tokens(dims.vocab)
eng = Engine(dims, device_weights(dims, dev, 1), 0, max_batch=B, max_beam=5)
enc = eng.encode(torch.randn(B, dims.n_mels, 3000, device=dev) * 0.5)
prompt = [tok.sot, tok.lang_id("en"), tok.transcribe]
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    out = eng.generate(enc, prompt, tok, beam_size=5, max_length=int(sys.argv[2]) if len(sys.argv) > 2 else 48)
    torch.cuda.synchronize(); print("iter", it, time.time() - t0, len(out[0].sequences_ids[0]), flush=True)
