"""large-v3, B=32 windows, beam 5 (whisperx's default asr_options): decode timing."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.engine import Engine
from bench import device_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 448
dev = torch.device("cuda:0"); dims = model_dims("large-v3"); tok = special_tokens(dims.vocab)
eng = Engine(dims, device_weights(dims, dev, 1234), 0, max_batch=B, max_beam=5)
enc = eng.encode(torch.randn(B, 128, 3000, device=dev) * 0.5)
prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
res = {}
for beam in ([int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (1, 5)):
    eng.generate(enc, prompt, tok, beam_size=beam, max_length=32)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = eng.generate(enc, prompt, tok, beam_size=beam, max_length=steps)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = len(out[0].sequences_ids[0])
    res[f"beam{beam}"] = {"seconds": dt, "tokens": n, "ms_per_step": dt / max(n, 1) * 1e3}
print(json.dumps(res, indent=1), "workspace GB", eng.workspace_bytes / 1e9)
