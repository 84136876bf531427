"""Quick large-v3 timing: kernels in isolation + one batch end to end (no CPU baseline)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.engine import Engine
from manual_whisper_b200 import audio as A, _lib
from bench import device_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
dims = model_dims("large-v3"); tok = special_tokens(dims.vocab)
eng = Engine(dims, device_weights(dims, dev, 1234), 0, max_batch=B)
plan = A.LogMelPlan(128, 0, max_chunks=B)
audio = torch.randn(B * 480000, device=dev) * 0.1
offs = torch.arange(B, dtype=torch.int64, device=dev) * 480000
lens = torch.full((B,), 480000, dtype=torch.int32, device=dev)
feat = torch.empty(B, 128, 3000, device=dev); feat_t = torch.empty(B, 3002, 128, device=dev, dtype=torch.bfloat16)
prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
res = {}
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out
res["logmel_ms"], _ = timed(lambda: plan.chunks(audio, offs, lens, out=feat, out_t=feat_t))
res["encode_ms"], enc = timed(lambda: eng.encode_time_major(feat_t))
res["generate_ms"], out = timed(lambda: eng.generate(enc, prompt, tok, beam_size=1), n=2)
res["ms_per_decode_step"] = res["generate_ms"] / 224
res["batch_total_ms"] = res["logmel_ms"] + res["encode_ms"] + res["generate_ms"]
res["rtfx_30s_windows"] = B * 30.0 / (res["batch_total_ms"] / 1e3)
for which, name, nbytes in [(0, "cross_attn", B * 1500 * 2560 * 2), (1, "skinny_fc1", 5120 * 1280 * 2), (2, "skinny_dxd", 1280 * 1280 * 2)]:
    ms = eng.bench_kernel(which, B, 96)
    res[name] = {"us": ms * 1e3, "GBps": nbytes / ms / 1e6}
for parts, name in [(1, "embed"), (2, "ln"), (4, "gemm"), (8, "self"), (16, "cross"), (32, "logits"), (6, "ln+gemm"), (31, "layers"), (63, "layers+logits")]:
    res["step_" + name + "_ms"] = eng.bench_step(B, parts, 10)
print(json.dumps(res, indent=1))
