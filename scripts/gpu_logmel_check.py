"""Scratch GPU check for the log-mel kernels: parity vs the oracle + event timing."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200 import audio as A, _lib
from oracle.logmel import log_mel_spectrogram as oracle_logmel, log_mel_chunks

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
res = {}
t = np.arange(480000) / 16000
cases = {"noise": (0.1 * rng.standard_normal(480000)).astype(np.float32),
         "chirp": (0.5 * np.sin(2 * np.pi * (200 + 50 * t) * t)).astype(np.float32),
         "zeros": np.zeros(480000, np.float32)}
for name, a in cases.items():
    for nm in (80, 128):
        got = A.log_mel_spectrogram(a, nm, device=dev).cpu()
        ref = oracle_logmel(a, nm)
        res[f"{name}_{nm}"] = float((got - ref).abs().max())
for n in (1, 399, 400, 16000, 479999, 480000):
    a = (0.1 * rng.standard_normal(n)).astype(np.float32)
    got = A.log_mel_spectrogram(a, 128, padding=480000 - n, device=dev).cpu()
    ref = oracle_logmel(a, 128, padding=480000 - n)
    res[f"len{n}"] = float((got - ref).abs().max())
# chunked API
N = 16000 * 300
audio = (0.1 * rng.standard_normal(N)).astype(np.float32)
offs = np.array([0, 123457, 1000001, 4000000, N - 5], dtype=np.int64)
lens = np.array([480000, 333333, 17, 480000, 5], dtype=np.int32)
plan = A.get_plan(128, dev)
d_audio = torch.from_numpy(audio).to(dev)
out_t = torch.empty((5, 3002, 128), dtype=torch.bfloat16, device=dev)
got = plan.chunks(d_audio, torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev), out_t=out_t)
ref = log_mel_chunks(audio, offs, lens, 128)
res["chunks"] = float((got.cpu() - ref).abs().max())
res["chunks_t"] = float((out_t[:, 1:3001].float().cpu() - ref.transpose(1, 2).bfloat16().float()).abs().max())
res["chunks_t_pad"] = float(out_t[:, [0, 3001]].float().abs().max())
# timing: 128 chunks of 30 s (1.92 MB in + 1.536 MB out each), inputs > L2
B = 128
big = torch.from_numpy((0.1 * rng.standard_normal(B * 480000)).astype(np.float32)).to(dev)
o = torch.arange(B, dtype=torch.int64, device=dev) * 480000
l = torch.full((B,), 480000, dtype=torch.int32, device=dev)
plan2 = A.LogMelPlan(128, 0, max_chunks=B)
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
for _ in range(3):
    plan2.chunks(big, o, l, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 10
for _ in range(iters):
    plan2.chunks(big, o, l, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
res["ms_per_128_chunks"] = ms
res["us_per_chunk"] = ms * 1e3 / B
res["algo_GBps"] = B * 3.456e6 / (ms * 1e-3) / 1e9
res["launches"] = _lib.launch_count()
print(json.dumps(res, indent=1))
