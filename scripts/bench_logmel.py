#!/usr/bin/env python3
"""Secondary metric of BASELINE.json: log-mel GB/s (config 5 — throughput sweep 1 s .. 4 h, un-chunked semantics, 80 and
128 mels) plus the pipeline's chunked form, beside whisperx.audio.log_mel_spectrogram's math on the host cores
(oracle/logmel.py, torch.stft path).  Algorithmic bytes = 4N + 4*n_mels*(N//160)  (SURVEY.md §8d).
Prints one JSON object; run on a GPU box:  python scripts/bench_logmel.py [--gpu-only]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manual_whisper_b200 import audio as A

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
rows = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)
for n_mels in (80, 128):
    plan = A.get_plan(n_mels, dev)
    for label, secs in (("1s", 1), ("10s", 10), ("30s", 30), ("5min", 300), ("1h", 3600), ("4h", 14400)):
        n = secs * 16000
        x = (torch.randn(n, device=dev) * 0.1)
        nbytes = 4 * n + 4 * n_mels * (n // 160)
        for _ in range(3):
            out = plan.long(x)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = plan.long(x); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        row = {"n_mels": n_mels, "audio": label, "gpu_ms": ms, "gpu_GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
               "rtfx": secs / (ms / 1e3)}
        if "--gpu-only" not in sys.argv and secs <= 3600:
            from oracle.logmel import log_mel_spectrogram as cpu_logmel
            xc = x.cpu().numpy()
            cpu_logmel(xc[: min(n, 480000)], n_mels)
            t0 = time.perf_counter(); cpu_logmel(xc, n_mels); dt = time.perf_counter() - t0
            row.update(cpu_ms=dt * 1e3, cpu_GBps=nbytes / dt / 1e9, cpu_cores=os.cpu_count())
        rows.append(row)
        del x, out
# the pipeline's form: 128 windows of 30 s, per-window max
plan = A.LogMelPlan(128, 0, max_chunks=128)
B = 128
big = torch.randn(B * 480000, device=dev) * 0.1
o = torch.arange(B, dtype=torch.int64, device=dev) * 480000
l = torch.full((B,), 480000, dtype=torch.int32, device=dev)
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
for _ in range(3):
    plan.chunks(big, o, l, out=out)
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.chunks(big, o, l, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(json.dumps({"metric": "log-mel GB/s (algorithmic bytes)", "hbm_peak_GBps": peak, "l2": "flushed between timed iterations",
                  "chunked_128x30s": {"ms": ms, "us_per_window": ms * 1e3 / B, "GBps": B * 3.456e6 / ms / 1e6,
                                      "frac_of_hbm_peak": B * 3.456e6 / ms / 1e6 / peak},
                  "sweep": rows}, indent=1))
