"""Skinny GEMM time vs row count (greedy R=32 ... beam-5 R=160), large-v3 shapes, through mw_bench_kernel."""
import sys, json
import torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims
from manual_whisper_b200.engine import Engine
from bench import device_weights
dev = torch.device("cuda:0"); dims = model_dims("large-v3")
eng = Engine(dims, device_weights(dims, dev, 1234), 0, max_batch=32, max_beam=5)
rows = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 64, 96, 160]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = {}
for R in rows:
    out[R] = {"fc1_us": eng.bench_kernel(1, R, iters) * 1e3, "dxd_us": eng.bench_kernel(2, R, iters) * 1e3}
print(json.dumps(out, indent=1))
