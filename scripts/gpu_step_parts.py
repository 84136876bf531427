"""Decode-step anatomy by kernel class (mw_bench_step parts mask), large-v3.  args: B"""
import sys, json
import torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims
from manual_whisper_b200.engine import Engine
from bench import device_weights
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0"); dims = model_dims("large-v3")
eng = Engine(dims, device_weights(dims, dev, 1234), 0, max_batch=B)
res = {}
for parts, name in [(1, "embed"), (2, "ln"), (4, "gemm"), (8, "self"), (16, "cross"), (32, "logits"), (63, "layers+logits")]:
    res[name] = round(eng.bench_step(B, parts, 20), 4)
print(json.dumps(res))
