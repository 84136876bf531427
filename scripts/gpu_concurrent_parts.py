"""Aggregate throughput of one kernel class of the decode step when S replicas (shared weights, one stream + host thread each)
replay it concurrently - which class stops scaling with batches in flight?  args: S [iters]"""
import sys, json, threading
import torch
sys.path.insert(0, ".")
import manual_whisper_b200 as mw
from manual_whisper_b200.config import model_dims
from bench import device_weights, MODEL, BATCH
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0"); dims = model_dims(MODEL)
pipe = mw.load_model(MODEL, "cuda", compute_type="float16", language="zh", asr_options={"beam_size": 1},
                     vad_model=mw.InjectedVad([]), model=device_weights(dims, dev, seed=1234), max_batch=BATCH, streams_per_device=S)
GB = {"gemm": 14 * dims.d_model ** 2 * 2 * dims.dec_layers / 1e9, "cross": BATCH * 1500 * 2 * dims.d_model * 2 * dims.dec_layers / 1e9,
      "ln": 3 * BATCH * dims.d_model * 6 * dims.dec_layers / 1e9, "layers+logits": None}
GB["layers+logits"] = GB["gemm"] + GB["cross"] + dims.vocab * dims.d_model * 2 / 1e9
res = {}
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for parts, name in [(4, "gemm"), (16, "cross"), (2, "ln"), (63, "layers+logits")]:
    if only and name not in only:
        continue
    for n in sorted({1, 2, 4, S}):
        out = [0.0] * n
        def work(i):
            rep = pipe.replicas[i]
            s = rep.stream or torch.cuda.current_stream(rep.device)
            with torch.cuda.device(rep.device), torch.cuda.stream(s):
                out[i] = rep.engine.bench_step(BATCH, parts, iters)
        th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        [t.start() for t in th]; [t.join() for t in th]
        ms = max(out)
        res[f"{name}_x{n}"] = {"ms_per_step_each": round(ms, 3), "ms_per_batch_step_aggregate": round(ms / n, 3),
                               "TBps": round(GB[name] * n / ms, 2)}
print(json.dumps(res, indent=1))
