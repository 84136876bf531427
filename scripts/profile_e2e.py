"""Where the host time of model.transcribe(host_audio, batch_size=32) goes (bench.py's e2e leg): cProfile of one call."""
import cProfile, pstats, sys, time, io
import numpy as np, torch
sys.path.insert(0, ".")
import manual_whisper_b200 as mw
from manual_whisper_b200.config import model_dims
from bench import device_weights, MODEL, BATCH, HOUR_S
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0"); dims = model_dims(MODEL)
sd = device_weights(dims, dev, seed=1234)
audio_np, turns = mw.synthetic_speech(HOUR_S, seed=1)
pinned = torch.empty(len(audio_np), dtype=torch.float32, pin_memory=True); pinned.numpy()[:] = audio_np
pipe = mw.load_model(MODEL, "cuda", compute_type="float16", language="zh", asr_options={"beam_size": 1},
                     vad_model=mw.InjectedVad(turns), model=sd, max_batch=BATCH, streams_per_device=streams)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipe.transcribe(pinned.numpy(), batch_size=BATCH, language="zh")
    torch.cuda.synchronize(); print("transcribe s", time.perf_counter() - t0, flush=True)
pr = cProfile.Profile()
torch.cuda.synchronize(); t0 = time.perf_counter()
pr.enable(); pipe.transcribe(pinned.numpy(), batch_size=BATCH, language="zh"); torch.cuda.synchronize(); pr.disable()
print("profiled s", time.perf_counter() - t0)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:6000])
