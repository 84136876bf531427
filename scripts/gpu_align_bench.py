"""Forced-alignment throughput: XLSR-53-large wav2vec2-CTC (24 layers, d 1024) emissions + CTC
trellis/backtrack for B windows of 30 s, beside the fp32 torch CPU restatement (oracle/) on one window.
args: [B=16] [iters=3] [--no-cpu]      prints one JSON object."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manual_whisper_b200.w2v import W2vDims, random_init_w2v
from manual_whisper_b200.alignment import AlignEngine
from manual_whisper_b200 import synthetic_speech, _lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args else 16
iters = int(args[1]) if len(args) > 1 else 3
vocab = 3503                                    # the zh XLSR fine-tune's dictionary size
dims = W2vDims(vocab=vocab)
sd = random_init_w2v(dims, seed=1, std=0.02)
eng = AlignEngine(dims, sd, device_index=0, max_batch=B, max_samples=480000)
audio, _ = synthetic_speech(30.0 * B + 1, seed=3)
d_audio = torch.from_numpy(audio).cuda()
rng = np.random.default_rng(0)
lens = rng.integers(int(20 * 16000), 480001, size=B).astype(np.int32)       # 20-30 s windows, like merged VAD chunks
offs = (np.arange(B) * 480000).astype(np.int64)
tokens = [[int(x) for x in rng.integers(1, vocab, size=int(l / 16000 * 5))] for l in lens]   # ~5 characters per second
res = {"windows": B, "audio_s": float(lens.sum() / 16000), "workspace_GB": eng.workspace_bytes / 1e9}
em, frames = eng.emissions(d_audio, offs, lens)
eng.ctc_align(em, frames, tokens, 0)
torch.cuda.synchronize()
t_em, t_ctc = [], []
l0 = _lib.launch_count()
for _ in range(iters):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    em, frames = eng.emissions(d_audio, offs, lens)
    e1.record()
    ft, fs, ok = eng.ctc_align(em, frames, tokens, 0)
    e2.record()
    torch.cuda.synchronize()
    t_em.append(e0.elapsed_time(e1)); t_ctc.append(e1.elapsed_time(e2))
res.update(emissions_ms=float(np.median(t_em)), ctc_ms=float(np.median(t_ctc)), launches_per_call=(_lib.launch_count() - l0) // iters,
           aligned=int(ok.sum()))
res["rtfx"] = res["audio_s"] / ((res["emissions_ms"] + res["ctc_ms"]) / 1e3)
flops = sum(2.0 * f * (12 * dims.d_model ** 2 * dims.n_layers) + 4.0 * f * f * dims.d_model * dims.n_layers for f in frames.astype(float))
res["encoder_TFLOPs_per_s"] = flops / (res["emissions_ms"] / 1e3) / 1e12
if "--no-cpu" not in sys.argv:
    from oracle.wav2vec2 import OracleWav2Vec2
    from oracle import align as OA
    ora = OracleWav2Vec2(dims, sd)
    w = torch.from_numpy(audio[: int(lens[0])])
    t0 = time.perf_counter()
    with torch.no_grad():
        e = ora.emissions(w).numpy()
    t1 = time.perf_counter()
    path = OA.backtrack(OA.get_trellis(e, tokens[0], 0), e, tokens[0], 0)
    t2 = time.perf_counter()
    res["cpu"] = {"cores": os.cpu_count(), "threads": torch.get_num_threads(), "window_s": float(lens[0] / 16000),
                  "emissions_s": t1 - t0, "trellis_backtrack_s": t2 - t1, "rtfx": float(lens[0] / 16000) / (t2 - t0),
                  "kind": "port (fp32 torch wav2vec2 + numpy DP; whisperx runs the DP as a Python loop over frames too)"}
print(json.dumps(res, indent=1))
