"""BASELINE.json configs[3]: Whisper medium / small (80 mels), beam_size=5 with timestamp rules, batch_size=16 — sanity +
timing at real sizes, and parity against the oracle on `small` (first windows; skipped with --no-oracle)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import manual_whisper_b200 as mw
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.weights import random_init
res = {}
audio, turns = mw.synthetic_speech(64 * 28.0, seed=2)
wins = mw.merge_chunks(turns, 30)
for name in ("small", "medium"):
    dims = model_dims(name); tok = special_tokens(dims.vocab)
    sd = random_init(dims, seed=1234, scheme="lively")
    pipe = mw.load_model(name, "cuda", compute_type="float16", language="en", model=sd, max_batch=16, streams_per_device=2,
                         asr_options={"beam_size": 5, "patience": 1, "length_penalty": 1, "without_timestamps": False},
                         vad_model=mw.InjectedVad(turns))
    pipe.transcribe(audio[: 16000 * 120], batch_size=16)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = pipe.transcribe(audio, batch_size=16)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ids = [s["tokens"] for s in out["segments"]]
    ts_ok = all(i and i[0] >= tok.timestamp_begin and [t for t in i if t >= tok.timestamp_begin] == sorted(t for t in i if t >= tok.timestamp_begin) for i in ids)
    res[name] = {"windows": len(ids), "seconds": dt, "rtfx": len(audio) / 16000 / dt, "mean_len": float(np.mean([len(i) for i in ids])), "timestamp_rules_hold": ts_ok}
    if name == "small" and "--no-oracle" not in sys.argv:
        from oracle.logmel import log_mel_chunks
        from oracle.model import OracleWhisper
        from oracle.generate import generate, GenOptions
        n = 3
        offs = [int(w["start"] * 16000) for w in wins[:n]]; lens = [int(w["end"] * 16000) - o for w, o in zip(wins[:n], offs)]
        emu = OracleWhisper(dims, sd, emulate=True)
        with torch.no_grad():
            ref = generate(emu, emu.encode(log_mel_chunks(audio, offs, lens, 80)), [tok.sot, tok.lang_id("en"), tok.transcribe], tok,
                           GenOptions(beam_size=5))
        res["small_parity_first3"] = [ids[i] == ref[i].sequences_ids[0] for i in range(n)]
        res["small_common_prefix"] = [next((k for k, (a, b) in enumerate(zip(ids[i], ref[i].sequences_ids[0])) if a != b), min(len(ids[i]), len(ref[i].sequences_ids[0]))) for i in range(n)]
    del pipe
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
