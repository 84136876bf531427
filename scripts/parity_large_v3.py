#!/usr/bin/env python3
"""Parity at scale (usage: parity_large_v3.py N_WINDOWS MAX_LENGTH SCHEME [MODEL]).
Flagship-size parity (BASELINE.json configs[1] generalised to a few windows): Whisper large-v3, the SAME seeded
random-init weights (rounded to bf16 once) on both sides, greedy decoding of VAD windows of the synthetic recording:
   * log-mel: max abs error vs the oracle (tolerance 1e-4)
   * encoder output: relative L2 vs the fp32 oracle
   * greedy ids vs the fp32 oracle and vs the bf16-rounding oracle, first divergence + oracle top-2 margin there
Writes one JSON object.  CPU-heavy (the oracle runs large-v3 in fp32 on the host): N windows x ~10 s on 16 cores."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manual_whisper_b200 as mw
from manual_whisper_b200.config import model_dims, special_tokens
from manual_whisper_b200.weights import random_init
from oracle.logmel import log_mel_chunks
from oracle.model import OracleWhisper
from oracle.generate import generate, GenOptions

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MAXLEN = int(sys.argv[2]) if len(sys.argv) > 2 else 448
scheme = sys.argv[3] if len(sys.argv) > 3 else "survey"
MODEL = sys.argv[4] if len(sys.argv) > 4 else "large-v3"
dims = model_dims(MODEL); tok = special_tokens(dims.vocab)
t0 = time.time(); sd = random_init(dims, seed=1234, scheme=scheme); t_init = time.time() - t0
audio, turns = mw.synthetic_speech(N * 30.0 + 5, seed=1)
wins = mw.merge_chunks(turns, 30)[:N]
offs = [int(w["start"] * 16000) for w in wins]; lens = [int(w["end"] * 16000) - o for w, o in zip(wins, offs)]
pipe = mw.load_model(MODEL, "cuda", compute_type="float16", language="zh", asr_options={"beam_size": 1}, model=sd,
                     vad_model=mw.InjectedVad([(w["start"], w["end"]) for w in wins]), max_batch=max(N, 1), streams_per_device=1)
model = pipe.model
model.max_length = MAXLEN
res = pipe.transcribe(audio, batch_size=N)
got = [s["tokens"] for s in res["segments"]]
d_audio = torch.from_numpy(audio).cuda()
mel_gpu = model.plan.chunks(d_audio, torch.tensor(offs).cuda(), torch.tensor(lens, dtype=torch.int32).cuda()).cpu()
mel = log_mel_chunks(audio, offs, lens, dims.n_mels)
enc_gpu = model.engine.encode(mel_gpu.cuda()).float().cpu()
out = {"model": MODEL, "windows": N, "max_length": MAXLEN, "init_scheme": scheme, "weights_init_s": t_init,
       "logmel_max_abs_err": float((mel_gpu - mel).abs().max())}
prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
torch.set_num_threads(os.cpu_count() or 8)
for name, emu in (("fp32", False), ("bf16_rounding", True)):
    orc = OracleWhisper(dims, sd, emulate=emu)
    t0 = time.time()
    with torch.no_grad():
        enc = orc.encode(mel)
        if not emu:
            out["encoder_rel_l2_vs_fp32"] = float((enc_gpu - enc).norm() / enc.norm())
        ref, trace = generate(orc, enc, prompt, tok, GenOptions(beam_size=1, max_length=MAXLEN), return_trace=True)
    rows = []
    for b in range(N):
        a, c = got[b], ref[b].sequences_ids[0]
        k = next((i for i in range(min(len(a), len(c))) if a[i] != c[i]), None)
        if k is None and len(a) == len(c):
            rows.append({"window": b, "identical": True, "len": len(a)})
        else:
            k = min(len(a), len(c)) if k is None else k
            top = trace[k][b].topk(2).values
            rows.append({"window": b, "identical": False, "first_divergence": k, "oracle_top2_margin": float(top[0] - top[1]), "len": len(a)})
    # teacher-forced agreement: feed the ORACLE's ids to the engine and compare the per-step argmax after the logit rules
    from oracle.generate import apply_rules, expand_suppress
    sup = expand_suppress(tok, [-1], False)
    agree = total = 0
    worst = 0.0
    forced = np.array([prompt + ref[b].sequences_ids[0][:-1] for b in range(N)], dtype=np.int32)
    lg = model.engine.decoder_logits(model.engine.encode(mel_gpu.cuda()), forced).cpu()
    for b in range(N):
        ids = ref[b].sequences_ids[0]
        for i, t in enumerate(ids):
            row = apply_rules(lg[b, len(prompt) - 1 + i][None], [ids[:i]], tok, sup, tok.suppress_ids_begin, False, 50)[0]
            total += 1
            if int(row.argmax()) == t:
                agree += 1
            else:
                top = trace[i][b].topk(2).values
                worst = max(worst, float(top[0] - top[1]))
    out[f"teacher_forced_vs_{name}"] = {"steps": total, "argmax_agree": agree, "fraction": agree / max(total, 1),
                                      "largest_oracle_margin_among_disagreements": worst}
    div = [r for r in rows if not r["identical"]]
    out[f"greedy_vs_{name}"] = {"identical": sum(r["identical"] for r in rows), "of": N,
                               "mean_common_prefix": float(np.mean([r["len"] if r["identical"] else r["first_divergence"] for r in rows])),
                               "largest_margin_at_first_divergence": max([r["oracle_top2_margin"] for r in div], default=0.0),
                               "rows": rows if N <= 8 else div[:16], "oracle_seconds": time.time() - t0,
                               "unique_ids_window0": len(set(ref[0].sequences_ids[0]))}
print(json.dumps(out, indent=1))
