"""Scratch GPU check: decoder logits / greedy / beam vs the oracle, then large-v3 decode timing."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from manual_whisper_b200.config import model_dims, custom_dims, special_tokens, scaled_tokens
from manual_whisper_b200.weights import random_init, _keys
from manual_whisper_b200.engine import Engine
from manual_whisper_b200 import _lib
from oracle.model import OracleWhisper
from oracle.generate import generate, GenOptions
import faulthandler; faulthandler.dump_traceback_later(50, exit=True)
dev = torch.device("cuda:0"); torch.manual_seed(0)
res = {}
def check(dims, tok, B, scheme, beam_cases=(1,)):
    out = {}
    sd = random_init(dims, scheme=scheme)
    eng = Engine(dims, sd, 0, max_batch=B, max_beam=5)
    mel = (torch.randn(B, dims.n_mels, 2 * dims.n_audio_ctx) * 0.5).clamp(-1.5, 1.5)
    enc = eng.encode(mel.to(dev))
    orc = OracleWhisper(dims, sd); emu = OracleWhisper(dims, sd, emulate_bf16=True)
    enc_cpu = enc.float().cpu()     # feed the SAME encoder output to the oracle decoder
    rng = np.random.default_rng(0)
    n = 12
    toks = rng.integers(0, dims.vocab, size=(B, n)).astype(np.int32)
    print('logits...', flush=True); got = eng.decoder_logits(enc, toks).cpu(); print('logits ok', flush=True)
    with torch.no_grad():
        for name, o in (("fp32", orc), ("emu", emu)):
            ref = o.decode(torch.from_numpy(toks).long(), 0, o.cross_kv(enc_cpu), o.new_cache())
            out[f"logits_maxabs_vs_{name}"] = (got - ref).abs().max().item()
        out["logits_absmax"] = ref.abs().max().item()
        prompt = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
        for beam in beam_cases:
            for with_ts in (False, True):
                p = prompt if not with_ts else prompt[:-1]
                print('gen', beam, with_ts, flush=True); g = eng.generate(enc, p, tok, beam_size=beam, max_length=dims.n_text_ctx); print('gen ok', flush=True)
                for name, o in (("emu", emu), ("fp32", orc)):
                    r = generate(o, enc_cpu, p, tok, GenOptions(beam_size=beam, max_length=dims.n_text_ctx))
                    same = [g[b].sequences_ids[0] == r[b].sequences_ids[0] for b in range(B)]
                    first_div = []
                    for b in range(B):
                        a, c = g[b].sequences_ids[0], r[b].sequences_ids[0]
                        k = next((i for i in range(min(len(a), len(c))) if a[i] != c[i]), min(len(a), len(c)))
                        first_div.append(k)
                    out[f"beam{beam}_ts{int(with_ts)}_vs_{name}"] = {"identical": sum(same), "of": B, "first_div": first_div,
                        "len": [len(x.sequences_ids[0]) for x in g], "score_diff": max(abs(g[b].scores[0] - r[b].scores[0]) for b in range(B))}
                out[f"beam{beam}_ts{int(with_ts)}_sample"] = g[0].sequences_ids[0][:10]
    return out
small = custom_dims("test-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=200, n_text_ctx=64)
res["small_lively"] = check(small, scaled_tokens(2048), 3, "lively", beam_cases=(1, 5))
res["small_survey"] = check(small, scaled_tokens(2048), 3, "survey", beam_cases=(1, 5))
print(json.dumps(res), flush=True)
t = model_dims("tiny")
res["tiny_lively"] = check(t, special_tokens(t.vocab), 2, "lively", beam_cases=(1,))
print(json.dumps(res["tiny_lively"]), flush=True)
# large-v3 decode timing
dims = model_dims("large-v3"); tok = special_tokens(dims.vocab)
g = torch.Generator(device=dev); g.manual_seed(1)
sd = {}
for name, shape, kind in _keys(dims):
    if kind == "g": sd[name] = torch.ones(shape, device=dev)
    elif kind == "beta": sd[name] = torch.zeros(shape, device=dev)
    else: sd[name] = (torch.randn(shape, device=dev, generator=g) * 0.02)
sd["model.encoder.embed_positions.weight"] = torch.zeros(1500, dims.d_model, device=dev)
B = 32
eng = Engine(dims, sd, 0, max_batch=B); del sd
mel = torch.randn(B, 128, 3000, device=dev) * 0.5
enc = eng.encode(mel)
prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
eng.generate(enc, prompt, tok, beam_size=1, max_length=40)   # warm-up + graph capture
torch.cuda.synchronize(); l0 = _lib.launch_count(); t0 = time.time()
out = eng.generate(enc, prompt, tok, beam_size=1)
torch.cuda.synchronize(); dt = time.time() - t0
res["large_v3_decode_B32_s"] = dt
res["large_v3_decode_ms_per_step"] = dt / 224 * 1e3
res["large_v3_lens"] = [len(o.sequences_ids[0]) for o in out][:4]
res["launches"] = _lib.launch_count() - l0
res["workspace_GB"] = eng.workspace_bytes / 1e9
print(json.dumps({k: v for k, v in res.items() if k.startswith("large") or k in ("launches", "workspace_GB")}))
