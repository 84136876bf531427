#!/usr/bin/env python3
"""CPU-only probe of init schemes for the token-identity bar (oracle vs oracle, no GPU).

For a model size and an init scheme it decodes N synthetic windows greedily with the fp32 oracle and with the
rounding-emulating oracles (bf16 = what the engine stores; fp16 = analysis) and reports
  * the fp32 oracle's top-2 margin statistics (nats): median, share of steps under 0.05,
  * how often the greedy token changes inside a window (the ids are not one constant token),
  * windows identical / first divergence step of each emulation against fp32.
Used to choose PEAKED_EMB_STD in manual_whisper_b200/weights.py; results in profiles/parity_probe_r2.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--scheme", default="peaked")
    ap.add_argument("--emb-std", type=float, default=None)
    ap.add_argument("--windows", type=int, default=16)
    ap.add_argument("--max-new", type=int, default=224)
    ap.add_argument("--emu", default="bf16,fp16")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from manual_whisper_b200.config import model_dims, special_tokens
    from manual_whisper_b200.weights import random_init
    from manual_whisper_b200.vad import synthetic_speech, merge_chunks
    from oracle.logmel import log_mel_spectrogram
    from oracle.model import OracleWhisper
    from oracle.generate import generate, GenOptions
    torch.set_grad_enabled(False)
    dims = model_dims(args.model)
    tok = special_tokens(dims.vocab)
    kw = {"emb_std": args.emb_std} if args.scheme == "peaked" else {}
    sd = random_init(dims, seed=args.seed, scheme=args.scheme, **kw)
    audio, turns = synthetic_speech(30.0 * args.windows * 1.3, seed=2)
    wins = merge_chunks(turns, 30)[: args.windows]
    mels = []
    for w in wins:
        a = audio[int(w["start"] * 16000): int(w["end"] * 16000)]
        mels.append(log_mel_spectrogram(a, dims.n_mels, padding=480000 - len(a)))
    mel = torch.stack(mels)
    prompt = [tok.sot, tok.lang_id("zh"), tok.transcribe, tok.no_timestamps]
    opt = GenOptions(beam_size=1, max_length=2 * args.max_new)
    out = {"model": args.model, "scheme": args.scheme, "emb_std": args.emb_std, "windows": len(wins), "max_new": args.max_new}
    runs = {}
    for mode in [False] + [m for m in args.emu.split(",") if m]:
        t0 = time.time()
        orc = OracleWhisper(dims, sd, emulate=mode)
        enc = orc.encode(mel)
        res, trace = generate(orc, enc, prompt, tok, opt, return_trace=True)
        runs[mode or "fp32"] = ([r.sequences_ids[0] for r in res], trace)
        print(f"[{mode or 'fp32'}] {time.time() - t0:.1f}s", flush=True)
    ids32, trace32 = runs["fp32"]
    margins = []
    for lg in trace32:
        top = lg.topk(2, dim=-1).values
        margins.append((top[:, 0] - top[:, 1]).numpy())
    margins = np.stack(margins)          # [steps, B]
    changes = [sum(1 for i in range(1, len(s)) if s[i] != s[i - 1]) for s in ids32]
    out["margin_nats"] = {"median": float(np.median(margins)), "p01": float(np.quantile(margins, 0.01)),
                          "min": float(margins.min()), "share_under_0.05": float((margins < 0.05).mean()),
                          "share_under_0.01": float((margins < 0.01).mean())}
    out["token_changes_per_window"] = {"mean": float(np.mean(changes)), "max": int(max(changes)), "windows_with_any": int(sum(c > 0 for c in changes))}
    out["unique_first_tokens"] = len({s[0] for s in ids32 if s})
    out["lengths"] = [len(s) for s in ids32]
    for mode, (ids, _) in runs.items():
        if mode == "fp32":
            continue
        ident, div = 0, []
        for b, (a, c) in enumerate(zip(ids, ids32)):
            if a == c:
                ident += 1
                continue
            k = next((i for i in range(min(len(a), len(c))) if a[i] != c[i]), min(len(a), len(c)))
            div.append({"window": b, "step": k, "oracle_margin": float(margins[k, b]) if k < margins.shape[0] else None})
        out[f"identical_{mode}_vs_fp32"] = {"identical": ident, "of": len(ids32), "divergences": div}
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
