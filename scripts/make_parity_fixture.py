#!/usr/bin/env python3
"""Generates the committed token-identity fixtures tests/golden/parity_<model>_<scheme>.npz (CPU only, oracle only).

For N VAD windows of the seeded synthetic recording it stores what the ORACLE decodes greedily on seeded random-init
weights (same values the engine gets): ids of the fp32 oracle, ids of the storage-rounding oracle (fp16, the engine's
rounding points), and the fp32 oracle's top-2 margin (nats) at every step.  The GPU parity tests and
scripts/gpu_parity_stats.py regenerate the same audio and weights from the seeds recorded here, run the CUDA engine and
compare - so the slow part (large-v3 in fp32 on host cores) never runs on the GPU box.

  python scripts/make_parity_fixture.py --model tiny --scheme peaked --windows 128
  python scripts/make_parity_fixture.py --model large-v3 --scheme peaked --windows 128      # ~1 h on 8 cores
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


def fixture_path(model, scheme, tag=""):
    return os.path.join(ROOT, "tests", "golden", f"parity_{model.replace('-', '_')}_{scheme}{tag}.npz")


def inputs(model, scheme, n_windows, weight_seed, audio_seed, emb_std=None, language="zh", with_timestamps=False):
    """Weights, audio and window list, regenerated identically by the GPU side."""
    from manual_whisper_b200.config import model_dims, special_tokens
    from manual_whisper_b200.weights import random_init
    from manual_whisper_b200.vad import synthetic_speech, merge_chunks
    dims = model_dims(model)
    tok = special_tokens(dims.vocab)
    kw = {"emb_std": emb_std} if (scheme == "peaked" and emb_std is not None) else {}
    sd = random_init(dims, seed=weight_seed, scheme=scheme, **kw)
    audio, turns = synthetic_speech(30.0 * n_windows * 1.25 + 30.0, seed=audio_seed)
    wins = merge_chunks(turns, 30)[:n_windows]
    assert len(wins) == n_windows, (len(wins), n_windows)
    offs = np.array([int(w["start"] * 16000) for w in wins], dtype=np.int64)
    lens = np.array([int(w["end"] * 16000) for w in wins], dtype=np.int64) - offs
    prompt = [tok.sot, tok.lang_id(language), tok.transcribe] + ([] if with_timestamps else [tok.no_timestamps])
    return dims, tok, sd, audio, wins, offs, lens, prompt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--scheme", default="peaked")
    ap.add_argument("--windows", type=int, default=128)
    ap.add_argument("--max-new", type=int, default=224)
    ap.add_argument("--weight-seed", type=int, default=1234)
    ap.add_argument("--audio-seed", type=int, default=2)
    ap.add_argument("--emb-std", type=float, default=None)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--beam", type=int, default=1, help="beam_size (BASELINE config 4: 5, patience 1, length_penalty 1)")
    ap.add_argument("--timestamps", action="store_true", help="without_timestamps=False: prompt [sot, lang, task], timestamp rules on")
    ap.add_argument("--language", default="zh")
    ap.add_argument("--suffix", default="", help="appended to the fixture name (a second fixture of the same model, e.g. another audio seed)")
    args = ap.parse_args()
    from oracle.logmel import log_mel_spectrogram
    from oracle.model import OracleWhisper
    from oracle.generate import generate, GenOptions
    torch.set_grad_enabled(False)
    dims, tok, sd, audio, wins, offs, lens, prompt = inputs(args.model, args.scheme, args.windows, args.weight_seed,
                                                            args.audio_seed, args.emb_std, args.language, args.timestamps)
    N, S = args.windows, args.max_new
    opt = GenOptions(beam_size=args.beam, max_length=2 * S)
    scores = {k: np.zeros(N, dtype=np.float32) for k in ("ids_fp32", "ids_emu")}
    out = {k: np.full((N, S), -1, dtype=np.int32) for k in ("ids_fp32", "ids_emu")}
    margins = np.full((N, S), np.nan, dtype=np.float32)
    t0 = time.time()
    for name, emu in (("ids_fp32", False), ("ids_emu", "fp16")):
        orc = OracleWhisper(dims, sd, emulate=emu)
        for b0 in range(0, N, args.batch):
            b1 = min(N, b0 + args.batch)
            mel = torch.stack([log_mel_spectrogram(audio[offs[i]: offs[i] + lens[i]], dims.n_mels, padding=480000 - int(lens[i]))
                               for i in range(b0, b1)])
            enc = orc.encode(mel)
            if args.beam > 1:
                res, trace = generate(orc, enc, prompt, tok, opt), []
            else:
                res, trace = generate(orc, enc, prompt, tok, opt, return_trace=True)
            for i, r in enumerate(res):
                ids = r.sequences_ids[0]
                out[name][b0 + i, : len(ids)] = ids
                scores[name][b0 + i] = r.scores[0]
            if not emu:
                for s, lg in enumerate(trace):
                    top = lg.topk(2, dim=-1).values
                    margins[b0: b1, s] = (top[:, 0] - top[:, 1]).numpy()
            print(f"[{name}] windows {b0}..{b1} done at {time.time() - t0:.0f}s", flush=True)
    meta = {"model": args.model, "scheme": args.scheme, "windows": N, "max_new": S, "weight_seed": args.weight_seed,
            "audio_seed": args.audio_seed, "emb_std": args.emb_std, "prompt": prompt, "emu": "fp16", "beam": args.beam,
            "with_timestamps": bool(args.timestamps), "language": args.language,
            "generator": "scripts/make_parity_fixture.py", "seconds": time.time() - t0}
    tag = (f"_beam{args.beam}" if args.beam > 1 else "") + ("_ts" if args.timestamps else "") + args.suffix
    path = fixture_path(args.model, args.scheme, tag)
    np.savez_compressed(path, ids_fp32=out["ids_fp32"], ids_emu=out["ids_emu"], margins=margins.astype(np.float16),
                        scores_fp32=scores["ids_fp32"], scores_emu=scores["ids_emu"],
                        offs=offs, lens=lens, meta=np.array(json.dumps(meta)))
    m = margins[~np.isnan(margins)]
    if m.size == 0:
        m = np.array([np.nan])
    same = int((out["ids_fp32"] == out["ids_emu"]).all(axis=1).sum())
    print(json.dumps({"path": path, "bytes": os.path.getsize(path), "margin_median": float(np.median(m)),
                      "share_under_0.05": float((m < 0.05).mean()), "emu_identical_to_fp32": same, "of": N,
                      "distinct_sequences": len({tuple(r) for r in out["ids_fp32"].tolist()}), **meta}, indent=1))


if __name__ == "__main__":
    main()
