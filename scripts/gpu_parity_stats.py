#!/usr/bin/env python3
"""Token-identity statistics of the CUDA engine against the committed oracle fixtures (tests/golden/parity_*.npz, made by
scripts/make_parity_fixture.py): the north-star bar "greedy ids identical on >= 99 % of chunks, divergences logged".

  python scripts/gpu_parity_stats.py --model tiny [--scheme peaked] [--windows N] [--out profiles/parity_tiny_peaked_r2.json]

The audio, windows and weights are regenerated from the seeds stored in the fixture (checked against its window table);
the engine decodes them through the public pipeline in batches of 32 and the ids are compared with the fp32 oracle's and
with the storage-rounding oracle's.  Also reports the SURVEY's forced-<eot> lengths U(40,120) (seed 3): identity of the
first L tokens of every window.  Nothing under oracle/ is executed here."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.make_parity_fixture import fixture_path, inputs      # noqa: E402


def engine_ids(meta, n_windows, batch=32, pipe=None):
    """ids the engine decodes for the first n_windows windows of the fixture's recording -> (ids [n, max_new] padded -1, pipe)"""
    import manual_whisper_b200 as mw
    beam, with_ts, language = meta.get("beam", 1), meta.get("with_timestamps", False), meta.get("language", "zh")
    dims, tok, sd, audio, wins, offs, lens, prompt = inputs(meta["model"], meta["scheme"], meta["windows"], meta["weight_seed"],
                                                            meta["audio_seed"], meta.get("emb_std"), language, with_ts)
    assert prompt == meta["prompt"]
    if pipe is None:
        pipe = mw.load_model(meta["model"], "cuda", compute_type="float16", language=language,
                             asr_options={"beam_size": beam, "without_timestamps": not with_ts},
                             vad_model=mw.InjectedVad([]), model=sd, max_batch=batch, streams_per_device=1)
    resident = pipe.upload(audio, offs[:n_windows], lens[:n_windows])
    res = pipe.run_device_batches(resident, offs[:n_windows], lens[:n_windows].astype(np.int32), batch)
    out = np.full((n_windows, meta["max_new"]), -1, dtype=np.int32)
    for i, (_, toks) in enumerate(res):
        out[i, : len(toks)] = toks
    return out, offs, lens, pipe


def compare(got, ref, margins, forced_len=None):
    """-> dict(identical, of, divergences[{window, step, oracle_margin}])"""
    n = got.shape[0]
    div = []
    for w in range(n):
        a, b = got[w], ref[w]
        if forced_len is not None:
            a, b = a[: forced_len[w]], b[: forced_len[w]]
        if np.array_equal(a, b):
            continue
        k = int(np.nonzero(a != b)[0][0])
        div.append({"window": w, "step": k, "oracle_margin": float(margins[w, k]) if k < margins.shape[1] else float("nan")})
    return {"identical": n - len(div), "of": n, "fraction": (n - len(div)) / n, "divergences": div}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--scheme", default="peaked")
    ap.add_argument("--windows", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--tag", default="", help="fixture suffix, e.g. _beam5_ts")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    fx = np.load(fixture_path(args.model, args.scheme, args.tag))
    meta = json.loads(str(fx["meta"]))
    n = args.windows or meta["windows"]
    got, offs, lens, pipe = engine_ids(meta, n, batch=args.batch)
    assert np.array_equal(offs[:n], fx["offs"][:n]) and np.array_equal(lens[:n], fx["lens"][:n]), "window table differs from the fixture"
    margins = fx["margins"].astype(np.float32)[:n]
    valid = margins[~np.isnan(margins)]
    if valid.size == 0:
        valid = np.array([np.nan], dtype=np.float32)
    from manual_whisper_b200 import _lib
    out = {"model": args.model, "scheme": args.scheme, "windows": n, "max_new": meta["max_new"], "beam": meta.get("beam", 1),
           "with_timestamps": meta.get("with_timestamps", False), "batch": args.batch,
           "engine_storage": "bf16" if _lib.load().mw_storage_dtype() == 1 else "fp16",
           "oracle_margin_nats": {"median": float(np.median(valid)), "share_under_0.05": float((valid < 0.05).mean()),
                                  "share_under_0.01": float((valid < 0.01).mean())},
           "distinct_sequences": len({tuple(r) for r in fx["ids_fp32"][:n].tolist()}),
           "token_changes_per_window_mean": float(np.mean([(np.diff(r[r >= 0]) != 0).sum() for r in fx["ids_fp32"][:n]])),
           "vs_fp32_oracle": compare(got, fx["ids_fp32"][:n], margins),
           "vs_rounding_oracle": compare(got, fx["ids_emu"][:n], margins)}
    forced = np.random.default_rng(3).integers(40, 121, size=n)
    out["forced_eot_U40_120_vs_fp32_oracle"] = compare(got, fx["ids_fp32"][:n], margins, forced)
    out["forced_eot_U40_120_vs_rounding_oracle"] = compare(got, fx["ids_emu"][:n], margins, forced)
    print(json.dumps(out, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
