"""ORACLE (test infrastructure, not product code) — logit rules, greedy and beam search.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Restates ``ctranslate2.models.Whisper.generate(features, prompts, beam_size, patience, length_penalty,
max_length, suppress_blank, suppress_tokens)`` as called by whisperx ``generate_segment_batched``
(SURVEY.md A.6, A.8), reached from /root/reference/transcribe.py:123.  CTranslate2 is absent; the
timestamp rules follow the in-container restatement transformers/generation/logits_process.py:1996-2043
(pinned by tests/test_oracle_generate.py).  Beam-search book-keeping is from memory of CT2's
decoding.cc and is THE DEFINITION the CUDA engine must match ("parity unpinned" for beam order).

Definition pinned here (per chunk, beam k, patience p, length penalty a):
  * step 0 expands a single live hypothesis; afterwards k live beams.
  * every step: logprob = log_softmax(masked logits); candidate = parent cumulative + logprob;
    take the top 2k candidates over k*V by (value desc, flat index asc).
  * walk them in order until k live successors are chosen; an <eot> candidate met on the way is
    recorded as finished (score = cum / length^a, length counts the eot) while fewer than
    round(k*p) are finished.
  * a chunk stops when round(k*p) hypotheses are finished or max_new tokens were produced (then the
    live beams are recorded, best first, score = cum / length^a, until round(k*p) are held).
  * result = finished sorted by score desc (stable); hypothesis 0 is what whisperx keeps.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

NEG_INF = float("-inf")


@dataclass
class GenOptions:
    beam_size: int = 5
    patience: float = 1.0
    length_penalty: float = 1.0
    max_length: int = 448
    suppress_blank: bool = True
    suppress_tokens: Optional[List[int]] = field(default_factory=lambda: [-1])
    max_initial_timestamp_index: int = 50
    num_hypotheses: int = 1


@dataclass
class GenResult:
    sequences_ids: List[List[int]]
    scores: List[float]


def expand_suppress(tok, suppress_tokens, with_timestamps: bool) -> List[int]:
    """``suppress_tokens=[-1]`` -> the model's non-speech list + control tokens."""
    ids = set()
    for t in suppress_tokens or []:
        if t == -1:
            ids.update(tok.suppress_ids)
        elif t >= 0:
            ids.add(int(t))
    if with_timestamps:
        ids.add(tok.no_timestamps)
    return sorted(i for i in ids if i < tok.vocab)


def apply_rules(logits: torch.Tensor, generated: List[List[int]], tok, suppress: List[int], suppress_begin: List[int],
                with_timestamps: bool, max_initial_timestamp_index: int) -> torch.Tensor:
    """Masks logits [R, V] in place-free fashion.  `generated[r]` = ids produced so far by row r."""
    out = logits.clone()
    if suppress:
        out[:, suppress] = NEG_INF
    for r, seq in enumerate(generated):
        if len(seq) == 0 and suppress_begin:
            out[r, suppress_begin] = NEG_INF
    if not with_timestamps:
        return out
    tb = tok.timestamp_begin
    for r, seq in enumerate(generated):
        last_ts = len(seq) >= 1 and seq[-1] >= tb
        pen_ts = len(seq) < 2 or seq[-2] >= tb
        if last_ts:
            if pen_ts:
                out[r, tb:] = NEG_INF
            else:
                out[r, : tok.eot] = NEG_INF
        ts = [t for t in seq if t >= tb]
        if ts:
            last = ts[-1] if (last_ts and not pen_ts) else ts[-1] + 1
            out[r, tb:last] = NEG_INF
        if len(seq) == 0:
            out[r, :tb] = NEG_INF
            if max_initial_timestamp_index is not None and max_initial_timestamp_index >= 0:
                out[r, tb + max_initial_timestamp_index + 1:] = NEG_INF
    lp = torch.log_softmax(out.float(), dim=-1)
    for r in range(out.shape[0]):
        if torch.logsumexp(lp[r, tb:], dim=-1) > lp[r, :tb].max():
            out[r, :tb] = NEG_INF
    return out


def max_new_tokens(prompt_len: int, max_length: int) -> int:
    return max(0, min(max_length // 2, max_length - prompt_len))


def _argmax_lowest(v: torch.Tensor) -> int:
    """argmax with lowest-index tie-break (what the CUDA argmax implements)."""
    m = v.max()
    return int(torch.nonzero(v == m)[0, 0])


def generate(model, enc: torch.Tensor, prompt: List[int], tok, opt: GenOptions,
             return_trace: bool = False):
    """enc [B, 1500, d]; one shared prompt (whisperx builds ``[prompt]*B``).  Returns list[GenResult]
    (and, with return_trace, per-step fp32 masked logits of the greedy path for margin analysis)."""
    B = enc.shape[0]
    P = len(prompt)
    # timestamp rules are off iff <|notimestamps|> follows <|startoftranscript|> in the prompt (get_prompt appends `prefix`
    # tokens AFTER it, so looking only at the last prompt token would switch the rules on for prefixed prompts)
    tail = prompt[prompt.index(tok.sot):] if tok.sot in prompt else prompt
    with_ts = tok.no_timestamps not in tail
    suppress = expand_suppress(tok, opt.suppress_tokens, with_ts)
    sup_begin = list(tok.suppress_ids_begin) if opt.suppress_blank else []
    n_new = max_new_tokens(P, opt.max_length)
    cross = model.cross_kv(enc)
    if opt.beam_size <= 1:
        return _greedy(model, cross, B, prompt, tok, opt, suppress, sup_begin, with_ts, n_new, return_trace)
    return _beam(model, cross, B, prompt, tok, opt, suppress, sup_begin, with_ts, n_new)


def _greedy(model, cross, B, prompt, tok, opt, suppress, sup_begin, with_ts, n_new, return_trace):
    P = len(prompt)
    cache = model.new_cache()
    ptoks = torch.tensor([prompt] * B, dtype=torch.long)
    if P > 1:
        model.decode(ptoks[:, :-1], 0, cross, cache)
    cur = ptoks[:, -1:]
    gen: List[List[int]] = [[] for _ in range(B)]
    done = [False] * B
    cum = [0.0] * B
    trace = []
    for step in range(n_new):
        logits = model.decode(cur, P - 1 + step, cross, cache)[:, 0]
        masked = apply_rules(logits, gen, tok, suppress, sup_begin, with_ts, opt.max_initial_timestamp_index)
        if return_trace:
            trace.append(masked.clone())
        lp = torch.log_softmax(masked.float(), dim=-1)
        nxt = []
        for r in range(B):
            t = _argmax_lowest(masked[r])
            if done[r]:
                t = tok.eot
            else:
                cum[r] += float(lp[r, t])
                if t == tok.eot:
                    done[r] = True
                else:
                    gen[r].append(t)
            nxt.append(t)
        if all(done):
            break
        cur = torch.tensor(nxt, dtype=torch.long)[:, None]
    res = []
    for r in range(B):
        length = len(gen[r]) + (1 if done[r] else 0)
        score = cum[r] / (max(length, 1) ** opt.length_penalty) if opt.length_penalty != 0 else cum[r]
        res.append(GenResult([gen[r]], [score]))
    return (res, trace) if return_trace else res


def _beam(model, cross, B, prompt, tok, opt, suppress, sup_begin, with_ts, n_new):
    k = opt.beam_size
    V = tok.vocab
    max_fin = max(1, int(round(k * opt.patience)))
    P = len(prompt)
    R = B * k
    cache = model.new_cache()
    cross_index = torch.arange(B).repeat_interleave(k)
    ptoks = torch.tensor([prompt] * R, dtype=torch.long)
    if P > 1:
        model.decode(ptoks[:, :-1], 0, cross, cache, cross_index)
    cur = ptoks[:, -1:]
    gen: List[List[int]] = [[] for _ in range(R)]
    cum = torch.zeros(B, k)
    cum[:, 1:] = NEG_INF  # step 0: one live hypothesis per chunk
    finished: List[List] = [[] for _ in range(B)]  # (score, ids)
    active = [True] * B

    def norm(c, length):
        return c / (max(length, 1) ** opt.length_penalty) if opt.length_penalty != 0 else c

    for step in range(n_new):
        logits = model.decode(cur, P - 1 + step, cross, cache, cross_index)[:, 0]
        masked = apply_rules(logits, gen, tok, suppress, sup_begin, with_ts, opt.max_initial_timestamp_index)
        lp = torch.log_softmax(masked.float(), dim=-1).view(B, k, V)
        cand = (cum[:, :, None] + lp).view(B, k * V)
        # top 2k by (value desc, flat index asc): stable sort of the negated values
        order = torch.sort(-cand, dim=1, stable=True).indices[:, : 2 * k]
        parent = torch.arange(R)
        nxt = torch.full((R,), tok.eot, dtype=torch.long)
        new_gen = [list(g) for g in gen]
        new_cum = torch.full((B, k), NEG_INF)
        for b in range(B):
            if not active[b]:
                for j in range(k):
                    new_gen[b * k + j] = gen[b * k + j]
                continue
            live = 0
            for c in order[b].tolist():
                val = float(cand[b, c])
                if val == NEG_INF:
                    break
                pj, t = divmod(c, V)
                src = gen[b * k + pj]
                if t == tok.eot:
                    if len(finished[b]) < max_fin:
                        finished[b].append((norm(val, len(src) + 1), list(src)))
                    continue
                row = b * k + live
                parent[row] = b * k + pj
                nxt[row] = t
                new_gen[row] = src + [t]
                new_cum[b, live] = val
                live += 1
                if live == k:
                    break
            if len(finished[b]) >= max_fin or live == 0:
                active[b] = False
        gen = new_gen
        cum = new_cum
        if not any(active):
            break
        cache = model.reorder_cache(cache, parent)
        cur = nxt[:, None]
    res = []
    for b in range(B):
        if active[b] or not finished[b]:
            lives = [(float(cum[b, j]), gen[b * k + j]) for j in range(k) if float(cum[b, j]) != NEG_INF]
            for c, ids in lives:
                if len(finished[b]) < max_fin:
                    finished[b].append((norm(c, len(ids)), list(ids)))
        fin = sorted(enumerate(finished[b]), key=lambda e: (-e[1][0], e[0]))
        fin = [e[1] for e in fin][: max(1, opt.num_hypotheses)]
        res.append(GenResult([ids for _, ids in fin], [s for s, _ in fin]))
    return res
