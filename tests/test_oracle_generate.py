"""Logit rules, greedy and beam search of the oracle: HF timestamp processor masks (golden), scripted-logit
unit cases for the search book-keeping, and the committed small-model ids."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from manual_whisper_b200.config import custom_dims, scaled_tokens, special_tokens
from manual_whisper_b200.weights import random_init
from oracle.generate import GenOptions, apply_rules, expand_suppress, generate, max_new_tokens
from oracle.model import OracleWhisper


def test_special_token_tables():
    t2, t3 = special_tokens(51865), special_tokens(51866)
    assert (t2.eot, t2.sot, t2.translate, t2.transcribe, t2.no_timestamps, t2.timestamp_begin) == (50257, 50258, 50358, 50359, 50363, 50364)
    assert (t3.translate, t3.transcribe, t3.no_timestamps, t3.timestamp_begin) == (50359, 50360, 50364, 50365)
    assert t3.lang_id("zh") == 50260 and t3.lang_id("yue") == 50358
    with pytest.raises(ValueError):
        t2.lang_id("yue")
    assert {50258, 50358, 50359, 50360, 50361, 50362}.issubset(t2.suppress_ids) and 220 not in t2.suppress_ids
    assert t2.suppress_ids_begin == [220, 50257]


def test_timestamp_rules_match_hf_processor_golden():
    g = json.load(open(os.path.join(GOLDEN, "timestamp_rules.json")))
    tok = scaled_tokens(g["vocab"])
    assert tok.timestamp_begin == g["timestamp_begin"] and tok.eot == g["eot"]
    for name, case in g["cases"].items():
        scores = torch.tensor([case["seed_scores"]])
        out = apply_rules(scores, [case["history"]], tok, suppress=[tok.no_timestamps], suppress_begin=[],
                          with_timestamps=True, max_initial_timestamp_index=50)
        got = torch.isinf(out[0]).nonzero().flatten().tolist()
        assert got == case["masked_is_inf"], name


def test_suppress_expansion_and_begin_rule():
    tok = scaled_tokens(2048)
    sup = expand_suppress(tok, [-1, 77], with_timestamps=False)
    assert 77 in sup and tok.sot in sup and tok.no_timestamps not in sup
    assert tok.no_timestamps in expand_suppress(tok, [-1], with_timestamps=True)
    lg = torch.zeros(2, tok.vocab)
    out = apply_rules(lg, [[], [5]], tok, sup, tok.suppress_ids_begin, False, 50)
    assert torch.isinf(out[0, tok.eot]) and torch.isinf(out[0, tok.blank]) and not torch.isinf(out[1, tok.eot])
    assert torch.isinf(out[:, 77]).all()


def test_max_new_tokens():
    assert max_new_tokens(4, 448) == 224 and max_new_tokens(228, 448) == 220 and max_new_tokens(448, 448) == 0


class Scripted:
    """Fake engine: logits depend only on (row history length, last token) through a table."""

    def __init__(self, vocab, table):
        self.vocab, self.table = vocab, table
        class D: dec_layers = 1
        self.dims = D()

    def cross_kv(self, enc):
        return [None]

    def new_cache(self):
        return [None]

    def decode(self, tokens, pos0, cross, cache, cross_index=None):
        R, n = tokens.shape
        out = torch.full((R, n, self.vocab), -20.0)
        for r in range(R):
            for i in range(n):
                for t, v in self.table.get(int(tokens[r, i]), {}).items():
                    out[r, i, t] = v
        return out

    @staticmethod
    def reorder_cache(cache, parent):
        return cache


def test_greedy_stops_at_eot_and_excludes_it():
    tok = scaled_tokens(2048)
    P = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
    table = {tok.no_timestamps: {300: 5.0}, 300: {301: 5.0}, 301: {tok.eot: 5.0}}
    res = generate(Scripted(tok.vocab, table), torch.zeros(2, 1, 1), P, tok, GenOptions(beam_size=1, suppress_tokens=[]))
    assert [r.sequences_ids[0] for r in res] == [[300, 301], [300, 301]]
    assert res[0].scores[0] <= 0 and abs(res[0].scores[0] - res[1].scores[0]) < 1e-6


def test_greedy_tie_breaks_to_lowest_index_and_first_step_suppresses_eot():
    tok = scaled_tokens(2048)
    P = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
    table = {tok.no_timestamps: {tok.eot: 9.0, 500: 3.0, 400: 3.0}, 400: {tok.eot: 1.0}}
    res = generate(Scripted(tok.vocab, table), torch.zeros(1, 1, 1), P, tok, GenOptions(beam_size=1, suppress_tokens=[]))
    assert res[0].sequences_ids[0] == [400]


def test_beam_prefers_higher_mean_logprob_and_honours_patience():
    tok = scaled_tokens(2048)
    P = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
    # greedy path: 300 then weak continuation; alternative 310 has a strong continuation
    table = {tok.no_timestamps: {300: 2.0, 310: 1.8}, 300: {301: 0.0, 302: 0.0, 303: 0.0, 304: 0.0},
             310: {311: 6.0}, 311: {tok.eot: 8.0}, 301: {tok.eot: 8.0}, 302: {tok.eot: 8.0}, 303: {tok.eot: 8.0}, 304: {tok.eot: 8.0}}
    m = Scripted(tok.vocab, table)
    g = generate(m, torch.zeros(1, 1, 1), P, tok, GenOptions(beam_size=1, suppress_tokens=[]))
    b = generate(m, torch.zeros(1, 1, 1), P, tok, GenOptions(beam_size=3, suppress_tokens=[], num_hypotheses=3))
    assert g[0].sequences_ids[0][0] == 300
    assert b[0].sequences_ids[0] == [310, 311]
    assert b[0].scores == sorted(b[0].scores, reverse=True) and len(b[0].sequences_ids) == 3
    p2 = generate(m, torch.zeros(1, 1, 1), P, tok, GenOptions(beam_size=3, patience=2.0, suppress_tokens=[], num_hypotheses=6))
    assert len(p2[0].sequences_ids) >= 3


def test_small_model_search_matches_committed_golden(golden_small):
    dims = custom_dims("golden-small", 80, 128, 2, 2, 2, 512, 2048, n_audio_ctx=100, n_text_ctx=32)
    tok = scaled_tokens(2048)
    orc = OracleWhisper(dims, random_init(dims, seed=11, scheme="lively"))
    pr = [tok.sot, tok.sot + 1, tok.transcribe, tok.no_timestamps]
    with torch.no_grad():
        enc = orc.encode(torch.from_numpy(golden_small["mel"]))
        greedy = generate(orc, enc, pr, tok, GenOptions(beam_size=1, max_length=32))
        beam = generate(orc, enc, pr, tok, GenOptions(beam_size=5, max_length=32))
        beam_ts = generate(orc, enc, pr[:-1], tok, GenOptions(beam_size=5, max_length=32))
    assert [r.sequences_ids[0] for r in greedy] == golden_small["greedy"].tolist()
    assert [r.sequences_ids[0] for r in beam] == golden_small["beam"].tolist()
    np.testing.assert_allclose([r.scores[0] for r in beam], golden_small["beam_scores"], atol=1e-4)
    for r, ref in zip(beam_ts, golden_small["beam_ts"].tolist()):
        assert r.sequences_ids[0] == [t for t in ref if t >= 0]
        ts = [t for t in r.sequences_ids[0] if t >= tok.timestamp_begin]
        assert r.sequences_ids[0][0] >= tok.timestamp_begin and ts == sorted(ts)
