"""Host-side logic of transcribe() that needs no GPU: word grouping against outputs of the reference's own add_word_timestamps
(tests/golden/ref_words.json, written by tests/golden/make_word_golden.py), segment slicing / seek rule against the oracle's
restatement of whisper/transcribe.py:350-410, and the per-window records the ranks exchange in speculative seek mode."""
import copy
import json
import os

import numpy as np
import pytest

from oracle import transcribe as otr

HERE = os.path.dirname(os.path.abspath(__file__))


def test_word_grouping_matches_reference_golden():
    from whisper_b200.timing import WordTiming, add_word_timestamps
    with open(os.path.join(HERE, "golden", "ref_words.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 100
    for c in cases:
        al = [WordTiming(w["word"], list(w["tokens"]), w["start"], w["end"], w["probability"]) for w in c["alignment"]]
        segs = copy.deepcopy(c["segments"])
        add_word_timestamps(segs, al, 50257, last_speech_timestamp=c["last_speech_timestamp"])
        assert segs == c["expected"]


@pytest.mark.parametrize("seed", range(6))
def test_segment_slicing_and_seek_rule_match_oracle(seed):
    """random token streams with the timestamp patterns that matter (pairs, single ending, none, unfinished tail)"""
    from whisper_b200.decoding import DecodingResult
    from whisper_b200.transcribe import _next_seek, _segments_from_tokens
    rng = np.random.default_rng(seed)
    tb = 50364
    for _ in range(40):
        toks, t = [], int(rng.integers(0, 200))
        for _ in range(int(rng.integers(1, 6))):
            toks.append(tb + t)
            toks += rng.integers(0, 50257, int(rng.integers(1, 6))).tolist()
            t += int(rng.integers(1, 300))
            if rng.random() < 0.8:
                toks.append(tb + min(t, 1500))
        if rng.random() < 0.3:
            toks = toks[:-1] or toks
        seek, size = 6000, int(rng.choice([3000, 1000]))
        want_segs, want_seek = otr.window_schedule_step(toks, seek, size, tb, seek * 0.01)
        r = DecodingResult(tokens=toks)
        got = _segments_from_tokens(toks, r, seek * 0.01, size * 0.01, seek, tb, keep_tail=False)
        assert [(round(s["start"], 6), round(s["end"], 6), s["tokens"]) for s in got] == [(round(a, 6), round(b, 6), tk) for a, b, tk in want_segs]
        assert _next_seek(toks, seek, size, tb) == want_seek
        # fixed windows keep what the reference would decode again: the segments cover every text token exactly once
        kept = _segments_from_tokens(toks, r, seek * 0.01, size * 0.01, seek, tb, keep_tail=True)
        flat = [t for s in kept for t in s["tokens"]]
        assert flat == toks[:len(flat)] and all(x >= tb for x in toks[len(flat):])     # (only a lone opening timestamp may be left over)


def test_window_records_survive_the_exchange():
    from whisper_b200.decoding import DecodingResult
    from whisper_b200.timing import Alignment
    from whisper_b200.transcribe import _Window
    al = Alignment(np.arange(3), np.arange(3), np.zeros((2, 4), np.float32), np.array([0.25, 0.5], np.float32), np.array([0.0, 0.4, 1.2]))
    w = _Window(6000, 3000, DecodingResult(tokens=[50364, 11, 50400], avg_logprob=-0.3, no_speech_prob=0.1, sum_logprob=-0.9, steps=4,
                                          candidates=5, temperature=0.2), al, [11])
    d = json.loads(json.dumps(w.pack()))                     # plain containers only
    u = _Window.unpack(d)
    assert (u.seek, u.segment_size, u.text_tokens) == (6000, 3000, [11])
    assert u.result == w.result
    assert np.allclose(u.alignment.jump_times, al.jump_times) and np.allclose(u.alignment.text_token_probs, al.text_token_probs)
