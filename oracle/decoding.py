"""Oracle: host decode loop, logit filters, greedy and beam search (fp32, CPU).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates whisper/decoding.py of the reference
(SuppressBlank :450-457, SuppressTokens :460-465, ApplyTimestampRules :468-532,
GreedyDecoder :299-325, BeamSearchDecoder :328-431, MaximumLikelihoodRanker :217-240,
DecodingTask._main_loop :707-737 and .run :740-816) without the tokenizer text handling.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .model import OracleModel

_ASSET = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                      "whisper.coreml_b200", "assets", "tokenizer_specials.json")


@dataclass(frozen=True)
class Specials:
    """The token ids the decode loop needs (dumped from the reference tokenizer by
    tests/golden/make_golden.py; whisper/tokenizer.py:147-275, :330-363)."""
    n_vocab: int
    sot: int
    eot: int
    sot_sequence: Tuple[int, ...]
    no_timestamps: int
    timestamp_begin: int
    no_speech: int
    blank: Tuple[int, ...]           # tokenizer.encode(" ")
    suppress: Tuple[int, ...]        # DecodingTask._get_suppress_tokens() for suppress_tokens="-1"

    @staticmethod
    def load(n_vocab: int) -> "Specials":
        with open(_ASSET) as f:
            d = json.load(f)[str(n_vocab)]
        return Specials(n_vocab, d["sot"], d["eot"], tuple(d["sot_sequence"]), d["no_timestamps"],
                        d["timestamp_begin"], d["no_speech"], tuple(d["blank"]), tuple(d["suppress"]))


@dataclass
class Options:
    """Subset of whisper/decoding.py:81-115 DecodingOptions that reaches the hot loop."""
    beam_size: Optional[int] = None
    sample_len: Optional[int] = None
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    suppress_blank: bool = True
    suppress_tokens: bool = True
    length_penalty: Optional[float] = None
    patience: Optional[float] = None
    prompt: Sequence[int] = ()


# ---------------------------------------------------------------------------------------------
def apply_filters(logits: torch.Tensor, tokens: torch.Tensor, sp: Specials, sample_begin: int,
                  opt: Options) -> None:
    """In-place logit masks in the reference order (decoding.py:581-597, :726-727)."""
    at_begin = tokens.shape[1] == sample_begin
    if opt.suppress_blank and at_begin:                                   # :455-457
        logits[:, list(sp.blank) + [sp.eot]] = -math.inf
    if opt.suppress_tokens:                                               # :464-465
        logits[:, list(sp.suppress)] = -math.inf
    if opt.without_timestamps:
        return
    tb = sp.timestamp_begin
    logits[:, sp.no_timestamps] = -math.inf                               # :481-482
    for k in range(tokens.shape[0]):                                      # :485-511
        seq = tokens[k, sample_begin:].tolist()
        last_ts = len(seq) >= 1 and seq[-1] >= tb
        penult_ts = len(seq) < 2 or seq[-2] >= tb
        if last_ts:
            if penult_ts:
                logits[k, tb:] = -math.inf
            else:
                logits[k, :sp.eot] = -math.inf
        stamps = [t for t in seq if t >= tb]
        if stamps:
            lim = stamps[-1] if (last_ts and not penult_ts) else stamps[-1] + 1
            logits[k, tb:lim] = -math.inf
    if at_begin:                                                          # :513-522
        logits[:, :tb] = -math.inf
        if opt.max_initial_timestamp:
            idx = round(opt.max_initial_timestamp / 0.02)
            logits[:, tb + idx + 1:] = -math.inf
    lp = F.log_softmax(logits.float(), dim=-1)                            # :525-532
    for k in range(tokens.shape[0]):
        if lp[k, tb:].logsumexp(dim=-1) > lp[k, :tb].max():
            logits[k, :tb] = -math.inf


class Greedy:
    """decoding.py:299-325 at temperature 0."""

    def __init__(self, eot):
        self.eot = eot

    def update(self, tokens, logits, sum_logprobs, model):
        nxt = logits.argmax(dim=-1)
        lp = F.log_softmax(logits.float(), dim=-1)
        cur = lp[torch.arange(lp.shape[0]), nxt]
        sum_logprobs += cur * (tokens[:, -1] != self.eot)
        nxt[tokens[:, -1] == self.eot] = self.eot
        tokens = torch.cat([tokens, nxt[:, None]], dim=-1)
        return tokens, bool((tokens[:, -1] == self.eot).all())

    def finalize(self, tokens, sum_logprobs):
        return [[F.pad(tokens[0, 0], (0, 1), value=self.eot)]], [sum_logprobs[0].tolist()]


class Beam:
    """decoding.py:328-431 for a single audio (n_audio == 1 is all the fork supports)."""

    def __init__(self, beam_size, eot, patience=None):
        self.bs, self.eot = beam_size, eot
        self.max_candidates = round(beam_size * (patience or 1.0))
        self.finished: Dict[tuple, float] = {}

    def update(self, tokens, logits, sum_logprobs, model: OracleModel):
        lp = F.log_softmax(logits.float(), dim=-1)
        scores, sources = {}, {}
        for j in range(self.bs):                                          # :366-373
            prefix = tokens[j].tolist()
            vals, idx = lp[j].topk(self.bs + 1)
            for v, t in zip(vals, idx):
                seq = tuple(prefix + [int(t)])
                scores[seq] = (sum_logprobs[j] + v).item()
                sources[seq] = j
        nxt, src, fin = [], [], {}
        for seq in sorted(scores, key=scores.get, reverse=True):          # :376-387
            if seq[-1] == self.eot:
                fin[seq] = scores[seq]
            else:
                sum_logprobs[len(nxt)] = scores[seq]
                nxt.append(seq); src.append(sources[seq])
                if len(nxt) == self.bs:
                    break
        tokens = torch.tensor(nxt)
        model.rearrange_kv_cache(src)                                     # :392
        for seq in sorted(fin, key=fin.get, reverse=True):                # :396-402
            if len(self.finished) >= self.max_candidates:
                break
            self.finished[seq] = fin[seq]
        self.last_sources = src
        return tokens, len(self.finished) >= self.max_candidates

    def finalize(self, tokens, sum_logprobs):
        """:411-431.  tokens (1, bs, n)."""
        sl = sum_logprobs.cpu()
        if len(self.finished) < self.bs:
            for j in list(np.argsort(sl[0]))[::-1]:
                self.finished[tuple(tokens[0, j].tolist() + [self.eot])] = sl[0][j].item()
                if len(self.finished) >= self.bs:
                    break
        return ([[torch.tensor(s) for s in self.finished.keys()]], [list(self.finished.values())])


@dataclass
class Result:
    tokens: List[int]
    avg_logprob: float
    no_speech_prob: float
    sum_logprob: float
    steps: int
    trace: List[List[int]]        # per-step beam token matrix (last column), for parity debugging


def decode_window(model: OracleModel, mel: Optional[torch.Tensor], sp: Specials, opt: Options) -> Result:
    """DecodingTask.run + _main_loop (decoding.py:707-816) for one 30-s window, temperature 0.
    mel=None reuses the encoder output already held by `model` (lets bench.py time the stages apart)."""
    model.reset()
    if mel is not None:
        model.encode(mel)
    sot_seq = list(sp.sot_sequence) + ([sp.no_timestamps] if opt.without_timestamps else [])
    initial = list(sot_seq)
    if opt.prompt:      # :628-638 - fixed-window sharding runs condition_on_previous_text=False
        raise NotImplementedError("prompt conditioning is outside the fixed-window hot path")
    sample_begin = len(initial)
    sot_index = initial.index(sp.sot)
    n_group = opt.beam_size or 1
    sample_len = opt.sample_len or model.dims.n_text_ctx // 2
    dec = Beam(opt.beam_size, sp.eot, opt.patience) if opt.beam_size else Greedy(sp.eot)

    tokens = torch.tensor([initial]).repeat_interleave(n_group, dim=0)
    sum_lp = torch.zeros(n_group)
    no_speech = math.nan
    trace = []
    steps = 0
    try:
        for i in range(sample_len):
            logits, _ = model.logits(tokens)
            if i == 0:                                                    # :716-720
                no_speech = logits[:, sot_index].float().softmax(dim=-1)[0, sp.no_speech].item()
            logits = logits[:, -1]
            apply_filters(logits, tokens, sp, sample_begin, opt)
            tokens, done = dec.update(tokens, logits, sum_lp, model)
            trace.append(tokens[:, -1].tolist())
            steps += 1
            if done or tokens.shape[-1] > model.dims.n_text_ctx:
                break
    finally:
        model.text_offset = 0                                             # cleanup_caching :186-187

    cands, lps = dec.finalize(tokens.reshape(1, n_group, -1), sum_lp.reshape(1, n_group))
    cands = [[t[sample_begin:(t == sp.eot).nonzero()[0, 0]] for t in s] for s in cands]  # :776-779
    lens = [len(t) for t in cands[0]]                                     # :226-240
    if opt.length_penalty is None:      # (a zero-length candidate raises in the reference; -inf here)
        with np.errstate(divide="ignore", invalid="ignore"):
            score = list(np.array(lps[0], dtype=np.float64) / np.array(lens, dtype=np.float64))
    else:
        score = [lp / (((5 + l) / 6) ** opt.length_penalty) for lp, l in zip(lps[0], lens)]
    sel = int(np.argmax(score))
    toks = cands[0][sel].tolist()
    return Result(toks, lps[0][sel] / (len(toks) + 1), no_speech, lps[0][sel], steps, trace)
