"""Decoding: counterpart of whisper/decoding.py for the paths the fork runs (temperature 0; greedy or
beam search; one audio window per call).  The loop itself - logits, logit filters, log-softmax, top-k,
beam bookkeeping, KV-cache permutation (decoding.py:707-737, 350-409, 450-532, 189-204) - executes on
the device inside b200DecodeWindow; this module prepares the initial tokens and ranks the candidates
(MaximumLikelihoodRanker, decoding.py:217-240)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib


@dataclass(frozen=True)
class DecodingOptions:                       # subset of whisper/decoding.py:81-115 that reaches the hot loop
    task: str = "transcribe"
    language: Optional[str] = "en"
    temperature: float = 0.0
    sample_len: Optional[int] = None
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Sequence[int]] = None
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0


@dataclass(frozen=True)
class DecodingResult:                        # whisper/decoding.py:118-128
    tokens: List[int] = field(default_factory=list)
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    sum_logprob: float = np.nan
    steps: int = 0
    candidates: int = 0


def decode(model, options: DecodingOptions = DecodingOptions(), window: Optional[int] = None) -> DecodingResult:
    """DecodingTask.run (decoding.py:740-816) for the window already encoded on the device."""
    if options.temperature != 0.0:
        raise NotImplementedError("sampling at temperature > 0 (decoding.py:307-310) is not on the B200 hot path")
    if options.patience not in (None, 1.0):
        raise NotImplementedError("patience != 1")
    if options.prompt:
        raise NotImplementedError("prompt conditioning: windows are decoded independently (condition_on_previous_text=False)")
    if window is not None:
        model.select_window(window)
    sp = model.specials
    initial = list(sp.sot_sequence) + ([sp.no_timestamps] if options.without_timestamps else [])
    n_ctx = model.dims.n_text_ctx
    sample_len = options.sample_len or n_ctx // 2
    bs = options.beam_size or 0
    n_cand = max(bs, 1)
    max_ts = -1
    if options.max_initial_timestamp:
        max_ts = round(options.max_initial_timestamp / 0.02)              # decoding.py:591-594 (time_precision 30/1500)
    init = np.array(initial, dtype=np.int32)
    toks = np.empty((n_cand, n_ctx + 1), dtype=np.int32)
    lens = np.empty(n_cand, dtype=np.int32); lps = np.empty(n_cand, dtype=np.float32); nsp = np.empty(1, dtype=np.float32)
    steps = model.lib.b200DecodeWindow(init.ctypes.data_as(_lib.i32p), len(init), bs, sample_len,
                                       1 if options.without_timestamps else 0, max_ts,
                                       toks.ctypes.data_as(_lib.i32p), lens.ctypes.data_as(_lib.i32p),
                                       lps.ctypes.data_as(_lib.f32p), nsp.ctypes.data_as(_lib.f32p))
    _lib.check_errors("b200DecodeWindow")
    valid = [i for i in range(n_cand) if lens[i] >= 0]
    if not valid:
        raise RuntimeError("b200DecodeWindow returned no candidate")
    if options.length_penalty is None:                                   # decoding.py:226-233
        with np.errstate(divide="ignore", invalid="ignore"):
            scores = [float(lps[i]) / float(lens[i]) for i in valid]
    else:
        scores = [float(lps[i]) / (((5 + int(lens[i])) / 6) ** options.length_penalty) for i in valid]
    best = valid[int(np.argmax(scores))]
    n0 = len(initial)
    tokens = toks[best, n0:n0 + lens[best]].tolist()
    return DecodingResult(tokens=tokens, avg_logprob=float(lps[best]) / (len(tokens) + 1), no_speech_prob=float(nsp[0]),
                          sum_logprob=float(lps[best]), steps=int(steps), candidates=len(valid))
