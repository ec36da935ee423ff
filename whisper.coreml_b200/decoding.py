"""Decoding: counterpart of whisper/decoding.py for the paths the fork runs (greedy, beam search at temperature 0, best_of
sampling at temperature > 0; one audio window per decode, many windows per call).  The loop itself - logits, logit filters,
log-softmax, top-k / Categorical sampling, beam bookkeeping, KV-cache permutation (decoding.py:707-737, 299-409, 450-532,
189-204) - executes on the device inside b200DecodeWindowsEx; this module prepares the initial tokens and ranks the candidates
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
    best_of: Optional[int] = None                # independent samples per window at temperature > 0 (decoding.py:88)
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Sequence[int]] = None
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    seed: int = 0                                # key of the device's counter-based sampler (temperature > 0)


@dataclass(frozen=True)
class DecodingResult:                        # whisper/decoding.py:118-128
    tokens: List[int] = field(default_factory=list)
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    sum_logprob: float = np.nan
    steps: int = 0
    candidates: int = 0
    temperature: float = 0.0


def decode_windows(model, options: DecodingOptions, windows: Sequence[int]) -> List[DecodingResult]:
    """DecodingTask.run (decoding.py:740-816) for several windows already encoded on the device.  The windows are
    independent (condition_on_previous_text=False), so the library decodes them concurrently (b200DecodeWindows)."""
    if options.temperature < 0.0:
        raise ValueError("temperature must be >= 0")
    if options.beam_size is not None and options.best_of is not None:
        raise ValueError("beam_size and best_of can't be given together")            # decoding.py:543
    if options.temperature == 0.0 and options.best_of is not None:
        raise ValueError("best_of with greedy sampling (T=0) is not compatible")      # decoding.py:545-546
    if options.temperature > 0.0 and options.beam_size is not None:
        raise ValueError("beam search runs at temperature 0 (whisper/transcribe.py:196-202 drops beam_size for t > 0)")
    if options.patience not in (None, 1.0):
        raise NotImplementedError("patience != 1")
    if options.prompt:
        raise NotImplementedError("prompt conditioning: windows are decoded independently (condition_on_previous_text=False)")
    windows = [int(w) for w in windows]
    if not windows:
        return []
    sp = model.specials
    initial = list(sp.sot_sequence) + ([sp.no_timestamps] if options.without_timestamps else [])
    n_ctx = model.dims.n_text_ctx
    sample_len = options.sample_len or n_ctx // 2
    bs = options.beam_size or 0
    n_group = 1 if bs else (options.best_of or 1)                           # decoding.py:549
    n_cand = max(bs, n_group)
    max_ts = -1
    if options.max_initial_timestamp:
        max_ts = round(options.max_initial_timestamp / 0.02)              # decoding.py:591-594 (time_precision 30/1500)
    nw = len(windows)
    init = np.array(initial, dtype=np.int32)
    wins = np.array(windows, dtype=np.int32)
    toks = np.empty((nw, n_cand, n_ctx + 1), dtype=np.int32)
    lens = np.empty((nw, n_cand), dtype=np.int32); lps = np.empty((nw, n_cand), dtype=np.float32)
    nsp = np.empty(nw, dtype=np.float32); steps = np.zeros(nw, dtype=np.int32)
    model.lib.b200DecodeWindowsEx(wins.ctypes.data_as(_lib.i32p), nw, init.ctypes.data_as(_lib.i32p), len(init), bs, n_group,
                                  float(options.temperature), int(options.seed) & 0xFFFFFFFFFFFFFFFF, sample_len,
                                  1 if options.without_timestamps else 0, max_ts,
                                  toks.ctypes.data_as(_lib.i32p), lens.ctypes.data_as(_lib.i32p),
                                  lps.ctypes.data_as(_lib.f32p), nsp.ctypes.data_as(_lib.f32p), steps.ctypes.data_as(_lib.i32p))
    _lib.check_errors("b200DecodeWindows")
    out = []
    n0 = len(initial)
    for w in range(nw):
        valid = [i for i in range(n_cand) if lens[w, i] >= 0]
        if not valid:
            raise RuntimeError("b200DecodeWindows returned no candidate")
        if options.length_penalty is None:                               # decoding.py:226-233
            with np.errstate(divide="ignore", invalid="ignore"):
                scores = [float(lps[w, i]) / float(lens[w, i]) for i in valid]
        else:
            scores = [float(lps[w, i]) / (((5 + int(lens[w, i])) / 6) ** options.length_penalty) for i in valid]
        best = valid[int(np.argmax(scores))]
        tokens = toks[w, best, n0:n0 + lens[w, best]].tolist()
        out.append(DecodingResult(tokens=tokens, avg_logprob=float(lps[w, best]) / (len(tokens) + 1), no_speech_prob=float(nsp[w]),
                                  sum_logprob=float(lps[w, best]), steps=int(steps[w]), candidates=len(valid), temperature=float(options.temperature)))
    return out


def decode(model, options: DecodingOptions = DecodingOptions(), window: Optional[int] = None) -> DecodingResult:
    """DecodingTask.run (decoding.py:740-816) for one window already encoded on the device."""
    if window is not None:
        model.select_window(window)
    return decode_windows(model, options, [model.current_window])[0]
