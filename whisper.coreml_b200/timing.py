"""Word-timestamp numerics on the device: counterpart of whisper/timing.py (median_filter :19-54,
dtw :141-160, find_alignment :163-231).  Word splitting / punctuation merging (:234-376) is host
string code that needs the tokenizer vocabulary; pass the reference's tokenizer to get words,
otherwise timings are reported per token."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .audio import TOKENS_PER_SECOND


def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    """whisper/timing.py:19-54: median filter along the last dimension with reflect padding."""
    pad_width = filter_width // 2
    if x.shape[-1] <= pad_width:
        return x                                               # F.pad requires the padding width to be smaller than the input dimension
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    xc = x.detach().to("cpu", torch.float32).contiguous()
    y = torch.empty_like(xc)
    length = xc.shape[-1]
    _lib.load().medianFilter(ctypes.cast(xc.data_ptr(), _lib.f32p), ctypes.cast(y.data_ptr(), _lib.f32p),
                             xc.numel() // length, length, filter_width)
    _lib.check_errors("medianFilter")
    return y.to(x.device)


def dtw(x) -> np.ndarray:
    """whisper/timing.py:141-160 (dtw_cpu tie rule, :95-100).  x: (N, M) cost matrix.  Returns a
    (2, path_len) int array: text indices, time indices."""
    xc = torch.as_tensor(x).detach().to("cpu", torch.float32).contiguous()
    n, m = xc.shape
    oi = np.empty(n + m, dtype=np.int32); oj = np.empty(n + m, dtype=np.int32)
    k = _lib.load().dtw(ctypes.cast(xc.data_ptr(), _lib.f32p), n, m,
                        oi.ctypes.data_as(_lib.i32p), oj.ctypes.data_as(_lib.i32p))
    _lib.check_errors("dtw")
    return np.stack([oi[:k], oj[:k]]).astype(np.int64)


@dataclass
class WordTiming:            # whisper/timing.py:154-160
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


@dataclass
class Alignment:
    text_indices: np.ndarray
    time_indices: np.ndarray
    matrix: np.ndarray                 # (n_text + 1, num_frames // 2): rows of text tokens + no_timestamps row
    text_token_probs: np.ndarray
    jump_times: np.ndarray             # start time of every aligned row, plus the trailing eot row


def align_tokens(sot_sequence: Sequence[int], no_timestamps: int, eot: int, text_tokens: Sequence[int], num_frames: int,
                 medfilt_width: int = 7) -> Optional[Alignment]:
    """Device part of find_alignment (whisper/timing.py:176-214) for the current window."""
    if len(text_tokens) == 0:
        return None
    tokens = np.array([*sot_sequence, no_timestamps, *text_tokens, eot], dtype=np.int32)
    n_tok, n_skip, F = len(tokens), len(sot_sequence), num_frames // 2
    n_rows, n_text = n_tok - 1 - n_skip, len(text_tokens)
    oi = np.empty(n_rows + F, dtype=np.int32); oj = np.empty(n_rows + F, dtype=np.int32)
    mat = np.empty((n_rows, F), dtype=np.float32); probs = np.empty(n_text, dtype=np.float32)
    k = _lib.load().b200AlignTokens(tokens.ctypes.data_as(_lib.i32p), n_tok, n_skip, num_frames, medfilt_width,
                                    oi.ctypes.data_as(_lib.i32p), oj.ctypes.data_as(_lib.i32p),
                                    mat.ctypes.data_as(_lib.f32p), probs.ctypes.data_as(_lib.f32p))
    _lib.check_errors("b200AlignTokens")
    ti, tj = oi[:k].astype(np.int64), oj[:k].astype(np.int64)
    jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)        # timing.py:216-217
    return Alignment(ti, tj, mat, probs, tj[jumps] / TOKENS_PER_SECOND)


def find_alignment(model, tokenizer, text_tokens: List[int], num_frames: int, *, medfilt_width: int = 7) -> List[WordTiming]:
    """whisper/timing.py:163-231.  `tokenizer` needs sot_sequence, no_timestamps, eot and, for word
    grouping, split_to_word_tokens (the reference tokenizer has all of them); without the latter each
    token is reported as its own 'word'."""
    al = align_tokens(tokenizer.sot_sequence, tokenizer.no_timestamps, tokenizer.eot, text_tokens, num_frames, medfilt_width)
    if al is None:
        return []
    if hasattr(tokenizer, "split_to_word_tokens"):
        words, word_tokens = tokenizer.split_to_word_tokens(list(text_tokens) + [tokenizer.eot])
    else:
        words = [str(t) for t in text_tokens] + [""]
        word_tokens = [[t] for t in text_tokens] + [[tokenizer.eot]]
    if len(word_tokens) <= 1:
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    start_times = al.jump_times[word_boundaries[:-1]]
    end_times = al.jump_times[word_boundaries[1:]]
    probs = [float(np.mean(al.text_token_probs[i:j])) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
    return [WordTiming(w, t, float(s), float(e), p) for w, t, s, e, p in zip(words, word_tokens, start_times, end_times, probs)]


PREPEND_PUNCTUATIONS = "\"'“¿([{-"
APPEND_PUNCTUATIONS = "\"'.。,，!！?？:：”)]}、"
SENTENCE_END_MARKS = ".。!！?？"


def merge_punctuations(alignment: List[WordTiming], prepended: str = PREPEND_PUNCTUATIONS, appended: str = APPEND_PUNCTUATIONS) -> None:
    """whisper/timing.py:234-265: glue leading punctuation to the word that follows it and trailing punctuation to the word before
    it (in place; a merged-away entry keeps an empty word and no tokens)."""
    # right to left: " (" + "word" -> " (word"; `tgt` is the nearest entry to the right that is not itself being prepended
    tgt = len(alignment) - 1
    for i in range(len(alignment) - 2, -1, -1):
        cur, nxt = alignment[i], alignment[tgt]
        if cur.word.startswith(" ") and cur.word.strip() in prepended:
            nxt.word, nxt.tokens = cur.word + nxt.word, cur.tokens + nxt.tokens
            cur.word, cur.tokens = "", []
        else:
            tgt = i
    # left to right: "word" + "." -> "word."; `tgt` is the nearest entry to the left that still owns text
    tgt = 0
    for j in range(1, len(alignment)):
        prev, cur = alignment[tgt], alignment[j]
        if not prev.word.endswith(" ") and cur.word in appended:
            prev.word, prev.tokens = prev.word + cur.word, prev.tokens + cur.tokens
            cur.word, cur.tokens = "", []
        else:
            tgt = j


def add_word_timestamps(segments: List[dict], alignment: List[WordTiming], eot: int, *, last_speech_timestamp: float,
                        prepend_punctuations: str = PREPEND_PUNCTUATIONS, append_punctuations: str = APPEND_PUNCTUATIONS) -> float:
    """whisper/timing.py:268-376 after its find_alignment call: `alignment` holds the word timings of the text tokens of all
    `segments` of one window (times relative to the window).  Clamps over-long words next to sentence ends and after pauses,
    merges punctuation, distributes the words over the segments and reconciles segment and word boundaries.  Adds "words" to
    every segment and returns the updated last_speech_timestamp."""
    if not segments:
        return last_speech_timestamp
    per_segment = [[t for t in seg["tokens"] if t < eot] for seg in segments]
    durations = np.array([w.end - w.start for w in alignment])
    durations = durations[durations.nonzero()]
    median = min(0.7, float(np.median(durations))) if len(durations) else 0.0
    longest = 2 * median
    if len(durations):                                                          # :299-308 long words at sentence boundaries
        for i in range(1, len(alignment)):
            w = alignment[i]
            if w.end - w.start > longest:
                if w.word in SENTENCE_END_MARKS:
                    w.end = w.start + longest
                elif alignment[i - 1].word in SENTENCE_END_MARKS:
                    w.start = w.end - longest
    merge_punctuations(alignment, prepend_punctuations, append_punctuations)
    from .audio import HOP_LENGTH, SAMPLE_RATE
    offset = segments[0]["seek"] * HOP_LENGTH / SAMPLE_RATE
    k = 0
    for seg, toks in zip(segments, per_segment):
        words, used = [], 0
        while k < len(alignment) and used < len(toks):                          # :317-332
            w = alignment[k]
            if w.word:
                words.append({"word": w.word, "start": round(offset + w.start, 2), "end": round(offset + w.end, 2), "probability": w.probability})
            used += len(w.tokens)
            k += 1
        if words:
            first = words[0]
            after_pause = first["end"] - last_speech_timestamp > median * 4     # :338-352 first words after a pause
            too_long = first["end"] - first["start"] > longest or (len(words) > 1 and words[1]["end"] - first["start"] > longest * 2)
            if after_pause and too_long:
                if len(words) > 1 and words[1]["end"] - words[1]["start"] > longest:
                    cut = max(words[1]["end"] / 2, words[1]["end"] - longest)
                    first["end"] = words[1]["start"] = cut
                first["start"] = max(0, first["end"] - longest)
            if seg["start"] < first["end"] and seg["start"] - 0.5 > first["start"]:      # :355-363 segment start vs first word
                first["start"] = max(0, min(first["end"] - median, seg["start"]))
            else:
                seg["start"] = first["start"]
            last = words[-1]
            if seg["end"] > last["start"] and seg["end"] + 0.5 < last["end"]:   # :366-374 segment end vs last word
                last["end"] = max(last["start"] + median, seg["end"])
            else:
                seg["end"] = last["end"]
            last_speech_timestamp = seg["end"]
        seg["words"] = words
    return last_speech_timestamp
