"""transcribe(): counterpart of whisper/transcribe.py:41-524 with condition_on_previous_text=False (windows are independent, which
is what lets them be batched, decoded concurrently and sharded over GPUs).

seek_mode="fixed" (default): fixed 30-s windows `seek = 3000 k`; windows are encoded in batches, decoded concurrently and sharded
`rank::world_size` with no collective in the loop.  A window that ends in an unfinished segment keeps its trailing tokens as a last
segment that ends at the window end (the reference would seek back to the last timestamp and decode that audio again, :383-388);
nothing is dropped, but a sentence that straddles a boundary is cut in two.
seek_mode="reference": the reference's data-dependent seek (:380-388, and :423-426 with word timestamps) - same segments as
whisper.transcribe().  Windows are decoded SPECULATIVELY on the fixed grid that starts at the current seek (sharded over the ranks,
batched on each), the ranks exchange the per-window results, and every rank then walks the reference's seek chain over them: a
window whose successor is on the grid is accepted, the first window that ends in an unfinished segment restarts the grid at its
last timestamp.  The host post-processing that depends on earlier windows (last_speech_timestamp) runs in that walk.

Temperature fallback (:188-228): `temperature` may be a tuple; a window whose result is too repetitive (compression ratio, needs a
tokenizer to turn tokens into text), too improbable (average log-probability) and not silence is decoded again at the next
temperature with `best_of` samples on the device (b200DecodeWindowsEx).  no_speech_threshold skips silent windows (:309-322).
The thresholds default to None / temperature 0 here (random-init weights must never fall back); the reference's defaults are
temperature=(0.0, 0.2, 0.4, 0.6, 0.8, 1.0), compression_ratio_threshold=2.4, logprob_threshold=-1.0, no_speech_threshold=0.6.

Word timestamps (:412-421): alignment (cross-attention, median filter, DTW) on the device; with the reference's tokenizer the
words are grouped and post-processed like add_word_timestamps (timing.py:207-231, 268-376), without one every text token is
reported as a word.  Not reproduced: prompt carry-over (:300-305), hallucination_silence_threshold (:428-470), clip_timestamps."""
from __future__ import annotations

import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram
from .decoding import DecodingOptions, DecodingResult, decode_windows
from .timing import (APPEND_PUNCTUATIONS, PREPEND_PUNCTUATIONS, Alignment, WordTiming, add_word_timestamps, align_tokens)

TIME_PRECISION = 0.02                                   # input_stride * HOP_LENGTH / SAMPLE_RATE (:250-252)
INPUT_STRIDE = 2                                        # N_FRAMES // n_audio_ctx (:249)


def _timestamp_structure(tokens: Sequence[int], timestamp_begin: int) -> Tuple[np.ndarray, List[int], bool]:
    t = np.array(tokens, dtype=np.int64)
    is_ts = t >= timestamp_begin
    single_timestamp_ending = is_ts[-2:].tolist() == [False, True]                       # :351
    consecutive = (np.where(is_ts[:-1] & is_ts[1:])[0] + 1).tolist() if len(t) > 1 else []   # :353-354
    return t, consecutive, single_timestamp_ending


def _segments_from_tokens(tokens: List[int], result: DecodingResult, time_offset: float, segment_duration: float, seek: int,
                          timestamp_begin: int, keep_tail: bool) -> List[dict]:
    """Slice a window's tokens into segments at consecutive timestamp pairs (transcribe.py:350-410).  keep_tail (fixed windows):
    the tokens after the last pair, which the reference re-decodes from the next seek, become a last segment up to the window end."""
    t, consecutive, single_timestamp_ending = _timestamp_structure(tokens, timestamp_begin)
    is_ts = t >= timestamp_begin

    def new_segment(start, end, toks):
        return {"seek": seek, "start": float(start), "end": float(end), "tokens": [int(x) for x in toks], "temperature": result.temperature,
                "avg_logprob": result.avg_logprob, "no_speech_prob": result.no_speech_prob}

    segs = []
    if consecutive:
        slices = consecutive + ([len(t)] if single_timestamp_ending else [])
        last = 0
        for cur in slices:
            sl = t[last:cur]
            segs.append(new_segment(time_offset + (sl[0] - timestamp_begin) * TIME_PRECISION,
                                    time_offset + (sl[-1] - timestamp_begin) * TIME_PRECISION, sl))
            last = cur
        if keep_tail and not single_timestamp_ending and last < len(t) and not is_ts[last:].all():     # (a lone opening timestamp is no segment)
            tail = t[last:]
            start = time_offset + (tail[0] - timestamp_begin) * TIME_PRECISION if is_ts[last] else segs[-1]["end"]
            segs.append(new_segment(start, time_offset + segment_duration, tail))
    else:
        duration = segment_duration
        stamps = t[is_ts]
        if len(stamps) > 0 and stamps[-1] != timestamp_begin:
            duration = (stamps[-1] - timestamp_begin) * TIME_PRECISION
        segs.append(new_segment(time_offset, time_offset + duration, t))
    return segs


def _next_seek(tokens: List[int], seek: int, segment_size: int, timestamp_begin: int) -> int:
    """(:376-388, :409) where the reference's loop continues after a window with these tokens."""
    t, consecutive, single_timestamp_ending = _timestamp_structure(tokens, timestamp_begin)
    if consecutive and not single_timestamp_ending:
        return seek + int(t[consecutive[-1] - 1] - timestamp_begin) * INPUT_STRIDE       # ignore the unfinished segment, seek to the last timestamp
    return seek + segment_size


def _compression_ratio(text: str) -> float:
    data = text.encode("utf-8")                          # whisper/utils.py compression_ratio
    return len(data) / len(zlib.compress(data))


@dataclass
class _Window:
    """What the device produced for one window; everything that depends on other windows is computed from it on the host."""
    seek: int
    segment_size: int
    result: DecodingResult
    alignment: Optional[Alignment] = None                # of the text tokens of the window's segments (word timestamps)
    text_tokens: Optional[List[int]] = None

    def pack(self) -> dict:                              # (plain containers: travels through all_gather_object)
        al = self.alignment
        return {"seek": self.seek, "segment_size": self.segment_size, "result": self.result.__dict__, "text_tokens": self.text_tokens,
                "alignment": None if al is None else {"jump_times": al.jump_times.tolist(), "probs": al.text_token_probs.tolist()}}

    @staticmethod
    def unpack(d: dict) -> "_Window":
        al = d["alignment"]
        alignment = None if al is None else Alignment(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 0), np.float32),
                                                      np.array(al["probs"], dtype=np.float32), np.array(al["jump_times"]))
        return _Window(d["seek"], d["segment_size"], DecodingResult(**d["result"]), alignment, d["text_tokens"])


def transcribe(model, audio: torch.Tensor, *, beam_size: Optional[int] = 5, best_of: Optional[int] = 5,
               temperature: Union[float, Tuple[float, ...]] = 0.0, compression_ratio_threshold: Optional[float] = None,
               logprob_threshold: Optional[float] = None, no_speech_threshold: Optional[float] = None,
               word_timestamps: bool = False, sample_len: Optional[int] = None, without_timestamps: bool = False,
               length_penalty: Optional[float] = None, window_batch: int = 8, rank: int = 0, world_size: int = 1, tokenizer=None,
               verbose: bool = False, seek_mode: str = "fixed", seed: int = 0,
               prepend_punctuations: str = PREPEND_PUNCTUATIONS, append_punctuations: str = APPEND_PUNCTUATIONS) -> dict:
    """audio: 1-D 16 kHz float waveform (CPU or CUDA).  Returns {"segments": [...], "windows": n, "seeks": [...], "language": "en"};
    with a tokenizer (the reference's) segments also carry "text".  Rank r of world_size handles windows r::world_size."""
    if seek_mode not in ("fixed", "reference"):
        raise ValueError(f"seek_mode {seek_mode!r}: 'fixed' or 'reference'")
    model.load()
    dims, sp = model.dims, model.specials
    dev = getattr(model, "device", None) or f"cuda:{model.device_index}"    # (tests/test_fallback_ladder.py drives the host logic with a scripted backend)
    mel = log_mel_spectrogram(audio.to(dev), dims.n_mels, padding=N_SAMPLES)          # transcribe.py:143 (global max per file)
    content_frames = mel.shape[-1] - N_FRAMES
    temperatures = [float(temperature)] if isinstance(temperature, (int, float)) else [float(t) for t in temperature]
    decode_steps: List[int] = []

    def usable(seek: int) -> bool:
        return seek < content_frames and min(N_FRAMES, content_frames - seek) * HOP_LENGTH / SAMPLE_RATE >= 1.0      # :295-298

    def needs_fallback(r: DecodingResult) -> bool:                                     # :207-226
        bad = False
        if compression_ratio_threshold is not None and tokenizer is not None and r.tokens:
            bad = bad or _compression_ratio(tokenizer.decode([t for t in r.tokens if t < sp.eot])) > compression_ratio_threshold
        if logprob_threshold is not None and r.avg_logprob < logprob_threshold:
            bad = True
        if no_speech_threshold is not None and r.no_speech_prob > no_speech_threshold and logprob_threshold is not None \
                and r.avg_logprob < logprob_threshold:
            bad = False                                                                # silence
        return bad

    def decode_batch(seeks: List[int]) -> List[_Window]:
        """encode + decode (with temperature fallback) + raw alignment of the windows at `seeks`, all on the device"""
        if not seeks:
            return []
        model.encode_windows(mel, seeks, content_frames)
        results: List[Optional[DecodingResult]] = [None] * len(seeks)
        pending = list(range(len(seeks)))
        for t in temperatures:                                                         # decode_with_fallback, a batch at a time
            if t > 0:
                opts = DecodingOptions(temperature=t, best_of=best_of, sample_len=sample_len, without_timestamps=without_timestamps,
                                       length_penalty=length_penalty, seed=seed)
            else:
                opts = DecodingOptions(beam_size=beam_size, sample_len=sample_len, without_timestamps=without_timestamps,
                                       length_penalty=length_penalty)
            for i, r in zip(pending, decode_windows(model, opts, pending)):
                results[i] = r
                decode_steps.append(r.steps)
            pending = [i for i in pending if needs_fallback(results[i])]
            if not pending:
                break
        out = []
        for w, seek in enumerate(seeks):
            win = _Window(seek, min(N_FRAMES, content_frames - seek), results[w])
            if word_timestamps and results[w].tokens and not skipped(results[w]):
                segs = _segments_from_tokens(results[w].tokens, results[w], 0.0, 0.0, seek, sp.timestamp_begin, seek_mode == "fixed")
                win.text_tokens = [t for s in segs for t in s["tokens"] if t < sp.eot]   # timing.py:283-287
                model.select_window(w)                                                 # the alignment reads this window's cross K/V
                win.alignment = align_tokens(sp.sot_sequence, sp.no_timestamps, sp.eot, win.text_tokens, win.segment_size)
            out.append(win)
        return out

    def skipped(r: DecodingResult) -> bool:                                            # :309-322
        if no_speech_threshold is None or not (r.no_speech_prob > no_speech_threshold):
            return False
        return not (logprob_threshold is not None and r.avg_logprob > logprob_threshold)

    state = {"last_speech": 0.0}

    def finish_window(win: _Window) -> Tuple[List[dict], int]:
        """segments (+ words, text) of one decoded window and the seek the reference continues from (host only)"""
        r, seek, segment_size = win.result, win.seek, win.segment_size
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        if skipped(r):
            return [], seek + segment_size
        if not r.tokens:                           # nothing decoded: the reference still reports the (empty) window as one segment
            cur = [{"seek": seek, "start": time_offset, "end": time_offset + segment_size * HOP_LENGTH / SAMPLE_RATE, "tokens": [],
                    "temperature": r.temperature, "avg_logprob": r.avg_logprob, "no_speech_prob": r.no_speech_prob}]
            if tokenizer is not None:
                cur[0]["text"] = ""
            if word_timestamps:
                cur[0]["words"] = []
            return cur, seek + segment_size
        cur = _segments_from_tokens(r.tokens, r, time_offset, segment_size * HOP_LENGTH / SAMPLE_RATE, seek, sp.timestamp_begin,
                                    seek_mode == "fixed")
        nxt = _next_seek(r.tokens, seek, segment_size, sp.timestamp_begin)
        if tokenizer is not None:
            for s in cur:
                s["text"] = tokenizer.decode([t for t in s["tokens"] if t < sp.eot])
        if word_timestamps:
            al, text = win.alignment, win.text_tokens or []
            if al is None:
                for s in cur:
                    s["words"] = []
            elif tokenizer is not None and hasattr(tokenizer, "split_to_word_tokens"):
                words, word_tokens = tokenizer.split_to_word_tokens(list(text) + [sp.eot])          # timing.py:207-231
                timings: List[WordTiming] = []
                if len(word_tokens) > 1:
                    bounds = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
                    for w, tk, a, b in zip(words, word_tokens, bounds[:-1], bounds[1:]):
                        timings.append(WordTiming(w, list(tk), float(al.jump_times[a]), float(al.jump_times[b]),
                                                  float(np.mean(al.text_token_probs[a:b]))))
                state["last_speech"] = add_word_timestamps(cur, timings, sp.eot, last_speech_timestamp=state["last_speech"],
                                                           prepend_punctuations=prepend_punctuations, append_punctuations=append_punctuations)
                _, _, single_ending = _timestamp_structure(r.tokens, sp.timestamp_begin)
                ends = [w["end"] for s in cur for w in s["words"]]
                if seek_mode == "reference" and not single_ending and ends and ends[-1] > time_offset:   # :423-426
                    nxt = round(ends[-1] * FRAMES_PER_SECOND)
            else:                                                                       # no tokenizer: one "word" per text token
                pos = 0
                for s in cur:
                    toks = [t for t in s["tokens"] if t < sp.eot]
                    s["words"] = [{"token": toks[k], "start": round(time_offset + float(al.jump_times[pos + k]), 2),
                                   "end": round(time_offset + float(al.jump_times[pos + k + 1]), 2),
                                   "probability": float(al.text_token_probs[pos + k])} for k in range(len(toks))]
                    pos += len(toks)
        if verbose:
            for s in cur:
                print(f"[{s['start']:.2f} --> {s['end']:.2f}] {len(s['tokens'])} tokens")
        if tokenizer is not None:                  # "if a segment is instantaneous or does not contain text, clear it" (:495-500);
            for s in cur:                          # without a tokenizer the caller gets the raw token segments
                if s["start"] == s["end"] or s["text"].strip() == "":
                    s["text"], s["tokens"] = "", []
                    if "words" in s:
                        s["words"] = []
        return cur, nxt

    def exchange(mine: List[_Window]) -> Dict[int, _Window]:
        """every rank's windows of a speculative round, on every rank (host objects; the only cross-rank exchange)"""
        if world_size == 1:
            return {w.seek: w for w in mine}
        import torch.distributed as dist
        parts = [None] * world_size
        dist.all_gather_object(parts, [w.pack() for w in mine])
        return {d["seek"]: _Window.unpack(d) for p in parts for d in p}

    segments: List[dict] = []
    seeks_done: List[int] = []
    if seek_mode == "fixed":
        mine = [s for s in range(0, content_frames, N_FRAMES) if usable(s)][rank::world_size]
        for b0 in range(0, len(mine), window_batch):
            for win in decode_batch(mine[b0:b0 + window_batch]):                       # independent windows decode concurrently on the device
                segments.extend(finish_window(win)[0])
                seeks_done.append(win.seek)
    else:
        seek, cache = 0, {}
        while usable(seek):
            if seek not in cache:                                                      # a new speculative grid from the current seek
                grid = [s for s in range(seek, content_frames, N_FRAMES) if usable(s)][:window_batch * world_size]
                cache = exchange(decode_batch(grid[rank::world_size]))
            win = cache.pop(seek)
            cur, nxt = finish_window(win)
            segments.extend(cur)
            seeks_done.append(seek)
            if nxt <= seek:                                                            # the reference would loop forever here
                break
            seek = nxt
    for i, s in enumerate(segments):
        s["id"] = i
    out = {"segments": segments, "windows": len(seeks_done), "seeks": seeks_done, "language": "en",
           "audio_seconds": audio.numel() / SAMPLE_RATE, "decode_steps": decode_steps}
    if tokenizer is not None:
        out["text"] = "".join(s.get("text", "") for s in segments)
    return out


def gather_sharded(result: dict, world_size: int) -> dict:
    """Merge per-rank results of a fixed-window run on every rank by window start (host-side; the only cross-rank exchange)."""
    if world_size == 1:
        return result
    import torch.distributed as dist
    parts = [None] * world_size
    dist.all_gather_object(parts, result)
    segs = sorted((s for p in parts for s in p["segments"]), key=lambda s: (s["seek"], s["start"]))
    for i, s in enumerate(segs):
        s["id"] = i
    merged = dict(result)
    merged.update(segments=segs, windows=sum(p["windows"] for p in parts), seeks=sorted(x for p in parts for x in p["seeks"]))
    return merged
