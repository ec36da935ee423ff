"""transcribe(): counterpart of whisper/transcribe.py:41-524 for the configuration the B200 hot path
targets - condition_on_previous_text=False, temperature 0.

seek_mode="fixed" (default): fixed 30-s windows, which makes every window independent, so windows are encoded in batches,
decoded concurrently (decode lanes) and sharded over GPUs (`rank::world_size`) with no collective in the loop.
seek_mode="reference": the reference's data-dependent seek (transcribe.py:380-388: a window that ends in an unfinished segment
makes the next window start at its last timestamp) - windows become sequential, one GPU; same segments as whisper.transcribe().
Segment slicing by timestamp tokens (:350-410), the zero-padded partial last window (:286-290) and word timestamps (:412-421)
are reproduced in both modes; temperature fallback (:188-228), prompt carry-over (:300-305) and the word-timestamp seek
adjustments (:423-470) are not."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram
from .decoding import DecodingOptions, DecodingResult, decode, decode_windows
from .timing import align_tokens


def _segments_from_tokens(tokens: List[int], result: DecodingResult, time_offset: float, segment_duration: float, seek: int,
                          timestamp_begin: int, eot: int) -> List[dict]:
    """Slice a window's tokens into segments at consecutive timestamp pairs (transcribe.py:350-410)."""
    time_precision = 0.02
    t = np.array(tokens, dtype=np.int64)
    is_ts = t >= timestamp_begin
    single_timestamp_ending = is_ts[-2:].tolist() == [False, True]
    consecutive = (np.where(is_ts[:-1] & is_ts[1:])[0] + 1).tolist() if len(t) > 1 else []

    def new_segment(start, end, toks):
        toks = [int(x) for x in toks]
        return {"seek": seek, "start": start, "end": end, "tokens": toks, "temperature": 0.0, "avg_logprob": result.avg_logprob,
                "no_speech_prob": result.no_speech_prob}

    segs = []
    if consecutive:
        slices = consecutive + ([len(t)] if single_timestamp_ending else [])
        last = 0
        for cur in slices:
            sl = t[last:cur]
            segs.append(new_segment(time_offset + (sl[0] - timestamp_begin) * time_precision,
                                    time_offset + (sl[-1] - timestamp_begin) * time_precision, sl))
            last = cur
    else:
        duration = segment_duration
        stamps = t[is_ts]
        if len(stamps) > 0 and stamps[-1] != timestamp_begin:
            duration = (stamps[-1] - timestamp_begin) * time_precision
        segs.append(new_segment(time_offset, time_offset + duration, t))
    return segs


def transcribe(model, audio: torch.Tensor, *, beam_size: Optional[int] = 5, word_timestamps: bool = False,
               sample_len: Optional[int] = None, without_timestamps: bool = False, length_penalty: Optional[float] = None,
               window_batch: int = 8, rank: int = 0, world_size: int = 1, tokenizer=None, verbose: bool = False,
               seek_mode: str = "fixed") -> dict:
    """audio: 1-D 16 kHz float waveform (CPU or CUDA).  Returns {"segments": [...], "windows": n, "language": "en"};
    with a tokenizer (the reference's) segments also carry "text".  Rank r of world_size handles windows r::world_size."""
    model.load()
    dims, sp = model.dims, model.specials
    dev = f"cuda:{model.device_index}"
    mel = log_mel_spectrogram(audio.to(dev), dims.n_mels, padding=N_SAMPLES)          # transcribe.py:143 (global max per file)
    content_frames = mel.shape[-1] - N_FRAMES
    seeks_all = [s for s in range(0, content_frames, N_FRAMES)
                 if min(N_FRAMES, content_frames - s) * HOP_LENGTH / SAMPLE_RATE >= 1.0]      # :295-298
    mine = seeks_all[rank::world_size]
    opts = DecodingOptions(beam_size=beam_size, sample_len=sample_len, without_timestamps=without_timestamps,
                           length_penalty=length_penalty)
    segments: List[dict] = []
    decode_steps: List[int] = []

    def finish_window(result, seek):
        """segments (+ word timestamps, text) of one decoded window; the window's cross K/V must be selected"""
        decode_steps.append(result.steps)
        segment_size = min(N_FRAMES, content_frames - seek)
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        if not result.tokens:
            return []
        cur = _segments_from_tokens(result.tokens, result, time_offset, segment_size * HOP_LENGTH / SAMPLE_RATE, seek,
                                    sp.timestamp_begin, sp.eot)
        if word_timestamps:
            per_seg = [[t for t in s["tokens"] if t < sp.eot] for s in cur]            # timing.py:283-287
            text_tokens = [t for seg in per_seg for t in seg]
            al = align_tokens(sp.sot_sequence, sp.no_timestamps, sp.eot, text_tokens, segment_size)
            pos = 0
            for s, toks in zip(cur, per_seg):
                words = []
                for k in range(len(toks)):
                    if al is None:
                        break
                    words.append({"token": toks[k], "start": round(time_offset + float(al.jump_times[pos + k]), 2),
                                  "end": round(time_offset + float(al.jump_times[pos + k + 1]), 2),
                                  "probability": float(al.text_token_probs[pos + k])})
                pos += len(toks)
                s["words"] = words
        if tokenizer is not None:
            for s in cur:
                s["text"] = tokenizer.decode([t for t in s["tokens"] if t < sp.eot])
        if verbose:
            for s in cur:
                print(f"[{s['start']:.2f} --> {s['end']:.2f}] {len(s['tokens'])} tokens")
        return cur

    if seek_mode == "reference":
        if world_size != 1:
            raise ValueError("seek_mode='reference' makes the windows sequential: it cannot be sharded")
        mine, seek = [], 0
        while seek < content_frames:                                      # transcribe.py:277-298 (one clip)
            segment_size = min(N_FRAMES, content_frames - seek)
            if segment_size * HOP_LENGTH / SAMPLE_RATE < 1.0:
                break
            model.encode_windows(mel, [seek], content_frames)
            result = decode_windows(model, opts, [0])[0]
            model.select_window(0)
            mine.append(seek)
            segments.extend(finish_window(result, seek))
            t = np.array(result.tokens, dtype=np.int64)
            is_ts = t >= sp.timestamp_begin
            consecutive = (np.where(is_ts[:-1] & is_ts[1:])[0] + 1).tolist() if len(t) > 1 else []
            if consecutive and is_ts[-2:].tolist() != [False, True]:      # unfinished last segment: seek to the last timestamp (:383-388)
                advance = int(t[consecutive[-1] - 1] - sp.timestamp_begin) * (N_FRAMES // dims.n_audio_ctx)
            else:
                advance = segment_size                                    # :380-382, :409
            if advance <= 0:                                              # the reference would loop forever here
                break
            seek += advance
    elif seek_mode != "fixed":
        raise ValueError(f"seek_mode {seek_mode!r}: 'fixed' or 'reference'")
    for b0 in range(0, len(mine) if seek_mode == "fixed" else 0, window_batch):
        batch = mine[b0:b0 + window_batch]
        model.encode_windows(mel, batch, content_frames)
        results = decode_windows(model, opts, range(len(batch)))          # independent windows decode concurrently on the device
        for w, seek in enumerate(batch):
            model.select_window(w)                                        # word timestamps read this window's cross K/V
            segments.extend(finish_window(results[w], seek))
    for i, s in enumerate(segments):
        s["id"] = i
    out = {"segments": segments, "windows": len(mine), "seeks": mine, "language": "en",
           "audio_seconds": audio.numel() / SAMPLE_RATE, "decode_steps": decode_steps}
    if tokenizer is not None:
        out["text"] = "".join(s.get("text", "") for s in segments)
    return out


def gather_sharded(result: dict, world_size: int) -> dict:
    """Merge per-rank results on every rank by window start (host-side; the only cross-rank exchange)."""
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
