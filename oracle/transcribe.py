"""CPU restatement of the reference's transcribe() main loop (whisper/transcribe.py:260-470) - TEST INFRASTRUCTURE, see
oracle/__init__.py - for the configuration the B200 path serves: condition_on_previous_text=False, temperature 0, no
compression / logprob / no-speech thresholds (random-init weights must never trigger a fallback), one clip, no word timestamps.
Unlike the fixed-window schedule of the sharded hot path, this keeps the reference's data-dependent seek: a window that ends in
an unfinished segment makes the next window start at its last timestamp (:380-388)."""
from typing import Dict, List

import torch

from . import audio as oa, decoding as od, model as om

N_FRAMES, HOP_LENGTH, SAMPLE_RATE = 3000, 160, 16000
INPUT_STRIDE = 2                                        # N_FRAMES // n_audio_ctx (:249)
TIME_PRECISION = 0.02                                   # input_stride * HOP_LENGTH / SAMPLE_RATE (:250-252)


def window_schedule_step(tokens: List[int], seek: int, segment_size: int, timestamp_begin: int, time_offset: float):
    """(:350-410) -> (segments [(start, end, tokens)], next seek) of one decoded window."""
    t = torch.tensor(tokens, dtype=torch.long)
    ts = t.ge(timestamp_begin)
    single_timestamp_ending = ts[-2:].tolist() == [False, True]
    consecutive = (torch.where(ts[:-1] & ts[1:])[0] + 1).tolist() if len(t) > 1 else []
    segs = []
    if consecutive:
        slices = consecutive + ([len(t)] if single_timestamp_ending else [])
        last = 0
        for cur in slices:
            sl = t[last:cur]
            segs.append((time_offset + (sl[0].item() - timestamp_begin) * TIME_PRECISION,
                         time_offset + (sl[-1].item() - timestamp_begin) * TIME_PRECISION, sl.tolist()))
            last = cur
        if single_timestamp_ending:
            seek += segment_size                        # no speech after the last timestamp (:380-382)
        else:                                           # ignore the unfinished segment, seek to the last timestamp (:383-388)
            seek += (t[last - 1].item() - timestamp_begin) * INPUT_STRIDE
    else:
        duration = segment_size * HOP_LENGTH / SAMPLE_RATE
        stamps = t[ts]
        if len(stamps) > 0 and stamps[-1].item() != timestamp_begin:
            duration = (stamps[-1].item() - timestamp_begin) * TIME_PRECISION
        segs.append((time_offset, time_offset + duration, t.tolist()))
        seek += segment_size
    return segs, seek


def transcribe(model: om.OracleModel, audio: torch.Tensor, sp: od.Specials, opt: od.Options, max_windows: int = 1000) -> Dict:
    mel = oa.log_mel_spectrogram(audio, model.dims.n_mels, padding=480000)          # :143
    content_frames = mel.shape[-1] - N_FRAMES
    seek, seeks, segments, window_tokens = 0, [], [], []
    while seek < content_frames and len(seeks) < max_windows:                       # :277-284 (one clip)
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        segment_size = min(N_FRAMES, content_frames - seek)
        if segment_size * HOP_LENGTH / SAMPLE_RATE < 1.0:                           # :295-298
            break
        mel_segment = oa.pad_or_trim(mel[:, seek:seek + segment_size], N_FRAMES).contiguous()   # ZERO padded (:288-290)
        res = od.decode_window(model, mel_segment, sp, opt)
        seeks.append(seek)
        window_tokens.append(list(res.tokens))
        segs, new_seek = window_schedule_step(res.tokens, seek, segment_size, sp.timestamp_begin, time_offset)
        segments += [dict(seek=seek, start=a, end=b, tokens=tk) for a, b, tk in segs]
        if new_seek <= seek:                            # the reference would loop forever on a window that ends at <|0.00|>
            break
        seek = new_seek
    return dict(seeks=seeks, window_tokens=window_tokens, segments=segments)
