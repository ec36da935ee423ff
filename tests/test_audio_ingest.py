"""Audio ingest (the step in front of log_mel_spectrogram; the reference shells out to ffmpeg, whisper/audio.py:45-62).

CPU part: the FLAC decoder of libwhisper_b200.so (host code, csrc/flac.cu) is bit-exact - the MD5 of the decoded PCM must equal the
MD5 the encoder stored in STREAMINFO (the format's own known-answer test) for the reference's tests/jfk.flac (44.1 kHz, stereo,
24 bit: LPC subframes, mid/side stereo, partitioned Rice codes), and hand-assembled streams cover the subframe kinds that file
does not use (constant, verbatim, fixed predictors, escaped partitions, wasted bits, left/side and side/right stereo).

GPU part: BASELINE.json configs[0] - tiny, greedy, tests/jfk.flac through load_audio() and transcribe(), against the oracle fed
with the same decoded samples."""
import hashlib
import os

import numpy as np
import pytest

JFK = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jfk.flac")       # copy of /root/reference/tests/jfk.flac


def _pcm_md5(pcm: np.ndarray, bits: int) -> str:
    width = (bits + 7) // 8
    raw = pcm.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :width]
    return hashlib.md5(raw.tobytes()).hexdigest()


def test_flac_decode_matches_streaminfo_md5():
    from whisper_b200.audio import decode_flac
    with open(JFK, "rb") as f:
        pcm, rate, bits, md5 = decode_flac(f.read())
    assert (rate, bits, pcm.shape) == (44100, 24, (485100, 2))                      # 11.0 s (tests/test_audio.py:8-19 of the reference)
    assert _pcm_md5(pcm, bits) == md5.hex()


# ---- a tiny FLAC writer (test infrastructure): one frame per block, one subframe kind per test ----------------------------
class _Bits:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, v, k):
        for i in range(k - 1, -1, -1):
            self.acc = (self.acc << 1) | ((v >> i) & 1)
            self.n += 1
            if self.n == 8:
                self.out.append(self.acc); self.acc, self.n = 0, 0

    def unary(self, q):
        self.put(0, q) if q else None
        self.put(1, 1)

    def align(self):
        if self.n:
            self.put(0, 8 - self.n)


def _rice(b, vals, k):
    for v in vals:
        u = (v << 1) ^ (v >> 63) if v >= 0 else ((-v) << 1) - 1
        b.unary(u >> k)
        if k:
            b.put(u & ((1 << k) - 1), k)


def _subframe(b, kind, x, bps, wasted=0):
    x = [int(v) >> wasted for v in x]
    bps -= wasted
    code = {"constant": 0, "verbatim": 1}.get(kind, None)
    if code is None:
        code = 8 + int(kind[-1])                                                    # "fixed0" .. "fixed4"
    b.put(0, 1); b.put(code, 6)
    if wasted:
        b.put(1, 1); b.unary(wasted - 1)
    else:
        b.put(0, 1)
    mask = (1 << bps) - 1
    if kind == "constant":
        b.put(x[0] & mask, bps)
    elif kind == "verbatim":
        for v in x:
            b.put(v & mask, bps)
    else:
        order = int(kind[-1])
        for v in x[:order]:
            b.put(v & mask, bps)
        res = np.array(x, dtype=np.int64)
        for _ in range(order):
            res = np.concatenate([res[:1], np.diff(res)])                          # fixed predictors are repeated differences
        res = res[order:].tolist()
        b.put(0, 2); b.put(1, 4)                                                    # Rice method 0, partition order 1: two partitions
        half = len(x) // 2
        first, second = res[:half - order], res[half - order:]
        b.put(3, 4); _rice(b, first, 3)                                             # partition 0: Rice parameter 3
        b.put(15, 4); b.put(20, 5)                                                  # partition 1: escape, 20 raw bits per residual
        for v in second:
            b.put(v & ((1 << 20) - 1), 20)


def _flac_stream(channels_pcm, bps, rate, kinds, ch_code, blocksize=192, wasted=0):
    n = len(channels_pcm[0])
    assert n % blocksize == 0
    raw = np.stack(channels_pcm, axis=1).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :(bps + 7) // 8]
    md5 = hashlib.md5(raw.tobytes()).digest()
    b = _Bits()
    for c in b"fLaC":
        b.put(c, 8)
    b.put(1, 1); b.put(0, 7); b.put(34, 24)                                         # last metadata block: STREAMINFO
    b.put(blocksize, 16); b.put(blocksize, 16); b.put(0, 24); b.put(0, 24)
    b.put(rate, 20); b.put(len(channels_pcm) - 1, 3); b.put(bps - 1, 5); b.put(n, 36)
    for c in md5:
        b.put(c, 8)
    left, right = channels_pcm[0], channels_pcm[-1]
    for f in range(n // blocksize):
        sl = slice(f * blocksize, (f + 1) * blocksize)
        b.put(0x3FFE, 14); b.put(0, 1); b.put(0, 1)
        b.put(1, 4)                                                                  # block size code 1 = 192
        b.put(0, 4)                                                                  # sample rate: from STREAMINFO
        b.put(ch_code, 4); b.put(0, 3); b.put(0, 1)                                  # sample size: from STREAMINFO
        b.put(f, 8)                                                                  # frame number (< 128: one byte)
        b.put(0, 8)                                                                  # CRC-8 (not checked: the MD5 covers the samples)
        if ch_code == 8:                                                             # left, side
            subs = [(left[sl], bps), (left[sl] - right[sl], bps + 1)]
        elif ch_code == 9:                                                           # side, right
            subs = [(left[sl] - right[sl], bps + 1), (right[sl], bps)]
        elif ch_code == 10:                                                          # mid, side
            subs = [((left[sl] + right[sl]) >> 1, bps), (left[sl] - right[sl], bps + 1)]
        else:
            subs = [(ch[sl], bps) for ch in channels_pcm]
        for (x, sb), kind in zip(subs, kinds):
            _subframe(b, kind, x, sb, wasted)
        b.align()
        b.put(0, 16)                                                                 # CRC-16 (not checked)
    return bytes(b.out)


@pytest.mark.parametrize("kinds,ch_code,wasted", [
    (("verbatim", "constant"), 1, 0), (("fixed1", "fixed2"), 1, 0), (("fixed3", "fixed4"), 8, 0), (("fixed0", "verbatim"), 9, 0),
    (("fixed2", "fixed1"), 10, 0), (("verbatim", "fixed2"), 1, 3), (("fixed4",), 0, 0)])
def test_flac_subframe_kinds(kinds, ch_code, wasted):
    from whisper_b200.audio import decode_flac
    rng = np.random.default_rng(len(kinds) * 100 + ch_code * 10 + wasted)
    n = 384
    t = np.arange(n)
    chans = []
    for c, kind in enumerate(kinds):
        if kind == "constant":
            x = np.full(n, -1234)
        else:
            x = (3000 * np.sin(0.05 * t * (c + 1))).astype(np.int64) + rng.integers(-40, 40, n)
        chans.append((x >> wasted) << wasted)
    if ch_code >= 8:                                                                 # the subframe kinds apply to the decorrelated channels
        chans = [chans[0], chans[0] - (chans[1] >> 2)]
    data = _flac_stream(chans, 16, 22050, kinds, ch_code, wasted=wasted)
    pcm, rate, bits, md5 = decode_flac(data)
    assert (rate, bits) == (22050, 16) and pcm.shape == (n, len(chans))
    assert np.array_equal(pcm, np.stack(chans, axis=1))
    assert _pcm_md5(pcm, bits) == md5.hex()


def test_flac_rejects_garbage():
    from whisper_b200.audio import decode_flac
    with pytest.raises((ValueError, RuntimeError)):
        decode_flac(b"RIFF" + bytes(100))


@pytest.mark.gpu
def test_config0_tiny_greedy_jfk_flac():
    """BASELINE.json configs[0]: Whisper tiny (random init), greedy decode of tests/jfk.flac (one 30-s window).  load_audio()
    decodes the FLAC, mixes to mono, resamples 44.1 -> 16 kHz on the device and rounds to the int16 grid like the reference's
    `ffmpeg -f s16le` pipe; the oracle gets the numpy restatement of the same steps on the same decoded samples."""
    import torch
    from oracle import audio as oa, decoding as od, model as om
    from tests._util import close_library, exported
    from whisper_b200.audio import decode_flac, load_audio
    from whisper_b200.model import ModelDimensions, WhisperB200
    from whisper_b200.transcribe import transcribe
    with open(JFK, "rb") as f:
        pcm, rate, bits, _ = decode_flac(f.read())
    mono = pcm.astype(np.float64).mean(axis=1) / float(1 << (bits - 1))
    want_audio = oa.resample_poly(mono, rate)
    want_audio = np.clip(np.round(want_audio * 32768.0), -32768, 32767) / 32768.0
    got_audio = load_audio(JFK).cpu().numpy()
    assert got_audio.shape == want_audio.shape == (176000,)                          # 11.0 s at 16 kHz
    assert np.abs(got_audio - want_audio).max() <= 1.0 / 32768.0 + 1e-7              # at most one int16 step where the rounding flips
    assert (np.abs(got_audio - want_audio) > 1e-7).mean() < 1e-3
    dims, ckpt, folder = exported("tiny", 1, 0.03)
    close_library()
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        res = transcribe(m, torch.from_numpy(got_audio.astype(np.float32)), beam_size=None, sample_len=48)
    finally:
        m.close()
    assert res["seeks"] == [0] and res["windows"] == 1
    mel = oa.log_mel_spectrogram(torch.from_numpy(got_audio.astype(np.float32)), dims.n_mels, padding=480000)
    content = mel.shape[-1] - 3000
    seg = oa.pad_or_trim(mel[:, :content], 3000).contiguous()                        # zero-padded partial window (transcribe.py:286-290)
    want = od.decode_window(om.OracleModel(dims, ckpt), seg, od.Specials.load(dims.n_vocab), od.Options(sample_len=48, beam_size=None))
    got = [t for s in res["segments"] for t in s["tokens"]]
    n = len(got)
    assert n >= 3 and sum(a == b for a, b in zip(got, want.tokens[:n])) / n >= 0.99, (got, want.tokens)
