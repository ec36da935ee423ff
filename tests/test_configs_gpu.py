"""The BASELINE.json configurations as parity cases (the bench measures only configs[3] at full length):
  configs[1]  base, beam_size=5, language en, one 30-s window
  configs[2]  small, word_timestamps=True, fixed 30-s windows of a longer clip (2 windows here)
  configs[3]  large-v3-turbo dims (32 enc / 4 dec layers, 128 mels), beam_size=5, encoder + crossKV + decoder256 + decoder1
Each compares the CUDA path (through the C ABI) with the CPU oracle on the same seeded weights and audio.
Tolerances (north_star): encoder / logits relative error <= 2e-2, token sequences >= 99% identical, DTW bit-exact."""
import numpy as np
import pytest
import torch

from oracle import audio as oa, decoding as od, model as om, synth, timing as ot
from tests._util import exported, rel

pytestmark = pytest.mark.gpu


def _agreement(a, b):
    return sum(x == y for x, y in zip(a, b)) / max(len(a), len(b), 1)


def _model(name, seed=0, scale=1.0):
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported(name, seed, scale)
    return dims, ckpt, WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()


def test_config1_base_beam5_single_window():
    from whisper_b200.decoding import DecodingOptions, decode
    dims, ckpt, m = _model("base")
    mel = oa.log_mel_spectrogram(synth.noise_audio(1, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
    m.encode_windows(mel.cuda(), [0])
    got = decode(m, DecodingOptions(beam_size=5, sample_len=32, language="en"), window=0)
    want = od.decode_window(om.OracleModel(dims, ckpt), mel, od.Specials.load(dims.n_vocab), od.Options(sample_len=32, beam_size=5))
    assert _agreement(got.tokens, want.tokens) >= 0.99, (got.tokens, want.tokens)
    assert got.steps == want.steps
    assert abs(got.avg_logprob - want.avg_logprob) <= 2e-2 * max(1.0, abs(want.avg_logprob))
    m.close()


def test_config2_small_word_timestamps_two_windows():
    from whisper_b200.transcribe import transcribe
    dims, ckpt, m = _model("small")                    # reference-default init (near-uniform logits would let bf16 flip near-ties)
    audio = torch.cat([synth.noise_audio(1, 480000), synth.noise_audio(2, 480000)])
    res = transcribe(m, audio, beam_size=5, word_timestamps=True, sample_len=20)
    assert res["windows"] == 2 and res["seeks"] == [0, 3000]
    orc = om.OracleModel(dims, ckpt)
    sp = od.Specials.load(dims.n_vocab)
    mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
    for w, seek in enumerate(res["seeks"]):
        want = od.decode_window(orc, mel[:, seek:seek + 3000].contiguous(), sp, od.Options(sample_len=20, beam_size=5))
        segs = [s for s in res["segments"] if s["seek"] == seek]
        got = [t for s in segs for t in s["tokens"]]
        assert _agreement(got, want.tokens) >= 0.99, (w, got, want.tokens)
        for s in segs:                                 # timing.py:268-376: one word entry per text token, times inside the window
            text = [t for t in s["tokens"] if t < sp.eot]
            assert len(s["words"]) == len(text)
            t0 = seek * 0.01
            prev = t0
            for wd in s["words"]:
                assert t0 - 1e-6 <= wd["start"] <= wd["end"] <= t0 + 30.0 + 1e-6
                assert wd["start"] >= prev - 1e-6       # DTW paths are monotone
                prev = wd["start"]
                assert 0.0 <= wd["probability"] <= 1.0
    m.close()


def test_config3_turbo_split_stages():
    """turbo dims: encoder output, cross K/V and an 8-step beam-5 decode (prompt + decoder1 steps) against the oracle."""
    import ctypes
    from whisper_b200 import _lib
    from whisper_b200.decoding import DecodingOptions, decode
    dims, ckpt, m = _model("turbo")
    orc = om.OracleModel(dims, ckpt)
    mel = oa.log_mel_spectrogram(synth.noise_audio(1, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
    m.encode_windows(mel.cuda(), [0])
    d, Ld, H = dims.n_text_state, dims.n_text_layer, dims.n_text_head
    xa = torch.empty(1500, d)
    m.lib.b200TestGetXa(ctypes.cast(xa.data_ptr(), _lib.f32p), 0)
    xa_ref = orc.encode(mel)
    assert rel(xa, xa_ref) < 2e-2, rel(xa, xa_ref)
    ck = torch.empty(Ld, H, 64, 1500); cv = torch.empty(Ld, H, 1500, 64)
    m.lib.b200TestGetCrossKV(ctypes.cast(ck.data_ptr(), _lib.f32p), ctypes.cast(cv.data_ptr(), _lib.f32p), 0)
    ck_ref, cv_ref = om.cross_kv(orc.w, dims, xa_ref)
    assert rel(ck, ck_ref) < 2e-2 and rel(cv, cv_ref) < 2e-2
    sp = od.Specials.load(dims.n_vocab)
    for beam in (None, 5):
        got = decode(m, DecodingOptions(beam_size=beam, sample_len=8), window=0)
        want = od.decode_window(orc, mel, sp, od.Options(sample_len=8, beam_size=beam))
        assert _agreement(got.tokens, want.tokens) >= 0.99, (beam, got.tokens, want.tokens)
        assert abs(got.sum_logprob - want.sum_logprob) <= 2e-2 * max(1.0, abs(want.sum_logprob)), (got.sum_logprob, want.sum_logprob)
    m.close()


def test_transcribe_reference_seek_and_partial_last_window():
    """transcribe(seek_mode="reference") against the oracle's restatement of the reference loop (itself pinned to
    whisper.transcribe() by tests/golden/ref_transcribe.npz): same data-dependent seeks, same tokens; the last window is partial
    (1024 content frames), so the zero padding of transcribe.py:286-290 is exercised too.  The fixed-window mode is checked on the
    same clip: its last window (10 s of content) must match the oracle on the zero-padded mel, not on the padded file's mel."""
    from oracle import transcribe as otr
    from whisper_b200.transcribe import transcribe
    dims, ckpt, m = _model("nano", 1, 0.03)
    orc = om.OracleModel(dims, ckpt)
    sp = od.Specials.load(dims.n_vocab)
    audio = torch.cat([synth.noise_audio(10 + i, 480000) for i in range(4)])[:1600000]
    want = otr.transcribe(orc, audio, sp, od.Options(sample_len=40, beam_size=5))
    got = transcribe(m, audio, beam_size=5, sample_len=40, seek_mode="reference")
    assert got["seeks"] == want["seeks"], (got["seeks"], want["seeks"])
    a = [t for s in got["segments"] for t in s["tokens"]]
    b = [t for s in want["segments"] for t in s["tokens"]]
    assert _agreement(a, b) >= 0.99, (a, b)
    assert [s["seek"] for s in got["segments"]] == [s["seek"] for s in want["segments"]]
    # fixed windows: 0, 3000, 6000, 9000 (the last one holds 1000 content frames)
    fixed = transcribe(m, audio, beam_size=5, sample_len=40)
    assert fixed["seeks"] == [0, 3000, 6000, 9000]
    mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
    content = mel.shape[-1] - 3000
    seg = oa.pad_or_trim(mel[:, 9000:content], 3000).contiguous()
    last = od.decode_window(orc, seg, sp, od.Options(sample_len=40, beam_size=5))
    got_last = [t for s in fixed["segments"] if s["seek"] == 9000 for t in s["tokens"]]      # the finished segments: a prefix of the decode
    assert len(got_last) >= 3 and _agreement(got_last, last.tokens[:len(got_last)]) >= 0.99, (got_last, last.tokens)
    unpadded = od.decode_window(orc, mel[:, 9000:12000].contiguous(), sp, od.Options(sample_len=40, beam_size=5))
    assert unpadded.tokens != last.tokens          # the case distinguishes zero padding from the padded file's silence frames
    m.close()


@pytest.mark.parametrize("tile", [1, 3])
def test_encoder_parity_with_forced_gemm_tiles(tile):
    """The whole encoder + crossKV (conv stem with its three row-shifted A maps and window batches, residual epilogues, second
    fragment-major output) through the 128x128 and the CTA-pair (cta_group::2, 256x256) GEMM kernels - the automatic dispatch only
    picks the pair kernel from four windows of a large model up, which no other test reaches."""
    import ctypes
    from whisper_b200 import _lib
    dims, ckpt, m = _model("tiny")
    orc = om.OracleModel(dims, ckpt)
    audio = torch.cat([synth.noise_audio(3, 480000), synth.noise_audio(4, 480000)])
    mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
    m.lib.b200TestGemmTile(tile)
    try:
        m.encode_windows(mel.cuda(), [0, 3000])
    finally:
        m.lib.b200TestGemmTile(0)
    d, Ld, H = dims.n_text_state, dims.n_text_layer, dims.n_text_head
    for w, seek in enumerate([0, 3000]):
        xa = torch.empty(1500, d)
        m.lib.b200TestGetXa(ctypes.cast(xa.data_ptr(), _lib.f32p), w)
        xa_ref = orc.encode(mel[:, seek:seek + 3000].contiguous())
        assert rel(xa, xa_ref) < 2e-2, (tile, w, rel(xa, xa_ref))
        ck = torch.empty(Ld, H, 64, 1500); cv = torch.empty(Ld, H, 1500, 64)
        m.lib.b200TestGetCrossKV(ctypes.cast(ck.data_ptr(), _lib.f32p), ctypes.cast(cv.data_ptr(), _lib.f32p), w)
        ck_ref, cv_ref = om.cross_kv(orc.w, dims, xa_ref)
        assert rel(ck, ck_ref) < 2e-2 and rel(cv, cv_ref) < 2e-2, (tile, w, rel(ck, ck_ref), rel(cv, cv_ref))
    m.close()


def test_transcribe_edge_inputs():
    """Ragged / degenerate inputs through transcribe(): a clip shorter than 1 s yields no window (transcribe.py:295-298), exactly
    30 s yields one, digital silence decodes like the oracle (the log-mel is then a constant: audio.py:152-156)."""
    from whisper_b200.transcribe import transcribe
    dims, ckpt, m = _model("nano", 1, 0.03)
    orc = om.OracleModel(dims, ckpt)
    sp = od.Specials.load(dims.n_vocab)
    for mode in ("fixed", "reference"):
        short = transcribe(m, synth.noise_audio(7, 8000), beam_size=5, sample_len=12, seek_mode=mode)
        assert short["windows"] == 0 and short["segments"] == [] and short["seeks"] == []
        one = transcribe(m, synth.noise_audio(8, 480000), beam_size=5, sample_len=12, seek_mode=mode)
        assert one["seeks"] == [0] and one["windows"] == 1
    silence = torch.zeros(480000)
    got = transcribe(m, silence, beam_size=None, sample_len=12)                    # greedy
    mel = oa.log_mel_spectrogram(silence, dims.n_mels, padding=480000)
    want = od.decode_window(orc, oa.pad_or_trim(mel[:, :3000], 3000).contiguous(), sp, od.Options(sample_len=12, beam_size=None))
    toks = [t for s in got["segments"] for t in s["tokens"]]
    assert len(toks) >= 1 and _agreement(toks, want.tokens[:len(toks)]) >= 0.99, (toks, want.tokens)
    m.close()
