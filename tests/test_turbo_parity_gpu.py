"""Parity at the dimensions the bench times (turbo: d = 1280, 4 decoder layers, V = 51866) and at large-v3 dims, against the CPU
oracle on the same seeded weights and audio, through the C ABI.  Tolerances (BASELINE.json north_star): logits / encoder output
relative error <= 2e-2, greedy / beam token sequences >= 99 % identical.

  * decoder256Predict / decoder1Predict logits at d = 1280 with "soft" (token-embedding scale 0.03, i.e. non-degenerate) weights,
    six decoder1 steps with beam permutations between them (whisper/decoder.py:241-257, 261-329; coreml.mm:245-277, 404-444);
  * a FULL-LENGTH (sample_len 224) greedy and beam-5 decode of one window, default and soft weights;
  * one large-v3-dims window: Xa, CK / CV, 8-step tokens;
  * the entry points nothing else exercises: decoder1Predict with one beam slot and the (1, 450) mask (decoder.py:246-248),
    decoder1StepFused, b200DecodeWindow."""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import audio as oa, decoding as od, model as om, synth
from tests._util import exported, rel

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _mel(dims, seed=1):
    return oa.log_mel_spectrogram(synth.noise_audio(seed, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()


def _f32p(t):
    from whisper_b200 import _lib
    return ctypes.cast(t.data_ptr(), _lib.f32p)


def _agreement(a, b):
    return sum(x == y for x, y in zip(a, b)) / max(len(a), len(b), 1)


def _backend(dims, folder, bs, n_align):
    from whisper_b200 import b200
    be = b200.B200(dims.n_audio_layer, dims.n_text_layer, dims.n_mels, dims.n_audio_state, dims.n_audio_head, dims.n_vocab, folder)
    be.bs = bs
    be.n_alignment_head = n_align
    return be


@pytest.fixture(scope="module")
def turbo_soft():
    """turbo dims, token embedding scale 0.03: logits of a few units, so the softmax is neither flat nor a one-hot"""
    dims, ckpt, folder = exported("turbo", 1, 0.03)
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims)
    xa_ref = orc.encode(mel)
    return dims, ckpt, folder, orc, mel, xa_ref


def test_turbo_soft_abi_logits(turbo_soft):
    """decoder256Predict + six decoder1Predict steps with beam permutations at d = 1280: logits relative error <= 2e-2."""
    from whisper_b200 import _lib
    dims, ckpt, folder, orc, mel, xa_ref = turbo_soft
    d, Ld = dims.n_text_state, dims.n_text_layer
    be = _backend(dims, folder, 5, len(orc.heads))
    lib = _lib.load()
    try:
        be.encoderPredict(mel[None])
        xa = torch.empty(1500, d)
        lib.b200TestGetXa(_f32p(xa), 0)
        assert rel(xa, xa_ref) < TOL, rel(xa, xa_ref)
        be.crossKVPredict()
        sp = od.Specials.load(dims.n_vocab)
        toks = torch.tensor([list(sp.sot_sequence)] * 5)
        n = toks.shape[1]
        orc.reset()
        lg_ref, chw_ref = orc.logits(toks)
        x = torch.cat([orc.embed(toks, 0), torch.zeros(5, 256 - n, d)], dim=1)
        mask = om.prefill_mask(n)
        for b in range(5):
            out_x, out_chw, _ = be.decoder256Predict(x[b:b + 1], mask, b)
            if b == 0:
                lg = out_x[0, :n] @ orc.w["decoder.token_embedding.weight"].t()     # decoder.py:238-240 (host side)
                assert rel(lg, lg_ref[0]) < TOL, rel(lg, lg_ref[0])
                assert rel(out_chw[:, :n], chw_ref) < TOL, rel(out_chw[:, :n], chw_ref)
        be.loadDecoder1()
        tokens = toks
        g = torch.Generator().manual_seed(5)
        for step in range(6):
            nxt = torch.randint(0, 50257, (5, 1), generator=g)
            tokens = torch.cat([tokens, nxt], dim=1)
            t = orc.text_offset
            lg_ref, _ = orc.logits(tokens)
            logits, _ = be.decoder1Predict(orc.embed(tokens[:, -1:], t), om.step_mask(t), t)
            assert rel(logits[:, 0], lg_ref[:, 0]) < TOL, (step, rel(logits[:, 0], lg_ref[:, 0]))
            # log-probabilities (what sampling sees): absolute error small against a spread of several units
            lp, lp_ref = F.log_softmax(logits[:, 0], -1), F.log_softmax(lg_ref[:, 0], -1)
            assert float((lp - lp_ref).abs().max()) < 5e-2, (step, float((lp - lp_ref).abs().max()))
            src = torch.randint(0, 5, (5,), generator=g).tolist()
            orc.rearrange_kv_cache(src)
            tokens = tokens[src]
            be.rearrange_mkv(src, orc.text_offset)
        t = orc.text_offset
        kv = torch.empty(2 * Ld, 5, t, d)
        lib.b200TestGetKV(_f32p(kv), t)
        assert rel(kv, orc.mkv[:, :, :t]) < TOL, rel(kv, orc.mkv[:, :, :t])
    finally:
        be.close()
        orc.reset()


def _first_difference(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), n if len(a) == len(b) else n)


@pytest.mark.parametrize("scale,seed", [(1.0, 0), (0.03, 1)])
@pytest.mark.parametrize("beam", [None, 5])
def test_turbo_full_length_tokens(scale, seed, beam):
    """One window, sample_len = 224 (the bench's length): token sequence >= 99 % identical to the oracle's."""
    from whisper_b200.decoding import DecodingOptions, decode
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("turbo", seed, scale)
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims)
    sp = od.Specials.load(dims.n_vocab)
    want = od.decode_window(orc, mel, sp, od.Options(sample_len=224, beam_size=beam))
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        m.encode_windows(mel.cuda(), [0])
        got = decode(m, DecodingOptions(sample_len=224, beam_size=beam), window=0)
    finally:
        m.close()
    k = _first_difference(got.tokens, want.tokens)
    assert _agreement(got.tokens, want.tokens) >= 0.99, (scale, beam, k, len(got.tokens), len(want.tokens), got.tokens[k:k + 4], want.tokens[k:k + 4])
    assert got.steps == want.steps
    assert abs(got.sum_logprob - want.sum_logprob) <= 2e-2 * max(1.0, abs(want.sum_logprob)), (got.sum_logprob, want.sum_logprob)


def test_large_v3_dims_one_window():
    """BASELINE configs[4] dims (32 + 32 layers, d = 1280, 128 mels): Xa, CK / CV, 8-step greedy and beam-5 tokens of one window."""
    from whisper_b200 import _lib
    from whisper_b200.decoding import DecodingOptions, decode
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("large-v3", 0, 1.0)
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims, 2)
    xa_ref = orc.encode(mel)
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        m.encode_windows(mel.cuda(), [0])
        d, Ld, H = dims.n_text_state, dims.n_text_layer, dims.n_text_head
        xa = torch.empty(1500, d)
        m.lib.b200TestGetXa(_f32p(xa), 0)
        assert rel(xa, xa_ref) < TOL, rel(xa, xa_ref)
        ck = torch.empty(Ld, H, 64, 1500); cv = torch.empty(Ld, H, 1500, 64)
        m.lib.b200TestGetCrossKV(_f32p(ck), _f32p(cv), 0)
        ck_ref, cv_ref = om.cross_kv(orc.w, dims, xa_ref)
        assert rel(ck, ck_ref) < TOL and rel(cv, cv_ref) < TOL, (rel(ck, ck_ref), rel(cv, cv_ref))
        sp = od.Specials.load(dims.n_vocab)
        for beam in (None, 5):
            got = decode(m, DecodingOptions(beam_size=beam, sample_len=8), window=0)
            want = od.decode_window(orc, None, sp, od.Options(sample_len=8, beam_size=beam))
            assert _agreement(got.tokens, want.tokens) >= 0.99, (beam, got.tokens, want.tokens)
            # eight log-probabilities of ~-0.015 each, from logits of magnitude ~40 after 32 bf16 layers: 2e-2 relative on the
            # logits would be ~0.8 per step; the sums must agree to 0.05
            assert abs(got.sum_logprob - want.sum_logprob) <= 5e-2 * max(1.0, abs(want.sum_logprob)), (got.sum_logprob, want.sum_logprob)
    finally:
        m.close()


def test_decoder1_single_beam_450_mask():
    """One beam slot: the step mask grows to (1, 450) (whisper/decoder.py:246-248); the extra column is -inf and must be ignored."""
    dims, ckpt, folder = exported("tiny", 0, 1.0)
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims)
    orc.encode(mel)
    d = dims.n_text_state
    be = _backend(dims, folder, 1, len(orc.heads))
    try:
        be.encoderPredict(mel[None])
        be.crossKVPredict()
        sp = od.Specials.load(dims.n_vocab)
        toks = torch.tensor([list(sp.sot_sequence)])
        n = toks.shape[1]
        orc.reset()
        orc.logits(toks)
        x = torch.cat([orc.embed(toks, 0), torch.zeros(1, 256 - n, d)], dim=1)
        be.decoder256Predict(x, om.prefill_mask(n), 0)
        be.loadDecoder1()
        tokens = toks
        g = torch.Generator().manual_seed(9)
        for step in range(4):
            tokens = torch.cat([tokens, torch.randint(0, 50257, (1, 1), generator=g)], dim=1)
            t = orc.text_offset
            lg_ref, _ = orc.logits(tokens)
            mask = torch.cat([om.step_mask(t), torch.full((1, 1), -math.inf)], dim=1)
            assert mask.shape == (1, 450)
            logits, _ = be.decoder1Predict(orc.embed(tokens[:, -1:], t), mask, t)
            assert logits.shape == (1, 1, dims.n_vocab)
            assert rel(logits[0, 0], lg_ref[0, 0]) < TOL, (step, rel(logits[0, 0], lg_ref[0, 0]))
    finally:
        be.close()
        orc.reset()


def test_decoder1_step_fused_and_decode_window():
    """decoder1StepFused: one token step + logit filters + log-softmax + top-(bs + 1) per beam on the device, against the oracle's
    filters (decoding.py:450-532) and topk (:366-369).  b200DecodeWindow: the single-window form of b200DecodeWindows."""
    from whisper_b200 import _lib
    from whisper_b200.decoding import DecodingOptions, decode
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("tiny", 0, 1.0)
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims)
    sp = od.Specials.load(dims.n_vocab)
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        m.encode_windows(mel.cuda(), [0])
        # a 6-step beam decode leaves the KV cache / slot tables of 5 live beams on the device; replay it on the oracle
        n0 = len(sp.sot_sequence)
        got = decode(m, DecodingOptions(beam_size=5, sample_len=6), window=0)
        want = od.decode_window(orc, mel, sp, od.Options(sample_len=6, beam_size=5))
        assert got.tokens == want.tokens
        # b200DecodeWindow (current window) == decode()
        init = np.array(sp.sot_sequence, dtype=np.int32)
        toks = np.empty((5, 449), dtype=np.int32); lens = np.empty(5, dtype=np.int32); lps = np.empty(5, dtype=np.float32)
        nsp = np.empty(1, dtype=np.float32)
        m.select_window(0)
        steps = m.lib.b200DecodeWindow(init.ctypes.data_as(_lib.i32p), len(init), 5, 6, 0, 50, toks.ctypes.data_as(_lib.i32p),
                                       lens.ctypes.data_as(_lib.i32p), lps.ctypes.data_as(_lib.f32p), nsp.ctypes.data_as(_lib.f32p))
        _lib.check_errors("b200DecodeWindow")
        assert steps == got.steps
        cands = [toks[i, n0:n0 + lens[i]].tolist() for i in range(5) if lens[i] >= 0]
        assert got.tokens in cands
        assert abs(float(nsp[0]) - got.no_speech_prob) < 1e-6
        # decoder1StepFused on a fresh prefill: history = the sot sequence + one timestamp token for every beam
        orc.reset()
        orc.encode(mel)
        hist = torch.tensor([list(sp.sot_sequence) + [sp.timestamp_begin]] * 5)
        orc.logits(hist[:, :n0])                                           # prefill
        be = m.backend
        x = torch.cat([orc.embed(hist[:, :n0], 0), torch.zeros(5, 256 - n0, dims.n_text_state)], dim=1)
        for b in range(5):
            be.decoder256Predict(x[b:b + 1], om.prefill_mask(n0), b)
        lg_ref, _ = orc.logits(hist)                                        # the step that consumes the timestamp token
        lg = lg_ref[:, -1].clone()
        od.apply_filters(lg, hist, sp, n0, od.Options(beam_size=5))
        lp_ref = F.log_softmax(lg.float(), dim=-1)
        vals, idx = lp_ref.topk(6, dim=-1)
        h = hist.to(torch.int32).contiguous().numpy()
        out_lp = np.empty((5, 6), dtype=np.float32); out_tok = np.empty((5, 6), dtype=np.int32)
        m.lib.decoder1StepFused(h.ctypes.data_as(_lib.i32p), h.shape[1], n0, h.shape[1] - 1, 0, 50,
                                out_lp.ctypes.data_as(_lib.f32p), out_tok.ctypes.data_as(_lib.i32p))
        _lib.check_errors("decoder1StepFused")
        for b in range(5):
            assert out_tok[b].tolist() == idx[b].tolist(), (b, out_tok[b], idx[b])
            # log-probabilities of logits that span ~40 units: 2e-2 relative on the logits is ~0.1 here
            assert np.allclose(out_lp[b], vals[b].numpy(), rtol=2e-2, atol=0.1), (b, out_lp[b], vals[b])
    finally:
        m.close()
