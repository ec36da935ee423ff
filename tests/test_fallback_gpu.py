"""Temperature sampling, fallback and the speculative seek.

  * Device sampler (whisper/decoding.py:307-313): the first sampled token over many seeds follows softmax(filtered logits / T) of
    the oracle; its reported log-probability is log_softmax of the UNPERTURBED filtered logits; a decode is reproducible from its
    seed, independent of how windows are batched, and different windows draw from different streams.
  * decode_with_fallback control flow (whisper/transcribe.py:188-228) and the no-speech skip (:309-322).
  * seek_mode="reference": the reference's seek chain reproduced with the windows decoded speculatively, on one rank and sharded
    over two ranks (gloo, both on cuda:0)."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import audio as oa, decoding as od, model as om, synth
from tests._util import close_library, exported

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(name, seed, scale):
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported(name, seed, scale)
    close_library()
    return dims, ckpt, WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()


def test_sampler_distribution_and_logprob():
    from whisper_b200.decoding import DecodingOptions, decode_windows
    dims, ckpt, m = _model("nano", 1, 0.03)
    try:
        mel = oa.log_mel_spectrogram(synth.noise_audio(1, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
        m.encode_windows(mel.cuda(), [0])
        sp = od.Specials.load(dims.n_vocab)
        orc = om.OracleModel(dims, ckpt)
        orc.encode(mel)
        toks = torch.tensor([list(sp.sot_sequence)])
        logits, _ = orc.logits(toks)
        lg = logits[:, -1].clone()
        od.apply_filters(lg, toks, sp, toks.shape[1], od.Options())
        T = 0.7
        p_ref = F.softmax(lg[0].double() / T, dim=-1).numpy()
        lp_plain = F.log_softmax(lg[0].float(), dim=-1).numpy()
        counts, n = {}, 600
        for seed in range(n):
            r = decode_windows(m, DecodingOptions(temperature=T, best_of=1, sample_len=1, seed=seed), [0])[0]
            assert len(r.tokens) <= 1
            if r.tokens:
                counts[r.tokens[0]] = counts.get(r.tokens[0], 0) + 1
                assert abs(r.sum_logprob - lp_plain[r.tokens[0]]) < 5e-2, (r.sum_logprob, lp_plain[r.tokens[0]])
        # (a sampled EOT leaves no token: the first step suppresses EOT, so every draw is a token)
        assert sum(counts.values()) == n
        assert all(p_ref[v] > 0 for v in counts)                             # nothing the logit filters masked is ever drawn
        for v in np.argsort(-p_ref)[:10]:                                     # multinomial counts within 4.5 sigma
            exp = p_ref[v] * n
            assert abs(counts.get(int(v), 0) - exp) <= 4.5 * math.sqrt(max(exp * (1 - p_ref[v]), 1.0)) + 2, (int(v), counts.get(int(v), 0), exp)
        # reproducible, and independent of batching: window 0 alone == window 0 inside a batch
        audio2 = torch.cat([synth.noise_audio(1, 480000), synth.noise_audio(2, 480000), synth.noise_audio(1, 480000)])
        mel2 = oa.log_mel_spectrogram(audio2, dims.n_mels, padding=480000)
        m.encode_windows(mel2.cuda(), [0, 3000, 6000])
        opts = DecodingOptions(temperature=0.9, best_of=5, sample_len=24, seed=7)
        a = decode_windows(m, opts, [0])[0]
        b = decode_windows(m, opts, [0, 1, 2])
        assert a.tokens == b[0].tokens and decode_windows(m, opts, [0])[0].tokens == a.tokens
        # windows 0 and 2 hold the same audio but draw from different streams (the window index keys the generator)
        assert b[0].tokens != b[2].tokens
        assert decode_windows(m, DecodingOptions(temperature=0.9, best_of=5, sample_len=24, seed=8), [0])[0].tokens != a.tokens
    finally:
        m.close()


def test_fallback_and_no_speech_skip():
    from whisper_b200.transcribe import transcribe
    dims, ckpt, m = _model("nano", 1, 0.03)
    try:
        audio = torch.cat([synth.noise_audio(10, 480000), synth.noise_audio(11, 480000)])
        base = transcribe(m, audio, beam_size=5, sample_len=24)
        assert all(s["temperature"] == 0.0 for s in base["segments"]) and len(base["decode_steps"]) == 2
        # a log-probability threshold nothing can reach: every temperature is tried, the last one's result is kept (:194-228)
        fb = transcribe(m, audio, beam_size=5, best_of=5, temperature=(0.0, 0.4, 0.8), logprob_threshold=10.0, sample_len=24)
        assert len(fb["decode_steps"]) == 6
        assert fb["segments"] and all(abs(s["temperature"] - 0.8) < 1e-6 for s in fb["segments"])
        # a threshold everything passes: no fallback, identical to the plain run
        ok = transcribe(m, audio, beam_size=5, best_of=5, temperature=(0.0, 0.4), logprob_threshold=-1e9, sample_len=24)
        assert [s["tokens"] for s in ok["segments"]] == [s["tokens"] for s in base["segments"]]
        # no-speech skip (:309-322): skipped unless the average log-probability clears the threshold
        skipped = transcribe(m, audio, beam_size=5, sample_len=24, no_speech_threshold=-1.0)
        assert skipped["segments"] == [] and skipped["windows"] == 2
        kept = transcribe(m, audio, beam_size=5, sample_len=24, no_speech_threshold=-1.0, logprob_threshold=-1e9)
        assert [s["tokens"] for s in kept["segments"]] == [s["tokens"] for s in base["segments"]]
    finally:
        m.close()


def _oracle_chain(dims, ckpt, audio, sample_len):
    from oracle import transcribe as otr
    orc = om.OracleModel(dims, ckpt)
    return otr.transcribe(orc, audio, od.Specials.load(dims.n_vocab), od.Options(sample_len=sample_len, beam_size=5))


def test_speculative_seek_single_rank():
    from whisper_b200.transcribe import transcribe
    dims, ckpt, m = _model("nano", 1, 0.03)
    try:
        audio = torch.cat([synth.noise_audio(10 + i, 480000) for i in range(4)])[:1600000]
        want = _oracle_chain(dims, ckpt, audio, 40)
        for batch in (1, 3, 8):
            got = transcribe(m, audio, beam_size=5, sample_len=40, seek_mode="reference", window_batch=batch)
            assert got["seeks"] == want["seeks"], (batch, got["seeks"], want["seeks"])
            assert [s["tokens"] for s in got["segments"]] == [s["tokens"] for s in want["segments"]]
    finally:
        m.close()





def test_speculative_seek_two_ranks():
    """the windows of every speculative round are split over two ranks (both on cuda:0), results exchanged with gloo; both ranks
    must arrive at the reference's seeks and segments (tests/golden/ref_transcribe.npz pins the oracle's chain to the reference)"""
    import json
    dims, ckpt, folder = exported("nano", 1, 0.03)
    audio = torch.cat([synth.noise_audio(10 + i, 480000) for i in range(4)])[:1600000]
    want = _oracle_chain(dims, ckpt, audio, 40)
    close_library()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tests", "_rank_reference_seek.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
    assert len(lines) == 2, p.stdout[-2000:]
    for l in lines:
        got = json.loads(l[7:])
        assert got["seeks"] == want["seeks"], (got["seeks"], want["seeks"])
        assert got["tokens"] == [s["tokens"] for s in want["segments"]]
