"""The host logic of transcribe() - temperature fallback (whisper/transcribe.py:188-228), no-speech skip (:309-322), segment
slicing and the data-dependent seek (:350-409), clearing of instantaneous segments (:495-500) - against THE REFERENCE ITSELF, on the
CPU: tests/golden/ref_fallback.json holds what whisper.transcribe() of /root/reference returned over a scripted decoder
(tests/_fallback_script.py, tests/golden/make_fallback_golden.py); here this repo's transcribe(seek_mode="reference") runs over the
same script through a scripted backend (no device: the decode, the encoder and the log-mel are replaced, nothing else is)."""
import json
import os

import pytest
import torch

from oracle import audio as oa, synth
from tests import _fallback_script as fs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fallback.json")


class _Tokenizer:
    def decode(self, tokens):
        return fs.fake_text(tokens)


class _Backend:
    """What transcribe() asks of a WhisperB200: dims, specials, encode_windows / select_window; `device` routes the audio to the CPU."""
    device = "cpu"

    def __init__(self):
        from whisper_b200.model import ModelDimensions, Specials
        self.dims = ModelDimensions(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4, n_vocab=51865,
                                    n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)
        self.specials = Specials.load(51865)
        self.seeks, self.calls = [], []

    def load(self):
        return self

    def encode_windows(self, mel, seeks, content_frames=None):
        self.seeks = list(seeks)

    def select_window(self, w):
        pass


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.mark.parametrize("scenario", range(12))
@pytest.mark.parametrize("window_batch", [1, 3])
def test_ladder_skip_and_seek_match_the_reference(golden, scenario, window_batch, monkeypatch):
    from whisper_b200 import transcribe as tr
    from whisper_b200.decoding import DecodingResult
    want = golden["scenarios"][scenario]
    assert want["scenario"] == scenario and tuple(golden["temperatures"]) == fs.TEMPERATURES
    backend = _Backend()

    def scripted_decode_windows(model, opts, pending):
        ti = fs.TEMPERATURES.index(round(float(opts.temperature), 3))
        assert (opts.beam_size is None) == (ti > 0)                      # beam search at T = 0, best_of sampling above (:193-199)
        out = []
        for i in pending:
            r = fs.scripted_result(scenario, model.seeks[i], ti)
            model.calls.append((model.seeks[i], ti))
            out.append(DecodingResult(tokens=list(r["tokens"]), avg_logprob=r["avg_logprob"], no_speech_prob=r["no_speech_prob"],
                                      temperature=float(opts.temperature), steps=1))
        return out

    monkeypatch.setattr(tr, "decode_windows", scripted_decode_windows)
    monkeypatch.setattr(tr, "log_mel_spectrogram", lambda audio, n_mels, padding=0: oa.log_mel_spectrogram(audio, n_mels, padding=padding))
    audio = synth.noise_audio(3, golden["seconds"] * 16000)
    res = tr.transcribe(backend, audio, beam_size=5, best_of=5, temperature=fs.TEMPERATURES, compression_ratio_threshold=2.4,
                        logprob_threshold=-1.0, no_speech_threshold=0.6, seek_mode="reference", window_batch=window_batch,
                        tokenizer=_Tokenizer())
    got = [(s["seek"], round(s["start"], 6), round(s["end"], 6), s["tokens"], round(s["temperature"], 6)) for s in res["segments"]]
    ref = [(s["seek"], round(s["start"], 6), round(s["end"], 6), s["tokens"], round(s["temperature"], 6)) for s in want["segments"]]
    assert got == ref
    # the windows visited, in order (the reference decodes each once per temperature it needs; this repo decodes a speculative grid
    # and keeps the windows whose start the seek rule confirms)
    visited = []
    for seek, _ in want["calls"]:
        if not visited or visited[-1] != seek:
            visited.append(seek)
    assert res["seeks"] == visited
    # and every (window, temperature) the reference needed was decoded here with the same script
    assert set(map(tuple, want["calls"])) <= set(backend.calls)
