"""Run the REFERENCE's transcribe() (imported from /root/reference) over a scripted decoder and record what it returns:
tests/golden/ref_fallback.json pins the temperature ladder (transcribe.py:188-228), the no-speech skip (:309-322), the segment
slicing and the data-dependent seek (:350-409), and the clearing of instantaneous segments (:495-500) of the reference itself.

    PYTHONPATH=/root/reference:/root/repo python tests/golden/make_fallback_golden.py"""
import json
import os
import sys

import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import whisper                                                              # noqa: E402
from whisper.decoding import DecodingResult                                 # noqa: E402
from tests import _fallback_script as fs                                    # noqa: E402
from oracle import synth                                                    # noqa: E402


class ScriptedModel:
    """The attributes whisper.transcribe() touches (transcribe.py:132-242) and a decode() that follows the script."""
    device = torch.device("cpu")
    is_multilingual, num_languages = True, 99
    dims = whisper.model.ModelDimensions(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4, n_vocab=51865,
                                         n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4)

    def __init__(self, scenario, mel):
        self.scenario, self.mel, self.calls = scenario, mel, []

    def decode(self, segment, options):
        # which window is this?  the segment is mel[:, seek : seek + size] zero-padded: find the seek whose first frames match
        probe = segment[:, :40]
        seek = next(s for s in range(0, self.mel.shape[-1] - 40) if torch.equal(self.mel[:, s:s + 40].to(probe.dtype), probe))
        ti = fs.TEMPERATURES.index(round(float(options.temperature), 3))
        r = fs.scripted_result(self.scenario, seek, ti)
        self.calls.append([seek, ti])
        text = fs.fake_text(r["tokens"])
        return DecodingResult(audio_features=None, language="en", tokens=list(r["tokens"]), text=text, avg_logprob=r["avg_logprob"],
                              no_speech_prob=r["no_speech_prob"], temperature=float(options.temperature),
                              compression_ratio=fs.compression_ratio(text))


def main():
    out = {"temperatures": list(fs.TEMPERATURES), "seconds": 170, "scenarios": []}
    audio = synth.noise_audio(3, 170 * 16000)
    mel = whisper.log_mel_spectrogram(audio, 80, padding=480000)
    for scenario in range(12):
        m = ScriptedModel(scenario, mel)
        res = whisper.transcribe(m, audio, temperature=fs.TEMPERATURES, compression_ratio_threshold=2.4, logprob_threshold=-1.0,
                                 no_speech_threshold=0.6, condition_on_previous_text=False, language="en", fp16=False, verbose=None)
        out["scenarios"].append({"scenario": scenario, "calls": m.calls,
                                 "segments": [{"seek": s["seek"], "start": s["start"], "end": s["end"], "tokens": s["tokens"],
                                               "temperature": s["temperature"]} for s in res["segments"]]})
        print(scenario, len(m.calls), "decode calls,", len(res["segments"]), "segments")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_fallback.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()
