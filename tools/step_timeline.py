"""Per-stage timeline (one CTA) of the persistent decoder step kernel on turbo dims (GPU)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import weights_folder
from oracle import model as om, synth
from whisper_b200.model import ModelDimensions, WhisperB200
from whisper_b200.audio import log_mel_spectrogram
from whisper_b200.decoding import DecodingOptions, decode
name = sys.argv[1] if len(sys.argv) > 1 else "turbo"
dims, folder, _ = weights_folder(name, 0)
m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
mel = log_mel_spectrogram(synth.noise_audio(1, 480000).cuda(), dims.n_mels, padding=480000)
m.encode_windows(mel, [0])
decode(m, DecodingOptions(beam_size=5, sample_len=40), window=0)      # warm
m.lib.b200TestStepTimeline(1, None, 0)
decode(m, DecodingOptions(beam_size=5, sample_len=5), window=0)
buf = np.zeros(2048, dtype=np.uint64)
n = m.lib.b200TestStepTimeline(0, buf.ctypes.data_as(ctypes.c_void_p), 2048)
t = buf[:n].astype(np.int64)
# marks per step: start; per layer: stages 0,2,3,5,6,7 have (prologue, units, barrier)=3 marks, stages 1,4 have (units, barrier)=2
names = []
for l in range(dims.n_text_layer):
    for st, has_pro in (("qkv", 1), ("self_attn", 0), ("out", 1), ("cross_q", 1), ("cross_attn", 0), ("cross_out", 1), ("mlp1", 1), ("mlp2", 1)):
        if has_pro: names.append(f"L{l}.{st}.prologue")
        names += [f"L{l}.{st}.units", f"L{l}.{st}.barrier"]
names += ["vocab.prologue", "vocab.units", "vocab.mark2", "vocab.barrier", "sample.units", "sample.barrier"]
per = 1 + len(names)
print("marks", n, "per step", per)
s0 = 2 * per
seg = t[s0:s0 + per]
d = np.diff(seg) / 1000.0
for k, v in zip(names, d):
    print(f"{k:28s} {v:8.2f} us")
print("step total", (seg[-1] - seg[0]) / 1000.0, "us")
