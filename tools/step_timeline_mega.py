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
LD = 640
buf = np.zeros(256 * LD, dtype=np.uint64)
n = m.lib.b200TestStepTimeline(0, buf.ctypes.data_as(ctypes.c_void_p), 256)
T = buf[:n * LD].astype(np.int64).reshape(n, LD)
T = T[T[:, 0] > 0]                      # CTAs that ran (B200_MEGA_CTAS < n_sms leaves the rest empty)
n = T.shape[0]
STAGES = ["qkv", "self_attn", "out_proj", "cross_q", "cross_attn", "cross_out", "mlp1", "mlp2"]
n_stages = dims.n_text_layer * 8 + 1
t0 = T[:, 0].min()
print(f"{n} CTAs; kernel start skew {(T[:, 0].max() - t0) / 1000:.2f} us")
print(f"{'stage':22s} {'first in':>9s} {'last in':>9s} | {'first ready':>11s} {'last ready':>10s} | {'first done':>10s} {'last done':>9s}   (us since start; in = entered, ready = prologue done)")
prev_done = T[:, 0]
for it in range(n_stages):
    name = f"L{it // 8}.{STAGES[it % 8]}" if it < n_stages - 1 else "vocab"
    ready, done = T[:, 2 * it + 1], T[:, 2 * it + 2]
    f = lambda a: (a - t0) / 1000.0
    print(f"{name:22s} {f(prev_done.min()):9.2f} {f(prev_done.max()):9.2f} | {f(ready.min()):11.2f} {f(ready.max()):10.2f} | {f(done.min()):10.2f} {f(done.max()):9.2f}")
    prev_done = done
tail = [("barrier", 2 * n_stages + 1), ("sample_partial", 2 * n_stages + 2), ("barrier", 2 * n_stages + 3), ("beam_update (CTA 0)", 2 * n_stages + 4)]
prev = T[:, 2 * n_stages]
for nm, k in tail:
    cur = T[:, k]
    ok = cur > 0
    if not ok.any(): break
    print(f"{nm:22s} {(prev[prev > 0].max() - t0) / 1000:9.2f} -> first {(cur[ok].min() - t0) / 1000:9.2f} last {(cur[ok].max() - t0) / 1000:9.2f}")
    prev = cur
print("step total", (T[:, :2 * n_stages + 5].max() - t0) / 1000.0, "us")
names = {1: "sync", 2: "sentinels", 11: "staged+stats", 12: "gamma/beta+sync", 16: "normalised", 17: "sync"}
for base, nm in ((200, "L0.qkv LN (embed)"), (220, "L0.mlp1 LN (LL)")):
    c = 100 % n
    prev = None
    out = []
    for k in sorted(names):
        v = T[c, base + k]
        if v == 0: continue
        if prev is not None: out.append(f"{names[k]}+{v - prev}")
        prev = v
    print(nm, "CTA", c, "cycles:", " ".join(out))
pn = {300: "sp.enter", 301: "sp.rules", 302: "sp.values+warp reduce", 303: "sp.block reduce", 304: "sp.warp top-k", 305: "sp.merge", 310: "bu.enter", 311: "bu.candidates", 312: "bu.sort (thread 0)", 313: "bu.finished pool", 314: "bu.permute"}
prev = None
out = []
for k in sorted(pn):
    v = T[0, k] if k < LD else 0
    if v == 0: continue
    if prev is not None: out.append(f"{pn[k]}+{v - prev}")
    prev = v
if out: print("tail probes (CTA 0, cycles):", " ".join(out))
