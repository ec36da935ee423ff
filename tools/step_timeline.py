"""Per-stage timeline of the batched persistent decoder step kernel on turbo dims (GPU).

    python tools/step_timeline.py [model] [windows]

Every CTA's thread 0 stores %globaltimer when it enters a stage, when the stage's prologue is done and when the stage is done
(decoder_batch.cu: DbDbg); this prints, per stage, the first / last CTA for each of the three marks."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import weights_folder
from oracle import synth
from whisper_b200.model import ModelDimensions, WhisperB200
from whisper_b200.audio import log_mel_spectrogram
from whisper_b200.decoding import DecodingOptions, decode_windows
name = sys.argv[1] if len(sys.argv) > 1 else "turbo"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dims, folder, _ = weights_folder(name, 0)
m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
mel = log_mel_spectrogram(synth.noise_audio(1, 480000 * W).cuda(), dims.n_mels, padding=480000)
m.encode_windows(mel, [3000 * w for w in range(W)])
decode_windows(m, DecodingOptions(beam_size=5, sample_len=40), range(W))      # warm
m.lib.b200TestStepTimeline(1, None, 0)
decode_windows(m, DecodingOptions(beam_size=5, sample_len=5), range(W))
LD = 640
buf = np.zeros(256 * LD, dtype=np.uint64)
n = m.lib.b200TestStepTimeline(0, buf.ctypes.data_as(ctypes.c_void_p), 256)
T = buf[:n * LD].astype(np.int64).reshape(n, LD)
SP = T[200:].reshape(-1, 8)             # marks of the separate sampling kernel: [CTA][entry, partial done, ticket, updated]
SP = SP[SP[:, 0] > 0]
T = T[:200]
T = T[T[:, 0] > 0]                      # CTAs that ran
n = T.shape[0]
STAGES = ["qkv", "self_attn", "out_proj", "cross_q", "cross_attn", "cross_out", "mlp1", "mlp2"]
n_stages = dims.n_text_layer * 8 + 1
t0 = T[:, 0].min()
print(f"{name}: {W} windows x 5 beams, {n} CTAs; kernel start skew {(T[:, 0].max() - t0) / 1000:.2f} us")
print(f"{'stage':22s} {'first in':>9s} {'last in':>9s} | {'first ready':>11s} {'last ready':>10s} | {'first done':>10s} {'last done':>9s}   (us since start; in = entered, ready = prologue done)")
prev_done = T[:, 0]
per_kind = {}
for it in range(n_stages):
    nm = f"L{it // 8}.{STAGES[it % 8]}" if it < n_stages - 1 else "vocab"
    ready, done = T[:, 2 * it + 1], T[:, 2 * it + 2]
    f = lambda a: (a - t0) / 1000.0
    print(f"{nm:22s} {f(prev_done.min()):9.2f} {f(prev_done.max()):9.2f} | {f(ready.min()):11.2f} {f(ready.max()):10.2f} | {f(done.min()):10.2f} {f(done.max()):9.2f}")
    kind = STAGES[it % 8] if it < n_stages - 1 else "vocab"
    per_kind.setdefault(kind, []).append((done.max() - prev_done.max()) / 1000.0)
    prev_done = done
if len(SP):
    f = lambda a: (a - t0) / 1000.0
    print(f"sampling kernel ({len(SP)} CTAs): entry first {f(SP[:, 0].min()):.2f} last {f(SP[:, 0].max()):.2f} | partial done first {f(SP[:, 1].min()):.2f} last {f(SP[:, 1].max()):.2f}"
          f" | ticket last {f(SP[:, 2].max()):.2f} | updated {f(SP[:, 3].max()):.2f}")
print("step total", (T[:, :2 * n_stages + 1].max() - t0) / 1000.0, "us")
print("mean us per stage kind (last done -> last done):", {k: round(float(np.mean(v)), 2) for k, v in per_kind.items()})

probe = os.environ.get("B200_STEP_PROBE")
if probe is not None:
    it = int(probe)
    print(f"inner marks of stage {it}: SM cycles since the CTA's first inner mark, and since the previous mark (median over CTAs; thread 0 only)")
    cols = [k for k in range(39) if (T[:, 600 + k] > 0).sum() * 2 >= T.shape[0] or (T[:, 600 + k] > 0).sum() >= 20]
    prev = None
    for k in cols:
        ok = (T[:, 600 + k] > 0) & (T[:, 600 + cols[0]] > 0)
        since0 = np.median(T[ok, 600 + k] - T[ok, 600 + cols[0]])
        d = ""
        if prev is not None:
            ok2 = ok & (T[:, 600 + prev] > 0)
            d = f"{np.median(T[ok2, 600 + k] - T[ok2, 600 + prev]):8.0f}"
        print(f"  mark {k:2d}: {since0:9.0f} {d}   ({int(ok.sum())} CTAs)")
        prev = k
    print("  poll retries of the last copy batch (thread 0), max over CTAs:", int(T[:, 639].max()))
