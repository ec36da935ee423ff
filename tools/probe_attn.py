"""GPU probe: time + parity of the tcgen05 flash-attention kernel on the encoder's shape (1500 tokens, 20 heads, W windows)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import whisper_b200._lib as L
lib = L.load()
for W in (2, 8):
    n_tok, heads = 1500, 20
    d = heads * 64
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(W, n_tok, 3 * d, device="cuda", generator=g)
    qkv[..., :d] *= 0.5
    qkv = qkv.bfloat16().contiguous()
    out = torch.empty(W, n_tok, d, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    ms = lib.b200TestAttentionTime(qkv.data_ptr(), out.data_ptr(), n_tok, heads, W, 50)
    q, k, v = [t.float().view(W, n_tok, heads, 64).transpose(1, 2) for t in qkv.split(d, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(W, n_tok, d)
    rel = float((out.float() - ref).norm() / ref.norm())
    fl = 4.0 * n_tok * n_tok * d * W
    print(f"W={W}: {ms*1e3:.1f} us = {fl/ms/1e9:.0f} TF/s, rel err {rel:.2e}", flush=True)
L.check_errors("probe")

# per-block timeline of CTA (0,0,0)
import numpy as np
W, n_tok, heads = 2, 1500, 20
d = heads * 64
qkv = (torch.randn(W, n_tok, 3 * d, device="cuda") * 0.7).bfloat16().contiguous()
out = torch.empty(W, n_tok, d, device="cuda", dtype=torch.bfloat16)
buf = np.zeros(24 * 16, dtype=np.uint64)
torch.cuda.synchronize()
n = lib.b200TestAttentionTimeline(qkv.data_ptr(), out.data_ptr(), n_tok, heads, W, buf.ctypes.data, 24)
m = buf.reshape(24, 16).astype(np.int64)
t0 = m[0, 0]
print("softmax warp: 0 enter, 1 S ready, 2 S loaded, 3 P buffer free, 4 P published; MMA thread: 5 loop top, 8 K ready, 9 S buffer free, 10 QK issued, "
      "6 QK committed, 7 P seen, 11 PV issued, 12 PV committed   (cycles since block 0 enter)")
order = [0, 1, 2, 3, 4, 5, 8, 9, 10, 6, 7, 11, 12]
print("blk " + " ".join(f"{k:>7d}" for k in order))
for j in range(n):
    print(f"{j:3d} " + " ".join(f"{int(m[j, k] - t0):7d}" if m[j, k] else "      -" for k in order))
r = slice(3, n - 2)
print("softmax: P-pub to P-pub", np.diff(m[:n, 4])[2:].mean(), "| wait S", (m[r,1]-m[r,0]).mean(), "ld S", (m[r,2]-m[r,1]).mean(), "max (+wait P buffer)", (m[r,3]-m[r,2]).mean(), "exp+st", (m[r,4]-m[r,3]).mean())
print("MMA thread: top->K ready", (m[r,8]-m[r,5]).mean(), "->S free", (m[r,9]-m[r,8]).mean(), "->QK issued", (m[r,10]-m[r,9]).mean(), "->committed", (m[r,6]-m[r,10]).mean(),
      "->P seen", (m[r,7]-m[r,6]).mean(), "->PV issued", (m[r,11]-m[r,7]).mean(), "->PV committed", (m[r,12]-m[r,11]).mean())
