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
