"""Weight export: reference state_dict -> .b2w containers read by libwhisper_b200.so.

The analogue of the reference's convert_encoder.py / convert_ckv.py / convert_decoder256.py /
convert_decoder.py + convert_coreml.sh: those bake the weights into four .mlmodelc bundles under
./coreml/<model>/; this writes <folder>/Encoder.b2w, CrossKV.b2w and Decoder.b2w (Decoder256 and
Decoder1 share one file).  Layout decisions made here, once, so the kernels stream weights in the
order they consume them:

  * GEMM operands (encoder, crossKV, prefill) are bf16 row-major [N, K] (nn.Linear layout) for TMA.
  * conv weights (d, C, 3) become [d, 3*Cpad] tap-major so the stem runs as three accumulating
    GEMM passes over shifted views of the zero-framed input (no im2col copy).
  * q/k/v are fused into one [3d, d] matrix; the encoder's k rows carry the 64^-0.5 scale
    (whisper/encoder.py:28,38; exact in bf16: a power of two); k has no bias (encoder.py:32).
  * decoder query weights/biases carry 0.125 exactly as after load_state_dict
    (whisper/decoder.py:16-20,42).  Pass fused=True if the dict comes from model.state_dict().
  * decoder1 GEMV operands are additionally stored "fragment-major": 16x32 tiles laid out in
    mma.m16n8k16 A-fragment order, so a warp's 128-bit loads are one contiguous 512-byte run.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional

import numpy as np
import torch

DT_F32, DT_BF16, DT_I32 = 0, 1, 2


def _pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def to_frag(w: torch.Tensor) -> torch.Tensor:
    """[N, K] -> [ceil(N/16), K/32, 2, 32, 8] (row half, lane = (row%8)*4 + (k%32)//8, 8 k's)."""
    n, k = w.shape
    assert k % 32 == 0, k
    n16 = _pad_to(n, 16)
    if n16 != n:
        w = torch.cat([w, w.new_zeros(n16 - n, k)], dim=0)
    v = w.reshape(n16 // 16, 2, 8, k // 32, 4, 8)          # nt, half, g, kc, t, e
    return v.permute(0, 3, 1, 2, 4, 5).contiguous().reshape(n16 // 16, k // 32, 2, 32, 8)


class _Writer:
    def __init__(self):
        self.entries = []
        self.blobs = []
        self.off = 0

    def add(self, name: str, t: torch.Tensor, dtype: int):
        assert len(name) < 64, name
        t = t.detach().cpu().contiguous()
        if dtype == DT_BF16:
            raw = t.to(torch.bfloat16).view(torch.int16).numpy().tobytes()
        elif dtype == DT_F32:
            raw = t.to(torch.float32).numpy().tobytes()
        else:
            raw = t.to(torch.int32).numpy().tobytes()
        shape = list(t.shape)[:4] + [0] * (4 - min(t.dim(), 4))
        if t.dim() > 4:                                     # frag tensors: store flattened
            shape = [t.numel(), 0, 0, 0]
        self.entries.append((name, dtype, min(t.dim(), 4), shape, self.off, len(raw)))
        pad = _pad_to(len(raw), 256) - len(raw)
        self.blobs.append(raw + b"\0" * pad)
        self.off += len(raw) + pad

    def write(self, path: str):
        n = len(self.entries)
        data_offset = _pad_to(24 + n * 120, 256)
        with open(path + ".tmp", "wb") as f:
            f.write(struct.pack("<4sIQQ", b"B2W1", n, data_offset, self.off))
            for name, dt, nd, shape, off, nb in self.entries:
                f.write(struct.pack("<64sII4QQQ", name.encode(), dt, nd, *shape, off, nb))
            f.write(b"\0" * (data_offset - f.tell()))
            for b in self.blobs:
                f.write(b)
        os.replace(path + ".tmp", path)


def _dims_tensor(dims) -> torch.Tensor:
    return torch.tensor([dims.n_mels, dims.n_audio_ctx, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer,
                         dims.n_vocab, dims.n_text_ctx, dims.n_text_state, dims.n_text_head, dims.n_text_layer],
                        dtype=torch.int32)


def export_model(state: Dict[str, torch.Tensor], dims, folder: str, fused: bool = False) -> str:
    """Write Encoder.b2w / CrossKV.b2w / Decoder.b2w under `folder` (like ./coreml/<model>/)."""
    os.makedirs(folder, exist_ok=True)
    sd = {k: v.detach().float().cpu() for k, v in state.items()}
    qs = 1.0 if fused else 0.125                           # whisper/decoder.py:16-20
    d, dt = dims.n_audio_state, dims.n_text_state
    cpad = _pad_to(dims.n_mels, 64)

    # ---------------------------------------------------------------- Encoder
    w = _Writer()
    w.add("dims", _dims_tensor(dims), DT_I32)
    c1 = sd["encoder.conv1.weight"]                        # (d, C, 3)
    c1p = torch.zeros(d, 3, cpad)
    c1p[:, :, :dims.n_mels] = c1.permute(0, 2, 1)
    w.add("conv1.w", c1p.reshape(d, 3 * cpad), DT_BF16)
    w.add("conv1.b", sd["encoder.conv1.bias"], DT_F32)
    w.add("conv2.w", sd["encoder.conv2.weight"].permute(0, 2, 1).reshape(d, 3 * d), DT_BF16)
    w.add("conv2.b", sd["encoder.conv2.bias"], DT_F32)
    w.add("pos", sd["encoder.positional_embedding"], DT_F32)
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}."
        w.add(f"l{i}.attn_ln.w", sd[p + "attn_ln.weight"], DT_F32)
        w.add(f"l{i}.attn_ln.b", sd[p + "attn_ln.bias"], DT_F32)
        qkv = torch.cat([sd[p + "attn.query.weight"], sd[p + "attn.key.weight"] * 0.125, sd[p + "attn.value.weight"]])
        qkv_b = torch.cat([sd[p + "attn.query.bias"], torch.zeros(d), sd[p + "attn.value.bias"]])
        w.add(f"l{i}.qkv.w", qkv, DT_BF16)
        w.add(f"l{i}.qkv.b", qkv_b, DT_F32)
        w.add(f"l{i}.out.w", sd[p + "attn.out.weight"], DT_BF16)
        w.add(f"l{i}.out.b", sd[p + "attn.out.bias"], DT_F32)
        w.add(f"l{i}.mlp_ln.w", sd[p + "mlp_ln.weight"], DT_F32)
        w.add(f"l{i}.mlp_ln.b", sd[p + "mlp_ln.bias"], DT_F32)
        w.add(f"l{i}.mlp1.w", sd[p + "mlp.0.weight"], DT_BF16)
        w.add(f"l{i}.mlp1.b", sd[p + "mlp.0.bias"], DT_F32)
        w.add(f"l{i}.mlp2.w", sd[p + "mlp.2.weight"], DT_BF16)
        w.add(f"l{i}.mlp2.b", sd[p + "mlp.2.bias"], DT_F32)
    w.add("ln_post.w", sd["encoder.ln_post.weight"], DT_F32)
    w.add("ln_post.b", sd["encoder.ln_post.bias"], DT_F32)
    w.write(os.path.join(folder, "Encoder.b2w"))

    # ---------------------------------------------------------------- CrossKV (decoder.py:172-187)
    w = _Writer()
    w.add("dims", _dims_tensor(dims), DT_I32)
    ws, bs = [], []
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}.cross_attn."
        ws += [sd[p + "key.weight"], sd[p + "value.weight"]]
        bs += [torch.zeros(dt), sd[p + "value.bias"]]
    w.add("ckv.w", torch.cat(ws), DT_BF16)                 # [2*Ld*d, d], rows ordered (layer, k|v, d)
    w.add("ckv.b", torch.cat(bs), DT_F32)
    w.write(os.path.join(folder, "CrossKV.b2w"))

    # ---------------------------------------------------------------- Decoder (prefill + step)
    w = _Writer()
    w.add("dims", _dims_tensor(dims), DT_I32)
    emb = sd["decoder.token_embedding.weight"]
    w.add("tok_emb.w", emb, DT_BF16)
    w.add("tok_emb.frag", to_frag(emb), DT_BF16)
    w.add("pos_emb", sd["decoder.positional_embedding"], DT_F32)

    def lin(name, weight, bias):
        w.add(name + ".w", weight, DT_BF16)
        w.add(name + ".frag", to_frag(weight), DT_BF16)
        w.add(name + ".b", bias, DT_F32)

    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}."
        for ln in ("attn_ln", "cross_attn_ln", "mlp_ln"):
            w.add(f"l{i}.{ln}.w", sd[p + ln + ".weight"], DT_F32)
            w.add(f"l{i}.{ln}.b", sd[p + ln + ".bias"], DT_F32)
        lin(f"l{i}.qkv",
            torch.cat([sd[p + "attn.query.weight"] * qs, sd[p + "attn.key.weight"], sd[p + "attn.value.weight"]]),
            torch.cat([sd[p + "attn.query.bias"] * qs, torch.zeros(dt), sd[p + "attn.value.bias"]]))
        lin(f"l{i}.attn_out", sd[p + "attn.out.weight"], sd[p + "attn.out.bias"])
        lin(f"l{i}.cross_q", sd[p + "cross_attn.query.weight"] * qs, sd[p + "cross_attn.query.bias"] * qs)
        lin(f"l{i}.cross_out", sd[p + "cross_attn.out.weight"], sd[p + "cross_attn.out.bias"])
        lin(f"l{i}.mlp1", sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
        lin(f"l{i}.mlp2", sd[p + "mlp.2.weight"], sd[p + "mlp.2.bias"])
    w.add("ln.w", sd["decoder.ln.weight"], DT_F32)
    w.add("ln.b", sd["decoder.ln.bias"], DT_F32)
    w.write(os.path.join(folder, "Decoder.b2w"))
    return folder
