"""Oracle: fp32 CPU restatement of the four sub-models (encoder, crossKV, decoder256, decoder1).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Every function cites the reference lines it
restates (paths relative to /root/reference).  Weights are a flat ``dict[str, Tensor]`` keyed
exactly like the reference ``state_dict`` so the same dict can be fed to the reference through
``load_state_dict`` (tests/golden/make_golden.py) and to the B200 exporter.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
N_AUDIO_CTX = 1500      # whisper/encoder.py:21 (hard-coded 1500 split)
N_TEXT_CTX = 448        # whisper/decoder.py:243 (hard-coded 448 mask)
PREFILL_CTX = 256       # whisper/decoder.py:163  max_n_ctx_for_1st
HEAD_DIM = 64           # whisper/decoder.py:62-64


@dataclass(frozen=True)
class Dims:
    """whisper/model.py:18-29 ModelDimensions."""
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int

    def as_dict(self):
        return asdict(self)


def _dims(m, d, h, le, ld, v):
    return Dims(m, N_AUDIO_CTX, d, h, le, v, N_TEXT_CTX, d, h, ld)


# Upstream checkpoint dimensions (coreml/coremlTest.cpp:16-21).
DIMS: Dict[str, Dims] = {
    "nano": _dims(80, 128, 2, 2, 2, 51865),        # test-only micro model (not an upstream size)
    "tiny": _dims(80, 384, 6, 4, 4, 51865),
    "base": _dims(80, 512, 8, 6, 6, 51865),
    "small": _dims(80, 768, 12, 12, 12, 51865),
    "large-v3": _dims(128, 1280, 20, 32, 32, 51866),
    "turbo": _dims(128, 1280, 20, 32, 4, 51866),
}


# ----------------------------------------------------------------------------------------------
# random-init recipe: same *distributions* as the reference constructor's defaults
# (nn.Linear/nn.Conv1d: U(+-1/sqrt(fan_in)); nn.Embedding: N(0,1); LayerNorm: 1/0), fixed seed,
# decoder.positional_embedding (torch.empty in whisper/decoder.py:138) filled with N(0, 0.01).
# ----------------------------------------------------------------------------------------------
def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> Tensor:
    """whisper/encoder.py:10-16."""
    half = channels // 2
    inc = math.log(max_timescale) / (half - 1)
    inv = torch.exp(-inc * torch.arange(half))
    ang = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


def init_weights(dims: Dims, seed: int = 0, logit_scale: float = 1.0) -> Dict[str, Tensor]:
    """Checkpoint-form weights (decoder query NOT yet multiplied by 0.125)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, Tensor] = {}

    def uniform(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    def linear(prefix, n_out, n_in, bias=True):
        w[prefix + ".weight"] = uniform((n_out, n_in), n_in)
        if bias:
            w[prefix + ".bias"] = uniform((n_out,), n_in)

    def lnorm(prefix, n):
        # default init is (1, 0); perturb slightly so gamma/beta handling is actually tested
        w[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
        w[prefix + ".bias"] = 0.1 * torch.randn(n, generator=g)

    def block(prefix, d, cross):
        for name in (("attn", "cross_attn") if cross else ("attn",)):
            linear(f"{prefix}.{name}.query", d, d)
            linear(f"{prefix}.{name}.key", d, d, bias=False)
            linear(f"{prefix}.{name}.value", d, d)
            linear(f"{prefix}.{name}.out", d, d)
            lnorm(f"{prefix}.{name}_ln", d)
        linear(f"{prefix}.mlp.0", 4 * d, d)
        linear(f"{prefix}.mlp.2", d, 4 * d)
        lnorm(f"{prefix}.mlp_ln", d)

    da, dt = dims.n_audio_state, dims.n_text_state
    w["encoder.conv1.weight"] = uniform((da, dims.n_mels, 3), dims.n_mels * 3)
    w["encoder.conv1.bias"] = uniform((da,), dims.n_mels * 3)
    w["encoder.conv2.weight"] = uniform((da, da, 3), da * 3)
    w["encoder.conv2.bias"] = uniform((da,), da * 3)
    w["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, da)
    for i in range(dims.n_audio_layer):
        block(f"encoder.blocks.{i}", da, cross=False)
    lnorm("encoder.ln_post", da)

    w["decoder.token_embedding.weight"] = torch.randn(dims.n_vocab, dt, generator=g) * logit_scale
    w["decoder.positional_embedding"] = torch.randn(dims.n_text_ctx, dt, generator=g) * 0.01
    for i in range(dims.n_text_layer):
        block(f"decoder.blocks.{i}", dt, cross=True)
    lnorm("decoder.ln", dt)
    return w


def fuse_query_scale(ckpt: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """whisper/decoder.py:16-20,42: load_state_dict pre-hook multiplies every *decoder*
    ``query.weight`` / ``query.bias`` by 0.125.  Encoder weights are untouched."""
    out = {}
    for k, v in ckpt.items():
        out[k] = v * 0.125 if (k.startswith("decoder.") and "query" in k) else v
    return out


def default_alignment_heads(dims: Dims) -> List[Tuple[int, int]]:
    """whisper/model.py:55-58 / decoder.py:166-170: all heads of the last Ld/2 layers."""
    return [(l, h) for l in range(dims.n_text_layer // 2, dims.n_text_layer)
            for h in range(dims.n_text_head)]


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _lin(w, p, x):
    return F.linear(x, w[p + ".weight"], w.get(p + ".bias"))


def _ln(w, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), w[p + ".weight"], w[p + ".bias"], eps)


def _heads(x: Tensor) -> Tensor:          # (n, d) -> (H, n, 64)
    n, d = x.shape
    return x.view(n, d // HEAD_DIM, HEAD_DIM).transpose(0, 1)


def _merge(x: Tensor) -> Tensor:          # (H, n, 64) -> (n, d)
    h, n, _ = x.shape
    return x.transpose(0, 1).reshape(n, h * HEAD_DIM)


def encoder_forward(w: Dict[str, Tensor], dims: Dims, mel: Tensor) -> Tensor:
    """whisper/encoder.py:103-136 (block12 loop) for ONE window.  mel (n_mels, 3000) -> (1500, d).

    conv1(k3,p1)+GELU, conv2(k3,s2,p1)+GELU, + sinusoid pos-emb (:122-128); per block
    pre-LN(eps=1e-7) MHA with k scaled by 64^-0.5 and no mask (:36-59, :61-80); ln_post (:133).
    """
    x = F.gelu(F.conv1d(mel[None], w["encoder.conv1.weight"], w["encoder.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["encoder.conv2.weight"], w["encoder.conv2.bias"], stride=2, padding=1))
    x = x[0].t() + w["encoder.positional_embedding"]
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        y = _ln(w, p + ".attn_ln", x, 1e-7)
        q = _heads(_lin(w, p + ".attn.query", y))
        k = _heads(_lin(w, p + ".attn.key", y) * (HEAD_DIM ** -0.5))
        v = _heads(_lin(w, p + ".attn.value", y))
        a = torch.softmax(q @ k.transpose(1, 2), dim=-1) @ v
        x = x + _lin(w, p + ".attn.out", _merge(a))
        y = _ln(w, p + ".mlp_ln", x, 1e-7)
        x = x + _lin(w, p + ".mlp.2", F.gelu(_lin(w, p + ".mlp.0", y)))
    return _ln(w, "encoder.ln_post", x, 1e-7)


def cross_kv(w: Dict[str, Tensor], dims: Dims, xa: Tensor) -> Tuple[Tensor, Tensor]:
    """whisper/decoder.py:172-187.  xa (1500,d) -> CK (Ld,H,64,1500) pre-transposed, CV (Ld,H,1500,64)."""
    ck, cv = [], []
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}.cross_attn"
        ck.append(_heads(_lin(w, p + ".key", xa)).transpose(1, 2))
        cv.append(_heads(_lin(w, p + ".value", xa)))
    return torch.stack(ck), torch.stack(cv)


def _decoder_blocks(w, dims, x, mask, mkv, ck, cv, heads):
    """whisper/decoder.py:261-329 forwardBlocks (torch branch) on a batch.

    x (B, n, d); mask broadcastable to (B, H, n, n_keys); mkv None or (2Ld, B, 448, d).
    Returns ln(x), list of raw cross QK for `heads` (batch item 0), new KV (2Ld, B, n, d).
    """
    B, n, d = x.shape
    H = d // HEAD_DIM
    new_kv, chw = [], {}
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        y = _ln(w, p + ".attn_ln", x, 1e-5)
        q = _lin(w, p + ".attn.query", y)           # 0.125 already fused (decoder.py:16-20)
        k_new = _lin(w, p + ".attn.key", y)
        v_new = _lin(w, p + ".attn.value", y)
        k, v = k_new, v_new
        if mkv is not None:                           # decoder.py:58-60
            k = torch.cat([mkv[2 * i], k_new], dim=1)
            v = torch.cat([mkv[2 * i + 1], v_new], dim=1)
        qh = q.view(B, n, H, HEAD_DIM).permute(0, 2, 1, 3)
        kh = k.view(B, -1, H, HEAD_DIM).permute(0, 2, 3, 1)
        vh = v.view(B, -1, H, HEAD_DIM).permute(0, 2, 1, 3)
        a = torch.softmax(qh @ kh + mask, dim=-1) @ vh                  # decoder.py:66-69
        x = x + _lin(w, p + ".attn.out", a.permute(0, 2, 1, 3).reshape(B, n, d))
        y = _ln(w, p + ".cross_attn_ln", x, 1e-5)
        qh = _lin(w, p + ".cross_attn.query", y).view(B, n, H, HEAD_DIM).permute(0, 2, 1, 3)
        qk = qh @ ck[i][None]                                             # decoder.py:84-87
        a = torch.softmax(qk, dim=-1) @ cv[i][None]
        x = x + _lin(w, p + ".cross_attn.out", a.permute(0, 2, 1, 3).reshape(B, n, d))
        y = _ln(w, p + ".mlp_ln", x, 1e-5)
        x = x + _lin(w, p + ".mlp.2", F.gelu(_lin(w, p + ".mlp.0", y)))
        for (l, h) in heads:                          # decoder.py:306-308 (batch item 0, raw QK)
            if l == i:
                chw[(l, h)] = qk[0, h]
        new_kv += [k_new, v_new]                      # decoder.py:310-314 even=K odd=V
    return _ln(w, "decoder.ln", x, 1e-5), [chw[k] for k in heads], torch.stack(new_kv)


def prefill_mask(n_ctx: int) -> Tensor:
    """whisper/decoder.py:212-213: causal (256,256) with columns >= n_ctx masked."""
    m = torch.full((PREFILL_CTX, PREFILL_CTX), -math.inf).triu_(1)
    m[:, n_ctx:] = -math.inf
    return m


def step_mask(text_offset: int) -> Tensor:
    """whisper/decoder.py:242-245: (1,449): 0 for [0,t), -inf for [t,448), 0 for the new token."""
    m = torch.zeros(1, N_TEXT_CTX + 1)
    m[0, text_offset:N_TEXT_CTX] = -math.inf
    return m


def decoder256(w, dims, x256: Tensor, mask: Tensor, ck, cv, heads):
    """One ``decoder256Predict`` call (coreml/coreml.mm:279-327 contract; torch body
    whisper/decoder.py:281-329 with qk_mask.shape[0]==256).  x256 (256,d) already embedded+padded.
    Returns ln(x) (256,d), CHW (n_align,256,1500) raw QK, KV (2Ld,256,d)."""
    x, chw, kv = _decoder_blocks(w, dims, x256[None], mask, None, ck, cv, heads)
    return x[0], (torch.stack(chw) if chw else torch.zeros(0, PREFILL_CTX, N_AUDIO_CTX)), kv[:, 0]


def decoder1(w, dims, x: Tensor, mask: Tensor, mkv: Tensor, ck, cv):
    """One ``decoder1Predict`` call (coreml/coreml.mm:404-444; whisper/decoder.py:241-257 +
    :316-327).  x (bs,d) embedded token, mkv (2Ld,bs,448,d).  Returns logits (bs,V), new KV (2Ld,bs,d).
    The bs==1 'nn.Linear speed-up' padding (decoder.py:247-248, 281-285) changes no values and is
    not restated."""
    y, _, kv = _decoder_blocks(w, dims, x[:, None], mask, mkv, ck, cv, [])
    logits = y[:, 0] @ w["decoder.token_embedding.weight"].t()          # decoder.py:319-320
    return logits, kv[:, :, 0]


class OracleModel:
    """Stateful mirror of Whisper + PyTorchInference (whisper/model.py:31-68,
    whisper/decoding.py:145-204): keeps Xa, CK/CV, the 448-slot KV cache and text_offset."""

    def __init__(self, dims: Dims, ckpt: Dict[str, Tensor], alignment_heads=None):
        self.dims = dims
        self.w = fuse_query_scale({k: v.float() for k, v in ckpt.items()})
        self.heads = list(alignment_heads) if alignment_heads is not None else default_alignment_heads(dims)
        self.text_offset = 0
        self.mkv: Optional[Tensor] = None
        self.xa = self.ck = self.cv = None

    # -- sub-models -------------------------------------------------------------------------
    def encode(self, mel: Tensor) -> Tensor:
        self.xa = encoder_forward(self.w, self.dims, mel.float())
        return self.xa

    def embed(self, tokens: Tensor, offset: int) -> Tensor:
        """whisper/decoder.py:202."""
        n = tokens.shape[-1]
        return (self.w["decoder.token_embedding.weight"][tokens]
                + self.w["decoder.positional_embedding"][offset:offset + n])

    def logits(self, tokens: Tensor, new_audio: bool = True):
        """PyTorchInference.logits + TextDecoder.forward (whisper/decoding.py:151-184,
        whisper/decoder.py:189-259).  tokens (bs, n) int64.  Returns logits (bs, n, V) and CHW."""
        d = self.dims.n_text_state
        if self.text_offset == 0:                                       # prefill branch
            if new_audio:
                self.ck, self.cv = cross_kv(self.w, self.dims, self.xa)
            bs, n = tokens.shape
            x = self.embed(tokens, 0)
            x = torch.cat([x, torch.zeros(bs, PREFILL_CTX - n, d)], dim=1)
            mask = prefill_mask(n)
            outs, kvs, chw0 = [], [], None
            for b in range(bs):                                         # decoder.py:217-234
                xb, chw, kv = decoder256(self.w, self.dims, x[b], mask, self.ck, self.cv, self.heads)
                outs.append(xb[:n]); kvs.append(kv)
                if b == 0:
                    chw0 = chw[:, :n]
            kv = torch.stack(kvs, dim=1)                                # (2Ld, bs, 256, d)
            pad = torch.zeros(kv.shape[0], bs, N_TEXT_CTX - PREFILL_CTX, d)
            self.mkv = torch.cat([kv, pad], dim=2)                      # decoding.py:169-176
            logits = torch.stack(outs) @ self.w["decoder.token_embedding.weight"].t()
            self.text_offset += n
            return logits, chw0
        x = self.embed(tokens[:, -1:], self.text_offset)[:, 0]
        logits, kv = decoder1(self.w, self.dims, x, step_mask(self.text_offset), self.mkv, self.ck, self.cv)
        self.mkv[:, :, self.text_offset] = kv                           # decoding.py:177-180
        self.text_offset += 1
        return logits[:, None], None

    def rearrange_kv_cache(self, source_indices: List[int]):
        """whisper/decoding.py:189-200 / coreml/coreml.mm:251-277: permute rows [:text_offset]."""
        if list(source_indices) != list(range(len(source_indices))):
            t = self.text_offset
            self.mkv[:, :, :t] = self.mkv[:, source_indices, :t].clone()

    def reset(self):
        self.text_offset = 0
        self.mkv = None
