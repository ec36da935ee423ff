"""Model container: counterpart of whisper/model.py (Whisper, ModelDimensions) wired to the B200
plugin the way the reference wires `use_coreml` (whisper/__init__.py:161-177, encoder.py:109-111,
decoder.py:172-175, 205-259, 272-279)."""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .b200 import B200, dummy


@dataclass
class ModelDimensions:                      # whisper/model.py:18-29
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


@dataclass(frozen=True)
class Specials:
    """Token ids the decode loop needs (whisper/tokenizer.py:147-275, 330-363).  Build from the reference
    tokenizer with `Specials.from_tokenizer`, or from the shipped table for the two stock vocabularies."""
    n_vocab: int
    sot: int
    eot: int
    sot_sequence: Tuple[int, ...]
    no_timestamps: int
    timestamp_begin: int
    no_speech: int
    blank: Tuple[int, ...]
    suppress: Tuple[int, ...]

    @staticmethod
    def load(n_vocab: int) -> "Specials":
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "tokenizer_specials.json")
        with open(path) as f:
            d = json.load(f)[str(n_vocab)]
        return Specials(n_vocab, d["sot"], d["eot"], tuple(d["sot_sequence"]), d["no_timestamps"], d["timestamp_begin"],
                        d["no_speech"], tuple(d["blank"]), tuple(d["suppress"]))

    @staticmethod
    def from_tokenizer(tk, n_vocab: int) -> "Specials":
        """Same suppress list as DecodingTask._get_suppress_tokens for suppress_tokens="-1" (decoding.py:599-626)."""
        sup = sorted(set(list(tk.non_speech_tokens) + [tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm, tk.no_speech]))
        return Specials(n_vocab, tk.sot, tk.eot, tuple(tk.sot_sequence), tk.no_timestamps, tk.timestamp_begin, tk.no_speech,
                        tuple(tk.encode(" ")), tuple(sup))


class WhisperB200:
    """One model on one GPU.  Holds the plugin wrapper and the two embeddings the reference keeps on the
    Python side when the plugin is active (whisper/__init__.py:181-189 drops everything else)."""

    def __init__(self, dims: ModelDimensions, folder: str, token_embedding: Optional[torch.Tensor] = None,
                 positional_embedding: Optional[torch.Tensor] = None, device: int = 0, beam_slots: int = 5,
                 alignment_heads: Optional[Sequence[Tuple[int, int]]] = None, specials: Optional[Specials] = None):
        self.dims = dims
        self.folder = folder
        self.device_index = device
        self.backend = B200(dims.n_audio_layer, dims.n_text_layer, dims.n_mels, dims.n_audio_state, dims.n_audio_head,
                            dims.n_vocab, folder, device)
        self.lib = self.backend.obj
        self.token_embedding = token_embedding          # (V, d) fp32 host, only for the reference-style forward()
        self.positional_embedding = positional_embedding
        self.specials = specials or Specials.load(dims.n_vocab)
        self.text_offset = 0                            # whisper/model.py:65-68
        self.n_windows = 0
        if alignment_heads is None:                     # whisper/model.py:55-58: last half of the decoder layers
            alignment_heads = [(l, h) for l in range(dims.n_text_layer // 2, dims.n_text_layer) for h in range(dims.n_text_head)]
        self.alignment_heads = list(alignment_heads)
        self.backend.bs = beam_slots
        self.backend.n_alignment_head = len(self.alignment_heads)
        self._loaded = False

    # ---- loading ----------------------------------------------------------------------------------
    def load(self):
        if self._loaded:
            return self
        be = self.backend
        pairs = np.array(self.alignment_heads, dtype=np.int32).reshape(-1)
        self.lib.b200SetAlignmentHeads(pairs.ctypes.data_as(_lib.i32p), len(self.alignment_heads))
        be.loadEncoder(); be.loadCrossKV(); be.loadDecoder256(); be.loadDecoder1()
        sp = self.specials
        sup = np.array(sp.suppress, dtype=np.int32); blank = np.array(sp.blank, dtype=np.int32)
        self.lib.b200SetDecodeSpec(sp.sot, sp.eot, sp.no_timestamps, sp.timestamp_begin, sp.no_speech,
                                   sup.ctypes.data_as(_lib.i32p), len(sup), blank.ctypes.data_as(_lib.i32p), len(blank))
        _lib.check_errors("load")
        self._loaded = True
        return self

    def set_alignment_heads(self, pairs: Iterable[Tuple[int, int]]):
        """whisper/model.py:70-78 takes a base85 dump; here the decoded (layer, head) pairs."""
        self.alignment_heads = list(pairs)
        arr = np.array(self.alignment_heads, dtype=np.int32).reshape(-1)
        self.lib.b200SetAlignmentHeads(arr.ctypes.data_as(_lib.i32p), len(self.alignment_heads))

    def close(self):
        self.backend.close()
        self._loaded = False

    # ---- reference-style forward through the plugin ABI (host buffers) -----------------------------------
    def encoder(self, mel: torch.Tensor):
        """AudioEncoder.forward with use_coreml (whisper/encoder.py:109-111): mel (1, n_mels, 3000)."""
        self.load()
        return self.backend.encoderPredict(mel)

    def decoder(self, tokens: torch.Tensor, xa, text_offset: int):
        """TextDecoder.forward with use_coreml (whisper/decoder.py:189-259): returns (logits, cross_qks, dummy)."""
        self.load()
        if self.token_embedding is None:
            raise RuntimeError("reference-style decoder() needs the host token/positional embeddings")
        be = self.backend
        n_batch, n_ctx = tokens.shape
        x = self.token_embedding[tokens] + self.positional_embedding[text_offset:text_offset + n_ctx]
        if text_offset == 0:
            if xa is not None:
                be.crossKVPredict()
            max_n_ctx = 256
            qk_mask = (torch.ones(max_n_ctx, max_n_ctx) * -np.inf).triu_(1)
            qk_mask[:, n_ctx:] = -np.inf
            x = torch.cat([x, torch.zeros(n_batch, max_n_ctx - n_ctx, self.dims.n_text_state)], dim=1)
            outs, cross_qks = [], None
            for b in range(n_batch):                                    # one call per beam (decoder.py:217-234)
                _x, _chw, _ = be.decoder256Predict(x[b:b + 1], qk_mask, b)
                outs.append(_x[:, :n_ctx].clone())
                if b == 0:
                    cross_qks = _chw[:, :n_ctx].clone()
            logits = (torch.cat(outs, dim=0) @ self.token_embedding.t()).float()
        else:
            qk_mask = torch.cat([torch.zeros((1, text_offset)), torch.ones((1, 448 - text_offset)) * -np.inf,
                                 torch.zeros((1, 1))], dim=1)
            if n_batch == 1:
                qk_mask = torch.cat([qk_mask, torch.full((1, 1), -np.inf)], dim=1)
            logits, _ = be.decoder1Predict(x, qk_mask, text_offset)
            cross_qks = None
        return logits, cross_qks, dummy

    # ---- device-resident fast path --------------------------------------------------------------------------
    current_window = 0

    def encode_windows(self, mel: torch.Tensor, seeks: Sequence[int], content_frames: Optional[int] = None) -> int:
        """Batched encoder + crossKV over independent windows of a device-resident log-mel (n_mels, frames).  Frames at or past
        `content_frames` are fed as zeros (the reference's pad_or_trim of a partial last window, transcribe.py:286-290)."""
        self.load()
        if not mel.is_cuda:
            mel = mel.to(f"cuda:{self.device_index}")
        mel = mel.to(torch.float32).contiguous()
        arr = np.array(list(seeks), dtype=np.int32)
        self.lib.encoderPredictWindowsContent(ctypes.c_void_p(mel.data_ptr()), mel.shape[1],
                                              mel.shape[1] if content_frames is None else int(content_frames),
                                              arr.ctypes.data_as(_lib.i32p), len(arr))
        self.lib.crossKVPredictWindows(len(arr))
        _lib.check_errors("encode_windows")
        self.n_windows = len(arr)
        self.current_window = 0
        return self.n_windows

    def select_window(self, w: int):
        self.lib.b200SelectWindow(w)
        _lib.check_errors("select_window")
        self.current_window = int(w)

    def stage_times_ms(self, reset: bool = False) -> Dict[str, float]:
        buf = (ctypes.c_float * 7)()
        self.lib.b200GetStageTimes(buf, 1 if reset else 0)
        names = ["mel", "encoder", "crossKV", "decoder256", "decoder1", "sampling", "align"]
        return {n: float(v) for n, v in zip(names, buf)}
