"""ctypes wrapper around libwhisper_b200.so - the B200 counterpart of the reference's
``whisper/coreml.py`` (class ``Coreml``, whisper/coreml.py:19-244).

Same method names, argument meaning and return conventions:

  * results that stay inside the library (encoder output, cross K/V, the KV cache) come back as
    the ``dummy`` tensor (whisper/coreml.py:17,65,107);
  * ``decoder256Predict`` / ``decoder1Predict`` return long-lived preallocated fp32 host tensors
    that the next call overwrites (whisper/coreml.py:137-140,168,196-198,236);
  * loads are lazy and idempotent; per-stage wall-clock accumulators mirror
    whisper/coreml.py:9-13 and ``showB200PredictTime`` mirrors ``showCoremlPredictTime``.

Differences, on purpose: the library path is resolved next to this file instead of the
CWD-relative ``./coreml/<model>/coreml.so`` (whisper/coreml.py:21); model files are the ``.b2w``
containers written by ``export.py``; native errors raise instead of being logged and swallowed
(coreml.mm:54-56); an unloaded sub-model is loaded *and then run* on first use (the reference
returns ``None`` from that first call, whisper/coreml.py:51-53).
"""
from __future__ import annotations

import ctypes
import os
from timeit import default_timer as timer

import torch

from . import _lib

f32Ptr = _lib.f32p
logPredictTime = False

totalLoadTime = 0.0
totalEncoderTime = 0.0
totalDecoder1Time = 0.0
totalDecoder256Time = 0.0
totalCrossKVTime = 0.0

dummy = torch.ones((1))


def _ptr(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), f32Ptr)


class B200:
    def __init__(self, n_audio_layer: int, n_text_layer: int, n_mels: int, n_state: int, n_head: int, n_vocab: int,
                 modelFolder: str, device: int = 0):
        self.obj = _lib.load()
        self.n_audio_layer = n_audio_layer
        self.n_text_layer = n_text_layer
        self.n_mels = n_mels
        self.n_state = n_state
        self.n_head = n_head
        self.n_alignment_head = -1   # for decoder256
        self.bs = -1                 # for decoder1
        self.n_vocab = n_vocab
        self.modelFolder = modelFolder
        self.isEncoderLoaded = False
        self.isCrossKVLoaded = False
        self.isDecoder1Loaded = False
        self.isDecoder256Loaded = False
        self.obj.b200SetDevice(device)

    def _path(self, name: str) -> bytes:
        return os.path.join(self.modelFolder, name).encode()

    # ---- Encoder --------------------------------------------------------------------------------
    def loadEncoder(self):
        global totalLoadTime
        if self.isEncoderLoaded:
            return
        startT = timer()
        self.obj.loadEncoder(self.modelFolder.encode(), self.n_audio_layer, self.n_state, self.n_mels)
        _lib.check_errors("loadEncoder")
        self.isEncoderLoaded = True
        totalLoadTime += timer() - startT

    def encoderPredict(self, melSegment: torch.Tensor):
        global totalEncoderTime
        self.loadEncoder()
        startT = timer()
        melSegment = melSegment.to(torch.float32).contiguous()
        if melSegment.numel() != self.n_mels * 3000:
            raise ValueError(f"melSegment must be (1, {self.n_mels}, 3000), got {tuple(melSegment.shape)}")
        self.obj.encoderPredict(_ptr(melSegment))
        _lib.check_errors("encoderPredict")
        if logPredictTime:
            print(f"\tb200 encoder {timer()-startT:.3f}")
        totalEncoderTime += timer() - startT
        return dummy

    def closeEncoder(self):
        if not self.isEncoderLoaded:
            return
        self.obj.closeEncoder()
        self.isEncoderLoaded = False

    # ---- CrossKV --------------------------------------------------------------------------------
    def loadCrossKV(self):
        global totalLoadTime
        if self.isCrossKVLoaded:
            return
        startT = timer()
        self.obj.loadCrossKV(self._path("CrossKV.b2w"), self.n_text_layer, self.n_state)
        _lib.check_errors("loadCrossKV")
        self.isCrossKVLoaded = True
        totalLoadTime += timer() - startT

    def crossKVPredict(self):
        global totalCrossKVTime
        self.loadCrossKV()
        startT = timer()
        self.obj.crossKVPredict()
        _lib.check_errors("crossKVPredict")
        if logPredictTime:
            print(f"\tb200 crossKV {timer()-startT:.3f}")
        totalCrossKVTime += timer() - startT
        return dummy, dummy

    def closeCrossKV(self):
        if not self.isCrossKVLoaded:
            return
        self.obj.closeCrossKV()
        self.isCrossKVLoaded = False

    # ---- Decoder256 -----------------------------------------------------------------------------
    def loadDecoder256(self):
        global totalLoadTime
        if self.isDecoder256Loaded:
            return
        startT = timer()
        self.obj.loadDecoder256(self._path("Decoder.b2w"), self.n_text_layer, self.n_state, self.n_head,
                                self.n_alignment_head, self.bs)
        _lib.check_errors("loadDecoder256")
        max_n_ctx = 256
        self.out_x256 = torch.ones((1, max_n_ctx, self.n_state), dtype=torch.float32).contiguous()
        self.out_cross_head_weights256 = torch.ones((max(self.n_alignment_head, 0), max_n_ctx, 1500),
                                                    dtype=torch.float32).contiguous()
        self.isDecoder256Loaded = True
        totalLoadTime += timer() - startT

    def decoder256Predict(self, x: torch.Tensor, qk_mask: torch.Tensor, beam_idx: int):
        global totalDecoder256Time
        self.loadDecoder256()
        startT = timer()
        x = x.to(torch.float32).contiguous()
        qk_mask = qk_mask.to(torch.float32).contiguous()
        if x.numel() != 256 * self.n_state or qk_mask.numel() != 256 * 256:
            raise ValueError("decoder256Predict expects x (1,256,n_state) and qk_mask (256,256)")
        self.obj.decoder256Predict(_ptr(x), _ptr(qk_mask), _ptr(self.out_x256),
                                   _ptr(self.out_cross_head_weights256), beam_idx)
        _lib.check_errors("decoder256Predict")
        if logPredictTime:
            print(f"\tb200 decoder256 {timer()-startT:.3f}")
        totalDecoder256Time += timer() - startT
        return self.out_x256, self.out_cross_head_weights256, dummy

    def closeDecoder256(self):
        if not self.isDecoder256Loaded:
            return
        self.obj.closeDecoder256()
        self.isDecoder256Loaded = False

    # ---- Decoder1 -------------------------------------------------------------------------------
    def loadDecoder1(self):
        global totalLoadTime
        if self.isDecoder1Loaded:
            return
        startT = timer()
        self.obj.loadDecoder1(self._path("Decoder.b2w"), self.n_text_layer, self.n_state, self.n_head, self.n_vocab)
        _lib.check_errors("loadDecoder1")
        self.out_x1 = torch.ones((self.bs, 1, self.n_vocab), dtype=torch.float32).contiguous()
        self.new_masked_kv_caches1 = torch.ones((self.n_text_layer * 2, self.bs, 1, self.n_state),
                                                dtype=torch.float32).contiguous()
        self.isDecoder1Loaded = True
        totalLoadTime += timer() - startT

    def rearrange_mkv(self, indices, text_offset: int):
        indices = torch.as_tensor(indices).to(torch.int32).contiguous()
        self.obj.rearrange_mkv(ctypes.cast(indices.data_ptr(), _lib.i32p), int(text_offset))
        _lib.check_errors("rearrange_mkv")

    def decoder1Predict(self, x: torch.Tensor, qk_mask: torch.Tensor, text_offset: int):
        global totalDecoder1Time
        self.loadDecoder1()
        startT = timer()
        x = x.to(torch.float32).contiguous()
        qk_mask = qk_mask.to(torch.float32).contiguous()
        # the library reads bs rows of x (the beam slots given to loadDecoder256, coreml.mm:230) and a (1, 449) mask, (1, 450) for bs == 1
        if x.numel() != self.bs * self.n_state:
            raise ValueError(f"decoder1Predict expects x of shape ({self.bs}, 1, {self.n_state}) - the decoder was loaded with "
                             f"{self.bs} beam slots - got {tuple(x.shape)}")
        if qk_mask.numel() != (450 if self.bs == 1 else 449):
            raise ValueError(f"decoder1Predict expects a qk_mask of {450 if self.bs == 1 else 449} elements, got {qk_mask.numel()}")
        self.obj.decoder1Predict(_ptr(x), _ptr(qk_mask), int(text_offset), _ptr(self.out_x1))
        _lib.check_errors("decoder1Predict")
        if logPredictTime:
            print(f"\tb200 decoder1 {timer()-startT:.3f}")
        totalDecoder1Time += timer() - startT
        return self.out_x1, self.new_masked_kv_caches1

    def closeDecoder1(self):
        if not self.isDecoder1Loaded:
            return
        self.obj.closeDecoder1()
        self.isDecoder1Loaded = False

    def close(self):
        self.closeDecoder1(); self.closeDecoder256(); self.closeCrossKV(); self.closeEncoder()


def showB200PredictTime():
    """Mirror of showCoremlPredictTime (whisper/coreml.py:247-263)."""
    print("  --- B200 load model ---")
    print(f"  total load time    {totalLoadTime:.3f}s")
    print("  --- B200 predict ------")
    print(f"  Encoder            {totalEncoderTime:.3f}s")
    print(f"  CrossKVCaches      {totalCrossKVTime:.3f}s")
    print(f"  Decoder256         {totalDecoder256Time:.3f}s")
    print(f"  Decoder1           {totalDecoder1Time:.3f}s")
    print("  ---")
    print(f"  total predict time {totalEncoderTime+totalCrossKVTime+totalDecoder1Time+totalDecoder256Time:.3f}s")
    print("  -------------------------")
