"""ctypes loader for libwhisper_b200.so (the analogue of whisper/coreml.py:21 ``cdll.LoadLibrary``)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_long, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libwhisper_b200.so")      # B200_LIB: A/B runs against another build
_lib = None

f32p = POINTER(c_float)
i32p = POINTER(c_int)

# name -> (restype, argtypes); must list every symbol declared in include/whisper_b200.h
SIGNATURES = {
    # Part 1: reference ABI (coreml/coreml.h:5-31)
    "loadEncoder": (None, [c_char_p, c_int, c_int, c_int]),
    "closeEncoder": (None, []),
    "encoderPredict": (None, [f32p]),
    "loadCrossKV": (None, [c_char_p, c_int, c_int]),
    "closeCrossKV": (None, []),
    "crossKVPredict": (None, []),
    "loadDecoder256": (None, [c_char_p, c_int, c_int, c_int, c_int, c_int]),
    "closeDecoder256": (None, []),
    "decoder256Predict": (None, [f32p, f32p, f32p, f32p, c_int]),
    "loadDecoder1": (None, [c_char_p, c_int, c_int, c_int, c_int]),
    "closeDecoder1": (None, []),
    "rearrange_mkv": (None, [i32p, c_int]),
    "decoder1Predict": (None, [f32p, f32p, c_int, f32p]),
    # Part 2: additive entry points
    "b200LastError": (c_int, [c_char_p, c_int]),
    "b200KernelLaunchCount": (c_long, []),
    "b200SetDevice": (None, [c_int]),
    "b200SetAlignmentHeads": (None, [i32p, c_int]),
    "b200SelectWindow": (None, [c_int]),
    "b200SetDecodeSpec": (None, [c_int, c_int, c_int, c_int, c_int, i32p, c_int, i32p, c_int]),
    "logMelSpectrogram": (c_long, [f32p, c_long, c_long, c_int, f32p]),
    "logMelSpectrogramDev": (c_long, [c_void_p, c_long, c_long, c_int, c_void_p]),
    "b200Pcm16ToMonoDev": (c_long, [c_void_p, c_long, c_int, c_void_p]),
    "b200ResampleDev": (c_long, [c_void_p, c_long, c_int, c_void_p, c_long]),
    "b200FlacInfo": (c_int, [c_void_p, c_long, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_long), c_void_p]),
    "b200FlacDecode": (c_long, [c_void_p, c_long, i32p, c_long]),
    "encoderPredictWindows": (None, [c_void_p, c_long, i32p, c_int]),
    "encoderPredictWindowsContent": (None, [c_void_p, c_long, c_long, i32p, c_int]),
    "crossKVPredictWindows": (None, [c_int]),
    "b200DecodeWindow": (c_int, [i32p, c_int, c_int, c_int, c_int, c_int, i32p, i32p, f32p, f32p]),
    "b200DecodeWindows": (c_int, [i32p, c_int, i32p, c_int, c_int, c_int, c_int, c_int, i32p, i32p, f32p, f32p, i32p]),
    "b200DecodeWindowsEx": (c_int, [i32p, c_int, i32p, c_int, c_int, c_int, c_float, ctypes.c_ulonglong, c_int, c_int, c_int, i32p, i32p, f32p, f32p, i32p]),
    "decoder1StepFused": (None, [i32p, c_int, c_int, c_int, c_int, c_int, f32p, i32p]),
    "medianFilter": (None, [f32p, f32p, c_long, c_int, c_int]),
    "dtw": (c_int, [f32p, c_int, c_int, i32p, i32p]),
    "b200AlignTokens": (c_int, [i32p, c_int, c_int, c_int, c_int, i32p, i32p, f32p, f32p]),
    "b200GetStageTimes": (None, [f32p, c_int]),
    # test hooks
    "b200TestGemm": (None, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    "b200TestGemmTile": (None, [c_int]),
    "b200TestGemmTimeline": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "b200TestGemmTime": (ctypes.c_float, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "b200TestGetXa": (None, [f32p, c_int]),
    "b200TestGetCrossKV": (None, [f32p, f32p, c_int]),
    "b200TestGetKV": (None, [f32p, c_int]),
    "b200TestAttention": (None, [c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "b200TestAttentionTimeline": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int]),
    "b200TestAttentionTime": (ctypes.c_float, [c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "b200TestStepTimeline": (c_int, [c_int, c_void_p, c_int]),
}


def load() -> ctypes.CDLL:
    """Load the library or raise: the product has no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built - run `python whisper.coreml_b200/build.py`; "
                               "there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check_errors(where: str = "") -> None:
    """Raise if the library recorded an error since the last check (Part 1 calls are `void`)."""
    lib = load()
    buf = ctypes.create_string_buffer(1024)
    n = lib.b200LastError(buf, 1024)
    if n:
        raise RuntimeError(f"libwhisper_b200 reported {n} error(s){' in ' + where if where else ''}: "
                           f"{buf.value.decode(errors='replace')}")
