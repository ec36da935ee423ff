"""whisper.coreml_b200 - B200-native (sm_100a) drop-in for the plugin surface of wangchou/whisper.coreml.

The compute lives in ``libwhisper_b200.so`` (hand-written CUDA behind the C ABI of
``include/whisper_b200.h``); this package is the host-side mirror of the reference's
``whisper/coreml.py`` wrapper plus the callers either side of it.  There is no CPU fallback:
loading fails loudly when the library is missing.
"""
__version__ = "0.1.0"
