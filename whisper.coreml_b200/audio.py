"""Audio front-end on the device: the counterpart of whisper/audio.py:65-157 (log_mel_spectrogram,
pad_or_trim).  ffmpeg decoding (audio.py:45-62) is outside the hot path: callers pass a 16 kHz mono
float waveform."""
from __future__ import annotations

import ctypes

import torch

from . import _lib

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE          # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH              # 3000
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH   # 100
TOKENS_PER_SECOND = SAMPLE_RATE // (HOP_LENGTH * 2)  # 50


def log_mel_spectrogram(audio: torch.Tensor, n_mels: int = 80, padding: int = 0, device=None) -> torch.Tensor:
    """whisper/audio.py:110-157.  `audio`: 1-D float waveform (CPU or CUDA).  Returns (n_mels, n_frames)
    fp32 on the device the computation ran on: CUDA input (or `device` given) keeps the result in HBM
    for encoderPredictWindows; CPU input returns a CPU tensor through the host-pointer entry point."""
    lib = _lib.load()
    if audio.dim() != 1:
        raise ValueError("log_mel_spectrogram expects a mono waveform")
    if device is not None:
        audio = audio.to(device)
    audio = audio.to(torch.float32).contiguous()
    n = audio.numel()
    n_frames = (n + padding) // HOP_LENGTH
    out = torch.empty((n_mels, n_frames), dtype=torch.float32, device=audio.device)
    if audio.is_cuda:
        got = lib.logMelSpectrogramDev(ctypes.c_void_p(audio.data_ptr()), n, padding, n_mels, ctypes.c_void_p(out.data_ptr()))
    else:
        got = lib.logMelSpectrogram(ctypes.cast(audio.data_ptr(), _lib.f32p), n, padding, n_mels,
                                    ctypes.cast(out.data_ptr(), _lib.f32p))
    _lib.check_errors("logMelSpectrogram")
    if got != n_frames:
        raise RuntimeError(f"logMelSpectrogram wrote {got} frames, expected {n_frames}")
    return out


def pad_or_trim(array: torch.Tensor, length: int = N_FRAMES, *, axis: int = -1) -> torch.Tensor:
    """whisper/audio.py:65-88 for tensors."""
    if array.shape[axis] > length:
        array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
    if array.shape[axis] < length:
        pad_widths = [(0, 0)] * array.ndim
        pad_widths[axis] = (0, length - array.shape[axis])
        array = torch.nn.functional.pad(array, [p for sizes in pad_widths[::-1] for p in sizes])
    return array
