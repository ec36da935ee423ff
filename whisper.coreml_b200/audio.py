"""Audio front-end on the device: the counterpart of whisper/audio.py:45-157 (load_audio, log_mel_spectrogram, pad_or_trim).
load_audio here reads PCM WAV files with the standard library and down-mixes / resamples on the GPU (the reference pipes any
container through an ffmpeg subprocess, audio.py:45-62; other containers are out of scope - pass a waveform instead)."""
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


def resample_to_16k(audio: torch.Tensor, sample_rate: int) -> torch.Tensor:
    """Rational resampling of a 1-D CUDA waveform to 16 kHz on the device (csrc/resample.cu; the polyphase design of
    scipy.signal.resample_poly)."""
    lib = _lib.load()
    audio = audio.to(torch.float32).contiguous()
    if not audio.is_cuda:
        raise ValueError("resample_to_16k expects a CUDA tensor")
    n_out = lib.b200ResampleDev(None, audio.numel(), int(sample_rate), None, 0)
    out = torch.empty(n_out, dtype=torch.float32, device=audio.device)
    got = lib.b200ResampleDev(ctypes.c_void_p(audio.data_ptr()), audio.numel(), int(sample_rate), ctypes.c_void_p(out.data_ptr()), n_out)
    _lib.check_errors("b200ResampleDev")
    if got != n_out:
        raise RuntimeError(f"b200ResampleDev wrote {got} samples, expected {n_out}")
    return out


def load_audio(path: str, sr: int = SAMPLE_RATE, device: str = "cuda") -> torch.Tensor:
    """whisper/audio.py:25-62 for 16-bit PCM WAV files: mono fp32 waveform at 16 kHz, on the device."""
    import wave

    import numpy as np
    if sr != SAMPLE_RATE:
        raise ValueError("the hot path runs at 16 kHz")
    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{path}: only 16-bit PCM WAV is decoded here (got sample width {w.getsampwidth()}, {w.getcomptype()})")
        channels, rate, n = w.getnchannels(), w.getframerate(), w.getnframes()
        pcm = torch.from_numpy(np.frombuffer(w.readframes(n), dtype="<i2").copy()).to(device)
    lib = _lib.load()
    mono = torch.empty(pcm.numel() // channels, dtype=torch.float32, device=pcm.device)
    lib.b200Pcm16ToMonoDev(ctypes.c_void_p(pcm.data_ptr()), mono.numel(), channels, ctypes.c_void_p(mono.data_ptr()))
    _lib.check_errors("b200Pcm16ToMonoDev")
    return mono if rate == SAMPLE_RATE else resample_to_16k(mono, rate)
