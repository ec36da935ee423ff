"""Audio front-end on the device: the counterpart of whisper/audio.py:45-157 (load_audio, log_mel_spectrogram, pad_or_trim).
load_audio here decodes FLAC (csrc/flac.cu) and PCM WAV itself and down-mixes / resamples on the GPU (the reference pipes any
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


def decode_flac(data: bytes):
    """FLAC stream -> (int32 samples [n, channels] on the host, sample_rate, bits_per_sample, md5 of the PCM as stored by the
    encoder).  The entropy decoding is serial and runs on the host (csrc/flac.cu); everything after it runs on the device."""
    import numpy as np
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    rate, ch, bits, total = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_long()
    md5 = (ctypes.c_ubyte * 16)()
    rc = lib.b200FlacInfo(buf.ctypes.data_as(ctypes.c_void_p), buf.size, ctypes.byref(rate), ctypes.byref(ch), ctypes.byref(bits),
                          ctypes.byref(total), md5)
    _lib.check_errors("b200FlacInfo")
    if rc != 0:
        raise ValueError("not a FLAC stream")
    cap = total.value if total.value > 0 else buf.size * 8                      # unknown length: a sample takes at least a bit
    out = np.empty((cap, ch.value), dtype=np.int32)
    n = lib.b200FlacDecode(buf.ctypes.data_as(ctypes.c_void_p), buf.size, out.ctypes.data_as(_lib.i32p), cap)
    _lib.check_errors("b200FlacDecode")
    if n < 0:
        raise ValueError("FLAC decode failed")
    return out[:n], rate.value, bits.value, bytes(md5)


def _wav_pcm(path: str):
    """PCM WAV (8 / 16 / 24 / 32-bit integer samples) -> (int32 samples [n, channels], sample_rate, bits_per_sample)."""
    import wave

    import numpy as np
    with wave.open(path, "rb") as w:
        width, channels, rate, n = w.getsampwidth(), w.getnchannels(), w.getframerate(), w.getnframes()
        if w.getcomptype() != "NONE" or width not in (1, 2, 3, 4):
            raise ValueError(f"{path}: only integer PCM WAV is decoded here (sample width {width}, {w.getcomptype()})")
        raw = np.frombuffer(w.readframes(n), dtype=np.uint8)
    if width == 1:
        pcm = raw.astype(np.int32) - 128                                         # 8-bit WAV is unsigned
    elif width == 3:
        b = raw.reshape(-1, 3).astype(np.int32)
        pcm = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        pcm = np.where(pcm & 0x800000, pcm - (1 << 24), pcm).astype(np.int32)
    else:
        pcm = raw.view("<i2" if width == 2 else "<i4").astype(np.int32)
    return pcm.reshape(-1, channels), rate, 8 * width


def load_audio(path: str, sr: int = SAMPLE_RATE, device: str = "cuda", quantize_s16: bool = True) -> torch.Tensor:
    """whisper/audio.py:25-62 for FLAC and integer-PCM WAV files: mono fp32 waveform at 16 kHz, on the device.

    The reference pipes the file through `ffmpeg -f s16le -ac 1 -acodec pcm_s16le -ar 16000` and divides the int16 samples by
    32768.  Here the container is decoded by this library (FLAC: csrc/flac.cu, bit-exact against the stream's MD5; WAV: the
    standard library), the channels are averaged (ffmpeg's stereo -> mono mix), the waveform is resampled on the device
    (csrc/resample.cu) and, like the reference's s16 pipe, rounded to the int16 grid (`quantize_s16`).  Parity with ffmpeg's own
    resampling filter is unpinned (no ffmpeg in the build image): see DESIGN.md."""
    if sr != SAMPLE_RATE:
        raise ValueError("the hot path runs at 16 kHz")
    with open(path, "rb") as f:
        magic = f.read(4)
    if magic == b"fLaC":
        with open(path, "rb") as f:
            pcm, rate, bits, _ = decode_flac(f.read())
    elif magic == b"RIFF":
        pcm, rate, bits = _wav_pcm(path)
    else:
        raise ValueError(f"{path}: only FLAC and PCM WAV containers are decoded here; pass a waveform for anything else")
    lib = _lib.load()
    channels = pcm.shape[1]
    if bits == 16:                                                              # the common case keeps its dedicated kernel
        dev = torch.from_numpy(pcm.astype("<i2")).to(device)
        mono = torch.empty(pcm.shape[0], dtype=torch.float32, device=dev.device)
        lib.b200Pcm16ToMonoDev(ctypes.c_void_p(dev.data_ptr()), mono.numel(), channels, ctypes.c_void_p(mono.data_ptr()))
        _lib.check_errors("b200Pcm16ToMonoDev")
    else:
        mono = torch.from_numpy(pcm).to(device).to(torch.float32).mean(dim=1) / float(1 << (bits - 1))
    out = mono if rate == SAMPLE_RATE else resample_to_16k(mono, rate)
    if quantize_s16:
        out = torch.clamp(torch.round(out * 32768.0), -32768.0, 32767.0) / 32768.0
    return out
