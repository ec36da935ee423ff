"""Oracle: log-mel spectrogram (fp32 torch restatement + an fp64 direct-DFT cross-check).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates whisper/audio.py:110-157 and the
librosa Slaney filterbank that produced whisper/assets/mel_filters.npz (audio.py:92-107).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE, N_FFT, HOP, N_FRAMES, N_SAMPLES = 16000, 400, 160, 3000, 480000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * (200.0 / 3)
    log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return np.where(m >= 15.0, log, lin)


def mel_filterbank(n_mels: int) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels, htk=False, norm='slaney') -> (n_mels, 201) fp32.
    (audio.py:97-101 documents that this call produced the asset.)"""
    fft_f = np.linspace(0.0, SAMPLE_RATE / 2, N_FFT // 2 + 1)
    pts = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SAMPLE_RATE / 2), n_mels + 2))
    diff = np.diff(pts)
    ramps = pts[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / diff[:-1, None]
    upper = ramps[2:] / diff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (pts[2:n_mels + 2] - pts[:n_mels]))[:, None]
    return w.astype(np.float32)


def log_mel_spectrogram(audio: torch.Tensor, n_mels: int, padding: int = 0,
                        filters: torch.Tensor | None = None) -> torch.Tensor:
    """audio.py:145-157 verbatim in behaviour (torch.stft on CPU == the reference's own arithmetic)."""
    audio = audio.float()
    if padding > 0:
        audio = F.pad(audio, (0, padding))
    win = torch.hann_window(N_FFT)
    spec = torch.stft(audio, N_FFT, HOP, window=win, return_complex=True)
    power = spec[..., :-1].abs() ** 2
    fb = torch.from_numpy(mel_filterbank(n_mels)) if filters is None else filters
    x = torch.clamp(fb @ power, min=1e-10).log10()
    x = torch.maximum(x, x.max() - 8.0)
    return (x + 4.0) / 4.0


def log_mel_spectrogram_f64(audio: np.ndarray, n_mels: int, padding: int = 0) -> np.ndarray:
    """Exact (fp64, direct DFT) evaluation of the same definition - used to measure how far both the
    reference's fp32 STFT and the CUDA kernel sit from the true value."""
    a = np.asarray(audio, dtype=np.float64)
    if padding:
        a = np.concatenate([a, np.zeros(padding)])
    a = np.pad(a, (N_FFT // 2, N_FFT // 2), mode="reflect")                 # center=True, reflect
    n_frames = 1 + (len(a) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames - 1)[:, None]  # last frame dropped
    n = np.arange(N_FFT)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / N_FFT)                          # periodic hann
    k = np.arange(N_FFT // 2 + 1)
    ang = -2 * np.pi * np.outer(n, k) / N_FFT
    fr = a[idx] * win
    power = (fr @ np.cos(ang)) ** 2 + (fr @ np.sin(ang)) ** 2                # (frames, 201)
    mel = mel_filterbank(n_mels).astype(np.float64) @ power.T
    x = np.log10(np.maximum(mel, 1e-10))
    x = np.maximum(x, x.max() - 8.0)
    return (x + 4.0) / 4.0


def pad_or_trim(x: torch.Tensor, length: int = N_FRAMES) -> torch.Tensor:
    """audio.py:65-88 on the last axis."""
    if x.shape[-1] > length:
        x = x[..., :length]
    if x.shape[-1] < length:
        x = F.pad(x, (0, length - x.shape[-1]))
    return x


def resample_poly(x: np.ndarray, sr_in: int, sr_out: int = 16000) -> np.ndarray:
    """The polyphase resampler of csrc/resample.cu restated in numpy (fp64): Kaiser(5.0) windowed sinc, cut-off 1 / max(up, down),
    half length 10 * max(up, down), unit DC gain times `up` - the design of scipy.signal.resample_poly, against which
    tests/test_oracle_golden.py checks this function.  (The reference itself resamples inside an ffmpeg subprocess,
    whisper/audio.py:45-62; ffmpeg is not in /root/reference and not in this image: parity against it is unpinned.)"""
    import math
    g = math.gcd(int(sr_in), int(sr_out))
    up, down = sr_out // g, sr_in // g
    x = np.asarray(x, dtype=np.float64)
    if up == down:
        return x.copy()
    max_rate = max(up, down)
    half = 10 * max_rate
    m = np.arange(-half, half + 1, dtype=np.float64)
    h = (1.0 / max_rate) * np.sinc(m / max_rate) * np.i0(5.0 * np.sqrt(np.clip(1 - (m / half) ** 2, 0, 1))) / np.i0(5.0)
    h = h / h.sum() * up
    n_out = -(-len(x) * up // down)
    # y[n] = sum_k x[k] h[n down - k up + half] over |n down - k up| <= half: gather the <= 2 half / up + 1 taps of every output
    n = np.arange(n_out, dtype=np.int64)[:, None]
    t = n * down
    k_first = -((half - t) // up)                        # ceil((t - half) / up)
    k = k_first + np.arange(2 * half // up + 1, dtype=np.int64)[None, :]
    j = t - k * up + half
    ok = (k >= 0) & (k < len(x)) & (j >= 0) & (j <= 2 * half)
    return (np.where(ok, x[np.clip(k, 0, len(x) - 1)], 0.0) * np.where(ok, h[np.clip(j, 0, 2 * half)], 0.0)).sum(axis=1)
