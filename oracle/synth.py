"""Seeded synthetic inputs shared by the golden generator, the parity tests and bench.py
(Gaussian noise sigma=0.1, plus a high-dynamic-range clip for the mel test)."""
import math

import torch


def noise_audio(seed: int, n_samples: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_samples, generator=g) * 0.1


def hdr_audio(seed: int, n_samples: int) -> torch.Tensor:
    """Three tones + 1e-4 noise for the first half, digital silence for the second half."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float64) / 16000.0
    x = 0.5 * torch.sin(2 * math.pi * 440.0 * t) + 0.25 * torch.sin(2 * math.pi * 1234.5 * t) \
        + 0.05 * torch.sin(2 * math.pi * 5555.0 * t)
    x = x.float() + 1e-4 * torch.randn(n_samples, generator=g)
    x[n_samples // 2:] = 0.0
    return x
