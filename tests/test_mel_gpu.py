"""log-mel parity (BASELINE.json: max-abs error <= 1e-4 against the reference's fp32 result)."""
import numpy as np
import pytest
import torch

from oracle import audio as oa, synth
from tests._util import golden

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("kind", ["noise", "hdr"])
@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_matches_oracle_and_golden(kind, n_mels):
    from whisper_b200.audio import log_mel_spectrogram
    audio = synth.noise_audio(1, 160000) if kind == "noise" else synth.hdr_audio(2, 160000)
    mine = log_mel_spectrogram(audio, n_mels, padding=480000)
    ref = oa.log_mel_spectrogram(audio, n_mels, padding=480000)
    assert mine.shape == ref.shape
    assert float((mine - ref).abs().max()) <= TOL
    g = golden("mel")
    assert float((mine[:, ::7] - torch.from_numpy(g[f"{kind}_{n_mels}"])).abs().max()) <= TOL      # the reference itself
    exact = torch.from_numpy(oa.log_mel_spectrogram_f64(audio.numpy(), n_mels, padding=480000)).float()
    assert float((mine - exact).abs().max()) <= TOL


def test_mel_device_pointer_path_and_edges():
    from whisper_b200.audio import log_mel_spectrogram
    audio = synth.noise_audio(5, 48000 + 77)                 # not a multiple of the hop
    host = log_mel_spectrogram(audio, 80, padding=1000)
    dev = log_mel_spectrogram(audio.cuda(), 80, padding=1000)
    assert dev.is_cuda and torch.equal(dev.cpu(), host)
    ref = oa.log_mel_spectrogram(audio, 80, padding=1000)
    assert host.shape == ref.shape and float((host - ref).abs().max()) <= TOL
    silence = torch.zeros(16000)
    z = log_mel_spectrogram(silence, 80)
    assert torch.allclose(z, oa.log_mel_spectrogram(silence, 80))
