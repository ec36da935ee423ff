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


@pytest.mark.parametrize("sr", [48000, 44100, 22050, 8000, 16000])
def test_resample_to_16k_matches_oracle(sr):
    """csrc/resample.cu against the numpy restatement (fp64) of the same polyphase design: fp32 taps and accumulation."""
    from whisper_b200.audio import resample_to_16k
    x = torch.from_numpy(np.random.default_rng(sr).standard_normal(sr * 3 + 17).astype(np.float32)) * 0.1
    want = oa.resample_poly(x.double().numpy(), sr)
    got = resample_to_16k(x.cuda(), sr).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-6


def test_load_audio_wav_stereo_44k(tmp_path):
    """load_audio(): 16-bit PCM WAV -> mono fp32 / 32768 (whisper/audio.py:62) -> 16 kHz, all on the device."""
    import wave
    from whisper_b200.audio import load_audio
    rng = np.random.default_rng(5)
    pcm = (rng.standard_normal((44100, 2)) * 3000).astype("<i2")
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    got = load_audio(path, quantize_s16=False).cpu().numpy()
    mono = pcm.astype(np.float64).mean(axis=1) / 32768.0
    want = oa.resample_poly(mono, 44100)
    assert got.shape == want.shape == (16000,)
    assert np.abs(got - want).max() < 2e-6
    # default: rounded to the int16 grid like the reference's `ffmpeg -f s16le` pipe (whisper/audio.py:45-62)
    q = load_audio(path).cpu().numpy()
    assert np.abs(q * 32768.0 - np.round(q * 32768.0)).max() < 1e-3 and np.abs(q - want).max() <= 0.5 / 32768.0 + 2e-6
