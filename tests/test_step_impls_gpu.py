"""The decoder1 implementations - the batched persistent kernel (default; 8 or 4 consumer warps) and the one-window persistent
kernel it replaced (B200_STEP_IMPL=mega) - must agree with each other; run in subprocesses because the choice is made once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, json, torch
sys.path.insert(0, %r)
from oracle import audio as oa, decoding as od, model as om, synth
from tests._util import exported
from whisper_b200.model import ModelDimensions, WhisperB200
from whisper_b200.decoding import DecodingOptions, decode
out = {}
for name, seed, scale, sl in (("nano", 1, 0.03, 40), ("tiny", 0, 1.0, 24)):
    dims, ckpt, folder = exported(name, seed, scale)
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    mel = oa.log_mel_spectrogram(synth.noise_audio(1, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
    m.encode_windows(mel.cuda(), [0])
    for beam in (None, 5):
        r = decode(m, DecodingOptions(sample_len=sl, beam_size=beam), window=0)
        out[f"{name}_{beam}"] = [r.tokens, r.sum_logprob, r.steps]
    m.close()
print("RESULT" + json.dumps(out))
''' % ROOT


def _run(impl, warps="8"):
    env = dict(os.environ, B200_STEP_WARPS=warps)
    if impl:
        env["B200_STEP_IMPL"] = impl
    p = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1]
    return json.loads(line[6:])


@pytest.mark.parametrize("warps", ["8", "4"])
def test_batched_kernel_matches_one_window_kernel(warps):
    a, b = _run("", warps), _run("mega")
    assert a.keys() == b.keys()
    for k in a:
        same = sum(x == y for x, y in zip(a[k][0], b[k][0])) / max(len(a[k][0]), len(b[k][0]), 1)
        assert same >= 0.99, (k, a[k], b[k])
        assert a[k][2] == b[k][2]
        assert abs(a[k][1] - b[k][1]) <= 2e-2 * max(1.0, abs(b[k][1])), (k, a[k][1], b[k][1])
