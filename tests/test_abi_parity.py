"""GPU parity of the reference plugin ABI (encoder / crossKV / decoder256 / decoder1 / rearrange_mkv)
against the CPU oracle on the same seeded weights and inputs.  Tolerances from BASELINE.json:
encoder output and logits relative error <= 2e-2."""
import ctypes
import math

import pytest
import torch

from oracle import audio as oa, model as om, synth
from tests._util import exported, rel

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _mel(dims, seed=1):
    audio = synth.noise_audio(seed, 480000)
    return oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)[:, :3000].contiguous()


def _f32p(t):
    from whisper_b200 import _lib
    return ctypes.cast(t.data_ptr(), _lib.f32p)


@pytest.fixture(scope="module", params=["nano", "tiny"])
def setup(request):
    from whisper_b200 import b200, _lib
    dims, ckpt, folder = exported(request.param)
    be = b200.B200(dims.n_audio_layer, dims.n_text_layer, dims.n_mels, dims.n_audio_state, dims.n_audio_head,
                   dims.n_vocab, folder)
    orc = om.OracleModel(dims, ckpt)
    be.bs = 5
    be.n_alignment_head = len(orc.heads)
    yield dims, be, orc, _lib.load()
    be.close()


def test_encoder_crosskv_prefill_step(setup):
    dims, be, orc, lib = setup
    d, Ld, H = dims.n_text_state, dims.n_text_layer, dims.n_text_head
    mel = _mel(dims)
    # ---- encoder -------------------------------------------------------------------------------
    be.encoderPredict(mel[None])
    xa = torch.empty(1500, d)
    lib.b200TestGetXa(_f32p(xa), 0)
    xa_ref = orc.encode(mel)
    assert rel(xa, xa_ref) < TOL, rel(xa, xa_ref)
    # ---- crossKV (checked against the oracle applied to the oracle's Xa) -----------------------
    be.crossKVPredict()
    ck = torch.empty(Ld, H, 64, 1500); cv = torch.empty(Ld, H, 1500, 64)
    lib.b200TestGetCrossKV(_f32p(ck), _f32p(cv), 0)
    ck_ref, cv_ref = om.cross_kv(orc.w, dims, xa_ref)
    assert rel(ck, ck_ref) < TOL and rel(cv, cv_ref) < TOL, (rel(ck, ck_ref), rel(cv, cv_ref))
    # ---- decoder256, one call per beam like whisper/decoder.py:217-234 --------------------------
    toks = torch.tensor([[50258, 50259, 50359]] * 5)
    n = toks.shape[1]
    orc.reset()
    lg_ref, chw_ref = orc.logits(toks)
    x = orc.embed(toks, 0)
    x = torch.cat([x, torch.zeros(5, 256 - n, d)], dim=1)
    mask = om.prefill_mask(n)
    for b in range(5):
        out_x, out_chw, _ = be.decoder256Predict(x[b:b + 1], mask, b)
        if b == 0:
            lg = out_x[0, :n] @ orc.w["decoder.token_embedding.weight"].t()     # decoder.py:238-240 (host side)
            assert rel(lg, lg_ref[0]) < TOL, rel(lg, lg_ref[0])
            assert rel(out_chw[:, :n], chw_ref) < TOL, rel(out_chw[:, :n], chw_ref)
    kv = torch.empty(2 * Ld, 5, 256, d)
    lib.b200TestGetKV(_f32p(kv), 256)
    assert rel(kv, orc.mkv[:, :, :256]) < TOL
    # ---- decoder1 steps with beam permutations --------------------------------------------------
    be.loadDecoder1()
    tokens = toks
    g = torch.Generator().manual_seed(3)
    for step in range(6):
        nxt = torch.randint(0, 50257, (5, 1), generator=g)
        tokens = torch.cat([tokens, nxt], dim=1)
        t = orc.text_offset
        lg_ref, _ = orc.logits(tokens)
        xs = orc.embed(tokens[:, -1:], t)
        logits, _ = be.decoder1Predict(xs, om.step_mask(t), t)
        assert logits.shape == (5, 1, dims.n_vocab)
        assert rel(logits[:, 0], lg_ref[:, 0]) < TOL, (step, rel(logits[:, 0], lg_ref[:, 0]))
        # top-1 must agree wherever the oracle's margin is not within rounding
        top2 = lg_ref[:, 0].topk(2).values
        clear = (top2[:, 0] - top2[:, 1]) > 0.05 * lg_ref[:, 0].abs().max()
        assert (logits[:, 0].argmax(-1) == lg_ref[:, 0].argmax(-1))[clear].all()
        src = torch.randint(0, 5, (5,), generator=g).tolist()
        orc.rearrange_kv_cache(src)
        tokens = tokens[src]
        be.rearrange_mkv(src, orc.text_offset)
    t = orc.text_offset
    kv = torch.empty(2 * Ld, 5, t, d)
    lib.b200TestGetKV(_f32p(kv), t)
    assert rel(kv, orc.mkv[:, :, :t]) < TOL, rel(kv, orc.mkv[:, :, :t])


def test_errors_are_reported(setup):
    dims, be, orc, lib = setup
    from whisper_b200 import _lib
    be.loadDecoder256()
    x = torch.zeros(1, 256, dims.n_text_state)
    lib.decoder256Predict(_f32p(x), _f32p(torch.zeros(256, 256)), _f32p(x.clone()), None, 99)
    with pytest.raises(RuntimeError):
        _lib.check_errors("bad beam index")


def test_abi_smoke_driver_runs():
    """csrc/abi_smoke.cpp (the analogue of coreml/coremlTest.cpp): a plain C++ program linked against libwhisper_b200.so drives the
    13 reference entry points twice (load / predict / close) on exported tiny weights and exits 0 with finite outputs."""
    import os
    import subprocess
    from tests._util import close_library
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(here, "whisper.coreml_b200", "build", "abi_smoke")
    assert os.path.exists(exe), "build/abi_smoke missing: run __graft_entry__.build()"
    dims, ckpt, folder = exported("tiny")
    close_library()
    r = subprocess.run([exe, folder, str(dims.n_audio_layer), str(dims.n_text_layer), str(dims.n_audio_state), str(dims.n_mels),
                        str(dims.n_vocab), "5", str(len(om.default_alignment_heads(dims)))],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "abi_smoke: ok" in r.stdout, (r.returncode, r.stdout[-800:], r.stderr[-800:])
