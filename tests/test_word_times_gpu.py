"""Word times against the reference: the start / end of every text token reported by the CUDA path must equal the jump times of
the reference's find_alignment (whisper/timing.py:185-231) - computed here by the oracle from the same tokens and, for the golden
cases, by the reference itself (tests/golden/ref_*.npz: align_path_i / align_path_j).  The DTW is bit-exact given the same cost
matrix; the matrices differ by bf16 rounding, which may move a jump by one frame (0.02 s) where two cells tie: at most 1 % of the
tokens may differ, and then by at most 0.02 s."""
import numpy as np
import pytest
import torch

from oracle import audio as oa, decoding as od, model as om, synth, timing as ot
from tests._util import exported, golden

pytestmark = pytest.mark.gpu


def _jump_times(ti, tj):
    jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)             # timing.py:216-217
    return tj[jumps] / 50.0


def _oracle_alignment(orc, sp, mel_window, text_tokens, num_frames):
    """find_alignment (timing.py:176-217) on the oracle: (jump_times, text_token_probs)."""
    n_skip = len(sp.sot_sequence)
    tokens = [*sp.sot_sequence, sp.no_timestamps, *text_tokens, sp.eot]
    orc.reset()
    orc.encode(mel_window)
    logits, chw = orc.logits(torch.tensor([tokens]))
    orc.reset()
    sampled = logits[0, n_skip:, :sp.eot].float().softmax(dim=-1)                   # timing.py:187-190
    probs = sampled[np.arange(len(text_tokens)), text_tokens].numpy()
    mat = ot.alignment_matrix(chw, num_frames, n_skip)
    ti, tj = ot.dtw(-mat.double().numpy())
    _oracle_alignment.last = (mat.double().numpy(), ti, tj)                         # for _compare_paths
    return _jump_times(ti, tj), probs


def _compare_paths(al, what):
    """Ill-conditioned cost matrices (the 2-layer, 128-wide test model: its z-scored attention rows are almost flat, so the DTW
    path is decided by differences inside bf16 rounding and two near-optimal paths may lie seconds apart).  What must hold there:
    the matrix equals the oracle's to bf16 accuracy, the path is a valid monotone path from corner to corner, and its cost ON THE
    ORACLE'S MATRIX is within 1 % of the optimum - i.e. the DTW found a path the reference would score as (almost) its own."""
    mat, ti, tj = _oracle_alignment.last
    got = np.asarray(al.matrix, dtype=np.float64)
    assert got.shape == mat.shape, (what, got.shape, mat.shape)
    assert np.abs(got - mat).max() <= 0.15 and np.abs(got - mat).mean() <= 0.01, (what, float(np.abs(got - mat).max()), float(np.abs(got - mat).mean()))
    pi, pj = np.asarray(al.text_indices, dtype=np.int64), np.asarray(al.time_indices, dtype=np.int64)
    assert (pi[0], pj[0]) == (0, 0) and (pi[-1], pj[-1]) == (mat.shape[0] - 1, mat.shape[1] - 1), what
    di, dj = np.diff(pi), np.diff(pj)
    assert ((di >= 0) & (dj >= 0) & (di <= 1) & (dj <= 1) & (di + dj >= 1)).all(), what
    best, mine = float((-mat[ti, tj]).sum()), float((-mat[pi, pj]).sum())
    assert mine - best <= 0.01 * abs(best) + 1e-6, (what, mine, best)


def _compare_times(got, want, what, flat=False):
    """exact, except <= 1 % of the tokens (at least one) by one frame.  `flat`: the 2-layer, 128-wide test model - its z-scored
    attention rows differ by ~1e-3 between neighbouring frames, inside the bf16 rounding of the layers below, so a run of tokens
    that share a frame may move by a frame or two as a block: every time within 0.04 s and the mean shift below half a frame"""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    diff = np.abs(got - want)
    if flat:
        assert diff.max() <= 0.04 + 1e-6 and diff.mean() <= 0.01, (what, float(diff.max()), float(diff.mean()), got, want)
    else:
        assert diff.max() <= 0.02 + 1e-6, (what, float(diff.max()), got, want)
        assert (diff > 1e-6).sum() <= max(1, int(0.01 * len(got))), (what, int((diff > 1e-6).sum()), len(got))


@pytest.mark.parametrize("tag,name,seed,scale", [("nano", "nano", 0, 1.0), ("nano_soft", "nano", 1, 0.03), ("tiny", "tiny", 0, 1.0)])
def test_jump_times_match_reference_golden(tag, name, seed, scale):
    from whisper_b200.model import ModelDimensions, WhisperB200
    from whisper_b200.timing import align_tokens
    dims, ckpt, folder = exported(name, seed, scale)
    g = golden(tag)
    sp = od.Specials.load(dims.n_vocab)
    mel = oa.log_mel_spectrogram(synth.noise_audio(1, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        m.encode_windows(mel.cuda(), [0])
        tokens = g["align_tokens"].tolist()
        n_skip = len(sp.sot_sequence)
        text = tokens[n_skip + 1:-1]
        al = align_tokens(sp.sot_sequence, sp.no_timestamps, sp.eot, text, 3000)
    finally:
        m.close()
    ref_jumps = _jump_times(g["align_path_i"].astype(np.int64), g["align_path_j"].astype(np.int64))
    orc = om.OracleModel(dims, ckpt)
    jumps, probs = _oracle_alignment(orc, sp, mel, text, 3000)
    assert np.array_equal(jumps, ref_jumps)                                         # the oracle reproduces the reference's path exactly
    if tag == "nano":                                                               # sharp random weights on the 2-layer model: flat rows
        _compare_paths(al, tag + " vs oracle")
    else:
        _compare_times(al.jump_times, ref_jumps, tag + " vs reference golden", flat=name == "nano")
        _compare_times(al.jump_times, jumps, tag + " vs oracle", flat=name == "nano")
    assert np.allclose(al.text_token_probs, probs, atol=2e-2), float(np.abs(al.text_token_probs - probs).max())


def test_transcribe_word_times_match_oracle():
    """transcribe(word_timestamps=True) on two windows (BASELINE configs[2] shape, small dims): every word's start / end / probability
    against find_alignment of the oracle on the same text tokens (timing.py:283-287: the text tokens of all segments of a window)."""
    from whisper_b200.model import ModelDimensions, WhisperB200
    from whisper_b200.transcribe import transcribe
    dims, ckpt, folder = exported("small")
    audio = torch.cat([synth.noise_audio(1, 480000), synth.noise_audio(2, 480000)])
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        res = transcribe(m, audio, beam_size=5, word_timestamps=True, sample_len=20)
    finally:
        m.close()
    orc = om.OracleModel(dims, ckpt)
    sp = od.Specials.load(dims.n_vocab)
    mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
    n_words = 0
    for seek in res["seeks"]:
        segs = [s for s in res["segments"] if s["seek"] == seek]
        text = [t for s in segs for t in s["tokens"] if t < sp.eot]
        if not text:
            continue
        jumps, probs = _oracle_alignment(orc, sp, mel[:, seek:seek + 3000].contiguous(), text, 3000)
        t0 = seek * 0.01
        words = [w for s in segs for w in s["words"]]
        assert [w["token"] for w in words] == text
        _compare_times([w["start"] for w in words], np.round(t0 + jumps[:-1], 2), f"starts of window {seek}")
        _compare_times([w["end"] for w in words], np.round(t0 + jumps[1:], 2), f"ends of window {seek}")
        assert np.allclose([w["probability"] for w in words], probs, atol=2e-2)
        n_words += len(words)
    assert n_words > 0
