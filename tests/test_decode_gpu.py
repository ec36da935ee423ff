"""Device-resident decode loop (b200DecodeWindow) and word-timestamp alignment against the oracle and the
reference goldens.  BASELINE.json: greedy/beam token sequences >= 99% identical; DTW paths bit-exact given
the same cost matrix."""
import numpy as np
import pytest
import torch

from oracle import audio as oa, decoding as od, model as om, synth, timing as ot
from tests._util import close_library, exported, golden, rel

pytestmark = pytest.mark.gpu

CASES = {  # name -> (dims name, seed, logit_scale, sample_len, golden tag)
    "nano": ("nano", 0, 1.0, 24, "nano"),
    "nano_soft": ("nano", 1, 0.03, 40, "nano_soft"),
    "tiny": ("tiny", 0, 1.0, 24, "tiny"),
}


def _mel(dims):
    audio = synth.noise_audio(1, 480000)
    return oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)[:, :3000].contiguous()


def _agreement(a, b):
    n = max(len(a), len(b), 1)
    return sum(x == y for x, y in zip(a, b)) / n


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    from whisper_b200.model import ModelDimensions, WhisperB200
    name, seed, scale, sample_len, tag = CASES[request.param]
    dims, ckpt, folder = exported(name, seed, scale)
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    orc = om.OracleModel(dims, ckpt)
    mel = _mel(dims)
    m.encode_windows(mel.cuda(), [0])
    yield m, orc, dims, mel, sample_len, golden(tag)
    m.close()


@pytest.mark.parametrize("beam", [None, 5])
def test_tokens_match_oracle_and_reference(case, beam):
    from whisper_b200.decoding import DecodingOptions, decode
    m, orc, dims, mel, sample_len, g = case
    sp = od.Specials.load(dims.n_vocab)
    want = od.decode_window(orc, mel, sp, od.Options(sample_len=sample_len, beam_size=beam))
    got = decode(m, DecodingOptions(sample_len=sample_len, beam_size=beam), window=0)
    tag = "beam5" if beam else "greedy"
    ref_tokens = g[f"{tag}_tokens"].tolist()
    assert _agreement(got.tokens, want.tokens) >= 0.99, (got.tokens, want.tokens)
    assert _agreement(got.tokens, ref_tokens) >= 0.99, (got.tokens, ref_tokens)
    assert got.steps == want.steps
    assert abs(got.avg_logprob - want.avg_logprob) <= 2e-2 * max(1.0, abs(want.avg_logprob))
    assert abs(got.no_speech_prob - want.no_speech_prob) <= 2e-2


def test_without_timestamps_and_length_penalty(case):
    from whisper_b200.decoding import DecodingOptions, decode
    m, orc, dims, mel, sample_len, g = case
    sp = od.Specials.load(dims.n_vocab)
    want = od.decode_window(orc, mel, sp, od.Options(sample_len=12, beam_size=5, without_timestamps=True, length_penalty=1.0))
    got = decode(m, DecodingOptions(sample_len=12, beam_size=5, without_timestamps=True, length_penalty=1.0), window=0)
    assert _agreement(got.tokens, want.tokens) >= 0.99, (got.tokens, want.tokens)


def test_alignment(case):
    from whisper_b200.timing import align_tokens
    m, orc, dims, mel, sample_len, g = case
    sp = od.Specials.load(dims.n_vocab)
    tokens = g["align_tokens"].tolist()                      # [*sot_sequence, no_timestamps, *text, eot]
    n_skip = len(sp.sot_sequence)
    text = tokens[n_skip + 1:-1]
    al = align_tokens(sp.sot_sequence, sp.no_timestamps, sp.eot, text, 3000)
    # matrix close to the oracle's and to the reference golden (sub-sampled every 10 frames)
    orc.reset(); orc.xa = orc.encode(mel)
    _, chw = orc.logits(torch.tensor([tokens]))
    orc.reset()
    mat = ot.alignment_matrix(chw, 3000, n_skip)
    assert al.matrix.shape == tuple(mat.shape)
    assert rel(al.matrix, mat) < 5e-2, rel(al.matrix, mat)
    assert rel(al.matrix[:, ::10], g["align_matrix"]) < 5e-2
    # DTW: bit-exact given the same cost matrix
    oi, oj = ot.dtw(-al.matrix)
    assert np.array_equal(al.text_indices, oi) and np.array_equal(al.time_indices, oj)
    assert len(al.text_token_probs) == len(text) and np.all(al.text_token_probs >= 0) and np.all(al.text_token_probs <= 1)
    assert len(al.jump_times) == len(text) + 1              # one start time per matrix row (no_timestamps row + text rows)


def test_batched_windows_match_sequential():
    """b200DecodeWindows advances every window of a batch in ONE step kernel (windows x beams = the MMA N dimension, the weights
    streamed once per step): the tokens must be identical to decoding the windows one by one, for 2, 3 and 8 windows per batch,
    greedy and beam search (8 windows x 5 beams = 40 rows = 5 n-tiles; 3 windows = a ragged second n-tile)."""
    from whisper_b200.decoding import DecodingOptions, decode, decode_windows
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("tiny", 1, 0.03)                     # soft logits: the windows decode to different tokens
    close_library()
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        audio = torch.cat([synth.noise_audio(1 + i, 480000) for i in range(8)])
        mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
        m.encode_windows(mel.cuda(), [3000 * i for i in range(8)])
        for beam in (5, None):
            opts = DecodingOptions(sample_len=32, beam_size=beam)
            one_by_one = [decode(m, opts, window=w) for w in range(8)]
            assert len({tuple(r.tokens) for r in one_by_one}) > 1       # the windows really differ
            for group in ([0, 1], [2, 3, 4], list(range(8)), [7, 0, 3]):
                together = decode_windows(m, opts, group)
                for w, b in zip(group, together):
                    a = one_by_one[w]
                    assert a.tokens == b.tokens, (beam, group, w, a.tokens, b.tokens)
                    assert a.steps == b.steps
                    assert abs(a.sum_logprob - b.sum_logprob) <= 1e-3 * max(1.0, abs(a.sum_logprob))
                    assert abs(a.no_speech_prob - b.no_speech_prob) <= 1e-5
    finally:
        m.close()


def test_batched_windows_finish_at_different_steps():
    """Windows of one batch that finish at different steps (soft weights emit EOT early in some windows): a finished window keeps
    riding along in the batched step kernel and must neither change nor disturb the others."""
    from whisper_b200.decoding import DecodingOptions, decode, decode_windows
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("nano", 1, 0.03)
    close_library()
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    try:
        audio = torch.cat([synth.noise_audio(30 + i, 480000) for i in range(6)])
        mel = oa.log_mel_spectrogram(audio, dims.n_mels, padding=480000)
        m.encode_windows(mel.cuda(), [3000 * i for i in range(6)])
        orc = om.OracleModel(dims, ckpt)
        sp = od.Specials.load(dims.n_vocab)
        for beam in (None, 5):
            opts = DecodingOptions(sample_len=60, beam_size=beam)
            together = decode_windows(m, opts, range(6))
            for w in range(6):
                one = decode(m, opts, window=w)
                assert one.tokens == together[w].tokens and one.steps == together[w].steps, (beam, w, one.tokens, together[w].tokens)
            want = od.decode_window(orc, mel[:, :3000].contiguous(), sp, od.Options(sample_len=60, beam_size=beam))
            assert _agreement(together[0].tokens, want.tokens) >= 0.99, (beam, together[0].tokens, want.tokens)
    finally:
        m.close()


def test_decode_runs_into_the_context_limit():
    """sample_len larger than the text context: the loop must stop when the sequence exceeds n_text_ctx = 448 tokens
    (decoding.py:732), i.e. after 448 - len(sot_sequence) + 1 steps, with the KV cache written up to its last row (447)."""
    from whisper_b200.decoding import DecodingOptions, decode
    from whisper_b200.model import ModelDimensions, WhisperB200
    dims, ckpt, folder = exported("nano", 1, 0.03)
    m = WhisperB200(ModelDimensions(**dims.as_dict()), folder).load()
    mel = oa.log_mel_spectrogram(synth.noise_audio(21, 480000), dims.n_mels, padding=480000)[:, :3000].contiguous()
    m.encode_windows(mel.cuda(), [0])
    sp = od.Specials.load(dims.n_vocab)
    orc = om.OracleModel(dims, ckpt)
    for beam in (None, 5):
        want = od.decode_window(orc, mel, sp, od.Options(sample_len=1000, beam_size=beam))
        got = decode(m, DecodingOptions(sample_len=1000, beam_size=beam), window=0)
        assert got.steps == want.steps, (beam, got.steps, want.steps)
        assert want.steps <= 448 - len(sp.sot_sequence) + 1
        n = min(len(got.tokens), len(want.tokens))
        assert n > 0 and len(got.tokens) == len(want.tokens)
        # a long random-weight decode may flip a near-tie somewhere: the prefix up to the first difference must be long
        same = next((i for i in range(n) if got.tokens[i] != want.tokens[i]), n)
        assert same >= min(n, 64), (beam, same, n)
    m.close()
