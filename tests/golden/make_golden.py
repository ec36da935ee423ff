"""Golden-vector generator.  RUN IN THE BUILD CONTAINER ONLY (needs /root/reference):

    python tests/golden/make_golden.py

It imports the *reference* (wangchou/whisper.coreml, PyTorch CPU path, use_coreml=False), loads
the oracle's seeded weights into it through `load_state_dict` (so the decoder's 0.125 query
fusion, whisper/decoder.py:16-20, fires exactly as in `load_model`), runs the reference on seeded
synthetic inputs and writes
  * whisper.coreml_b200/assets/tokenizer_specials.json   (token ids from whisper/tokenizer.py)
  * tests/golden/*.npz                                    (small sub-sampled outputs)
It also compares the oracle against the reference on the full tensors and prints the errors;
the committed fixtures let the same check run anywhere (tests/test_oracle_golden.py).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import whisper                                              # noqa: E402  (the reference)
from whisper.decoding import DecodingOptions, DecodingTask   # noqa: E402
from whisper.model import ModelDimensions, Whisper           # noqa: E402
from whisper.tokenizer import get_tokenizer                   # noqa: E402
from whisper import timing as ref_timing                      # noqa: E402
from whisper import audio as ref_audio                        # noqa: E402

from oracle import model as om, decoding as od, audio as oa, timing as ot, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)


def dump_specials():
    spec = {}
    for n_vocab, n_lang in ((51865, 99), (51866, 100)):
        tk = get_tokenizer(True, num_languages=n_lang, language="en", task="transcribe")

        class _M:                                            # minimal stand-in for DecodingTask
            pass
        sup = sorted(set(list(tk.non_speech_tokens) + [tk.transcribe, tk.translate, tk.sot,
                                                       tk.sot_prev, tk.sot_lm, tk.no_speech]))
        spec[str(n_vocab)] = dict(sot=tk.sot, eot=tk.eot, sot_sequence=list(tk.sot_sequence),
                                  no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
                                  no_speech=tk.no_speech, blank=tk.encode(" "), suppress=sup,
                                  sot_prev=tk.sot_prev, transcribe=tk.transcribe, translate=tk.translate)
    path = os.path.join(ROOT, "whisper.coreml_b200", "assets", "tokenizer_specials.json")
    with open(path, "w") as f:
        json.dump(spec, f)
    return spec


def ref_model(name, seed, logit_scale=1.0):
    dims = om.DIMS[name]
    ckpt = om.init_weights(dims, seed, logit_scale)
    m = Whisper(ModelDimensions(**dims.as_dict()), False, name)
    missing = m.load_state_dict(ckpt, strict=True)
    return dims, ckpt, m.eval()


def sub(x, *steps):
    sl = tuple(slice(None, None, s) for s in steps)
    return x[sl].contiguous().numpy()


def rel(a, b):
    return float((a - b).norm() / b.norm())


def model_goldens(name, seed=0, sample_len=24, logit_scale=1.0, tag=None):
    dims, ckpt, ref = ref_model(name, seed, logit_scale)
    tag_name = tag or name
    orc = om.OracleModel(dims, ckpt)
    audio = synth.noise_audio(1, 480000)
    mel = ref_audio.log_mel_spectrogram(audio, dims.n_mels, padding=480000)[:, :3000].contiguous()
    g = {}
    # --- encoder / crossKV ------------------------------------------------------------------
    xa_ref = ref.encoder(mel[None])[0]
    xa = orc.encode(mel)
    print(name, "encoder  oracle-vs-ref rel", rel(xa, xa_ref))
    g["xa"] = sub(xa_ref, 25, 16); g["xa_norm"] = xa_ref.norm().numpy()
    ck_ref, cv_ref = ref.decoder.crossKVCaches(xa_ref[None])
    ck, cv = om.cross_kv(orc.w, dims, xa)
    print(name, "crossKV  rel", rel(ck, ck_ref), rel(cv, cv_ref))
    g["ck"] = sub(ck_ref, 1, 1, 8, 50); g["cv"] = sub(cv_ref, 1, 1, 50, 8)
    # --- prefill + steps through the reference decoder --------------------------------------
    sp = od.Specials.load(dims.n_vocab)
    toks = torch.tensor([list(sp.sot_sequence)] * 5)
    ref.text_offset = 0; ref.masked_kv_caches = None
    lg_ref, chw_ref, mkv = ref.decoder(toks, xa_ref[None], 0, None)
    orc.reset(); lg, chw = orc.logits(toks)
    print(name, "prefill logits rel", rel(lg, lg_ref), "chw rel", rel(chw, chw_ref))
    g["prefill_logits_top"] = lg_ref[0, -1].topk(16).indices.numpy()
    g["prefill_logits"] = sub(lg_ref[0], 1, 97)
    g["prefill_chw"] = sub(chw_ref, 1, 1, 25)
    cache = torch.cat([mkv, torch.zeros(mkv.shape[0], 5, 448 - 256, dims.n_text_state)], dim=2)
    t = toks.shape[1]
    nxt = torch.tensor([[11], [220], [50257], [1000], [51000]])
    lg1_ref, _, kv1 = ref.decoder(nxt, xa_ref[None], t, cache)
    lg1, _ = orc.logits(torch.cat([toks, nxt], dim=1))
    print(name, "step logits rel", rel(lg1, lg1_ref))
    g["step_tokens"] = nxt.numpy(); g["step_logits"] = sub(lg1_ref[:, 0], 1, 97)
    g["step_kv"] = sub(kv1[:, :, 0], 1, 1, 7)
    # --- full decode: greedy + beam 5 ---------------------------------------------------------
    for tag, kw in (("greedy", {}), ("beam5", dict(beam_size=5))):
        ref.text_offset = 0
        opts = DecodingOptions(language="en", fp16=False, sample_len=sample_len, **kw)
        r = ref.decode(mel, opts)
        o = od.decode_window(orc, mel, sp, od.Options(sample_len=sample_len, beam_size=kw.get("beam_size")))
        same = list(r.tokens) == list(o.tokens)
        print(name, tag, "ref tokens", r.tokens[:12], "... oracle identical:", same,
              "avg_logprob", r.avg_logprob, o.avg_logprob, "no_speech", r.no_speech_prob, o.no_speech_prob)
        g[f"{tag}_tokens"] = np.array(r.tokens, dtype=np.int64)
        g[f"{tag}_avg_logprob"] = np.float64(r.avg_logprob)
        g[f"{tag}_no_speech_prob"] = np.float64(r.no_speech_prob)
    # --- word-timestamp numerics: reference model.forward -> find_alignment internals ---------
    text_tokens = [int(x) for x in g["beam5_tokens"] if x < sp.eot][:20] or [11, 220, 1000]
    tokens = torch.tensor([*sp.sot_sequence, sp.no_timestamps, *text_tokens, sp.eot])
    ref.text_offset = 0; ref.masked_kv_caches = None
    out, chw_ref = ref(tokens[None])
    w = chw_ref[:, :, :1500].softmax(dim=-1)
    std, mean = torch.std_mean(w, dim=-2, keepdim=True, unbiased=False)
    w = ref_timing.median_filter((w - mean) / std, 7)
    mat = w.mean(axis=0)[len(sp.sot_sequence):-1]
    ti, tj = ref_timing.dtw(-mat)
    orc.reset(); _, chw_o = orc.logits(tokens[None], new_audio=False)
    mat_o = ot.alignment_matrix(chw_o, 3000, len(sp.sot_sequence))
    oi, oj = ot.dtw(-mat_o.numpy())
    print(name, "alignment matrix rel", rel(mat_o, mat), "dtw path equal:",
          len(oi) == len(ti) and bool((oi == ti).all() and (oj == tj).all()))
    g["align_tokens"] = tokens.numpy(); g["align_matrix"] = sub(mat, 1, 10)
    g["align_path_i"] = ti; g["align_path_j"] = tj
    np.savez_compressed(os.path.join(OUT, f"ref_{tag_name}.npz"), seed=np.int64(seed),
                        logit_scale=np.float64(logit_scale), **g)


def mel_goldens():
    g = {}
    for tag, audio in (("noise", synth.noise_audio(1, 160000)), ("hdr", synth.hdr_audio(2, 160000))):
        for m in (80, 128):
            ref = ref_audio.log_mel_spectrogram(audio, m, padding=480000)
            mine = oa.log_mel_spectrogram(audio, m, padding=480000)
            exact = torch.from_numpy(oa.log_mel_spectrogram_f64(audio.numpy(), m, padding=480000)).float()
            print("mel", tag, m, "oracle-vs-ref max", float((mine - ref).abs().max()),
                  "ref-vs-fp64 max", float((ref - exact).abs().max()))
            g[f"{tag}_{m}"] = ref[:, ::7].contiguous().numpy()
            g[f"{tag}_{m}_max"] = ref.max().numpy()
    np.savez_compressed(os.path.join(OUT, "ref_mel.npz"), **g)


def transcribe_goldens():
    """whisper.transcribe() of the reference on 100 s of noise with the nano_soft weights (its beam-5 output has consecutive
    timestamp pairs, so the data-dependent seek of transcribe.py:380-388 and the zero-padded partial last window are exercised)."""
    from oracle import transcribe as otr
    dims, ckpt, ref = ref_model("nano", 1, 0.03)
    audio = torch.cat([synth.noise_audio(10 + i, 480000) for i in range(4)])[:1600000]
    res = whisper.transcribe(ref, audio, beam_size=5, language="en", condition_on_previous_text=False, temperature=0.0,
                             compression_ratio_threshold=None, logprob_threshold=None, no_speech_threshold=None, sample_len=40, verbose=None)
    segs = res["segments"]
    g = dict(seg_seek=np.array([s["seek"] for s in segs], dtype=np.int64), seg_start=np.array([s["start"] for s in segs]),
             seg_end=np.array([s["end"] for s in segs]), seg_len=np.array([len(s["tokens"]) for s in segs], dtype=np.int64),
             seg_tokens=np.array([t for s in segs for t in s["tokens"]], dtype=np.int64))
    orc = om.OracleModel(dims, ckpt)
    got = otr.transcribe(orc, audio, od.Specials.load(dims.n_vocab), od.Options(sample_len=40, beam_size=5))
    print("transcribe: reference seeks", sorted(set(g["seg_seek"].tolist())), "oracle seeks", got["seeks"],
          "tokens equal", [t for s in got["segments"] for t in s["tokens"]] == g["seg_tokens"].tolist())
    np.savez_compressed(os.path.join(OUT, "ref_transcribe.npz"), **g)


def timing_goldens():
    """The reference's own known-answer tests (tests/test_timing.py:22-52, :67-84) run here,
    plus reference outputs on seeded inputs."""
    rng = np.random.RandomState(42)
    g = {}
    for n, m in ((10, 20), (32, 16), (123, 1500), (234, 189)):
        x = rng.randn(n, m).astype(np.float32)
        ri, rj = ref_timing.dtw_cpu(x.astype(np.float64))
        oi, oj = ot.dtw(x)
        assert len(ri) == len(oi) and (ri == oi).all() and (rj == oj).all(), "oracle dtw != reference"
        g[f"dtw_{n}_{m}_i"] = ri; g[f"dtw_{n}_{m}_j"] = rj
    # ties: quantised costs exercise the strict-< fall-through rule (timing.py:95-100)
    x = rng.randint(0, 3, size=(40, 60)).astype(np.float32)
    ri, rj = ref_timing.dtw_cpu(x.astype(np.float64)); oi, oj = ot.dtw(x)
    assert (ri == oi).all() and (rj == oj).all(), "oracle dtw tie rule != reference"
    g["dtw_ties_i"] = ri; g["dtw_ties_j"] = rj
    tg = torch.Generator().manual_seed(7)
    for shape in ((10,), (1, 15), (4, 5, 345), (3, 6, 40, 128)):
        x = torch.randn(*shape, generator=tg)
        for wdt in (3, 5, 7, 13):
            r = ref_timing.median_filter(x, wdt)
            o = ot.median_filter(x, wdt)
            assert torch.equal(r, o), f"oracle median != reference {shape} {wdt}"
    x = torch.randn(2, 9, 200, generator=torch.Generator().manual_seed(8))
    g["median7"] = ref_timing.median_filter(x, 7).numpy()
    print("timing: oracle == reference on dtw (5 cases) and median (16 cases)")
    np.savez_compressed(os.path.join(OUT, "ref_timing.npz"), **g)


if __name__ == "__main__":
    dump_specials()
    timing_goldens()
    transcribe_goldens()
    mel_goldens()
    for name in sys.argv[1:] or ["nano", "tiny"]:
        model_goldens(name)
    # near-uniform logits: exercises beam divergence, EOT/finished handling and timestamp rules
    model_goldens("nano", seed=1, sample_len=40, logit_scale=0.03, tag="nano_soft")
