"""CPU: repository contract - the product never touches the oracle, the library exports every symbol the
header declares, the host logic that needs no GPU works."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "whisper.coreml_b200")


def test_product_does_not_import_oracle_or_reference():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "/root/reference" in src or "oracle/" in src:
                    bad.append(f)
    assert not bad, bad


def test_no_triton_or_torch_compile_in_product():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import triton" not in src and "torch.compile" not in src, f


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(?:void|int|long|float)\s+([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr))
    assert {"loadEncoder", "encoderPredict", "crossKVPredict", "decoder256Predict", "decoder1Predict", "rearrange_mkv"} <= declared
    from whisper_b200 import _lib
    lib = _lib.load()                          # loading needs no GPU
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_reference_abi_prototypes_match_coreml_h():
    """Part 1 must be prototype-for-prototype coreml/coreml.h:5-31 (names and argument lists)."""
    hdr = open(os.path.join(ROOT, "include", "whisper_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    flat = re.sub(r"\s+", " ", hdr)
    for proto in ["void loadEncoder(const char* modelFolderPath, int n_layer, int n_state, int n_mels);",
                  "void closeEncoder();", "void encoderPredict(float* melSegment);",
                  "void loadCrossKV(const char* modelPath, int n_layer, int n_state);", "void closeCrossKV();", "void crossKVPredict();",
                  "void loadDecoder256(const char* modelPath, int n_layer, int n_state, int n_head, int n_alignment_head, int beam_size);",
                  "void closeDecoder256();",
                  "void decoder256Predict( float* x, float* qk_mask, float* out_x, float* out_cross_head_weights, int beam_idx);",
                  "void loadDecoder1(const char* modelPath, int n_layer, int n_state, int n_head, int n_vocab);", "void closeDecoder1();",
                  "void rearrange_mkv(int* indices, int text_offset);",
                  "void decoder1Predict( float* x, float* qk_mask, int text_offset, float* out_x);"]:
        assert proto in flat, proto


def test_missing_library_fails_loudly(monkeypatch):
    from whisper_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libwhisper_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_export_fragment_layout_roundtrip():
    from whisper_b200.export import to_frag
    w = torch.arange(40 * 64, dtype=torch.float32).reshape(40, 64)
    f = to_frag(w)
    assert f.shape == (3, 2, 2, 32, 8)
    for n, k in ((0, 0), (7, 31), (8, 8), (15, 63), (39, 40)):
        nt, r = divmod(n, 16); kc, kk = divmod(k, 32)
        assert f[nt, kc, r // 8, (r % 8) * 4 + kk // 8, kk % 8] == w[n, k]
    assert (f[2, :, 1] == 0).all()                  # rows 40..47 are padding


def test_export_writes_readable_containers(tmp_path):
    import struct
    from oracle import model as om
    from whisper_b200 import export
    dims = om.DIMS["nano"]
    export.export_model(om.init_weights(dims, 0), dims, str(tmp_path))
    for name in ("Encoder.b2w", "CrossKV.b2w", "Decoder.b2w"):
        raw = open(tmp_path / name, "rb").read()
        magic, n, off, nbytes = struct.unpack("<4sIQQ", raw[:24])
        assert magic == b"B2W1" and n > 0 and off % 256 == 0 and off + nbytes == len(raw)


def test_segment_slicing_follows_timestamp_pairs():
    from whisper_b200.decoding import DecodingResult
    from whisper_b200.transcribe import _segments_from_tokens
    tb, eot = 50364, 50257
    toks = [tb, 11, 12, tb + 100, tb + 100, 13, tb + 250, tb + 250]
    for keep_tail in (False, True):                                                  # (the tail here is a lone opening timestamp: no segment)
        segs = _segments_from_tokens(toks, DecodingResult(tokens=toks), 30.0, 30.0, 3000, tb, keep_tail)
        assert [(round(s["start"], 2), round(s["end"], 2)) for s in segs] == [(30.0, 32.0), (32.0, 35.0)]
    # fixed windows keep the text after the last pair (the reference would re-decode it from the next seek)
    toks2 = toks + [14, 15]
    segs = _segments_from_tokens(toks2, DecodingResult(tokens=toks2), 30.0, 30.0, 3000, tb, True)
    assert [(round(s["start"], 2), round(s["end"], 2)) for s in segs] == [(30.0, 32.0), (32.0, 35.0), (35.0, 60.0)]
    assert segs[-1]["tokens"] == [tb + 250, 14, 15]
    assert len(_segments_from_tokens(toks2, DecodingResult(tokens=toks2), 30.0, 30.0, 3000, tb, False)) == 2
    segs = _segments_from_tokens([11, 12, tb + 50], DecodingResult(), 0.0, 30.0, 0, tb, True)
    assert len(segs) == 1 and segs[0]["end"] == 1.0


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisper_b200.transcribe import gather_sharded
    seeks = list(range(0, 7 * 3000, 3000))[rank::world]
    part = {"segments": [{"seek": s, "start": s / 100.0, "end": s / 100.0 + 30, "tokens": [rank]} for s in seeks],
            "windows": len(seeks), "seeks": seeks}
    merged = gather_sharded(part, world)
    q.put((rank, [s["seek"] for s in merged["segments"]], merged["windows"]))
    dist.destroy_process_group()


def test_window_sharding_merge_world_size_2():
    """N > 1 path on CPU: ranks take windows r::W, results are merged by window start (gloo)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    for rank, seeks, windows in got:
        assert seeks == list(range(0, 7 * 3000, 3000)) and windows == 7
