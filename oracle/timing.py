"""Oracle: word-timestamp numerics (median filter, DTW, find_alignment pre-processing).

TEST INFRASTRUCTURE - see oracle/__init__.py.  The C bodies live in oracle/timing_c.c and are
compiled by oracle/build.py; small pure-Python versions are kept for cross-checking the C.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_timing.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "timing_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_dtw.restype = ctypes.c_int
    return _lib


def dtw(x: np.ndarray):
    """whisper/timing.py:141-160 dtw() CPU branch: dtw_cpu(x.double()).  Returns (text_idx, time_idx)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, m = x.shape
    oi = np.empty(n + m, dtype=np.int32); oj = np.empty(n + m, dtype=np.int32)
    k = _c().oracle_dtw(x.ctypes.data_as(ctypes.c_void_p), n, m,
                        oi.ctypes.data_as(ctypes.c_void_p), oj.ctypes.data_as(ctypes.c_void_p))
    return oi[:k].astype(np.int64), oj[:k].astype(np.int64)


def dtw_py(x: np.ndarray):
    """Pure-Python twin of oracle_dtw for small cases (timing.py:57-105)."""
    n, m = x.shape
    cost = np.full((n + 1, m + 1), np.inf, dtype=np.float32)
    trace = -np.ones((n + 1, m + 1), dtype=np.int8)
    cost[0, 0] = 0
    for j in range(1, m + 1):
        for i in range(1, n + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = np.float32(np.float64(x[i - 1, j - 1]) + np.float64(c))
            trace[i, j] = t
    trace[0, :] = 2
    trace[:, 0] = 1
    i, j, path = n, m, []
    while i > 0 or j > 0:
        path.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i, j = i - 1, j - 1
        elif t == 1:
            i -= 1
        else:
            j -= 1
    p = np.array(path[::-1]).T
    return p[0], p[1]


def median_filter(x: torch.Tensor, width: int) -> torch.Tensor:
    """whisper/timing.py:19-54 along the last dim."""
    x = x.float().contiguous()
    length = x.shape[-1]
    rows = x.numel() // max(length, 1)
    y = torch.empty_like(x)
    _c().oracle_median_filter(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()),
                              ctypes.c_long(rows), ctypes.c_int(length), ctypes.c_int(width))
    return y


def alignment_matrix(chw: torch.Tensor, num_frames: int, n_skip: int, medfilt_width: int = 7,
                     qk_scale: float = 1.0) -> torch.Tensor:
    """whisper/timing.py:194-204: raw QK (heads, tokens, 1500) -> matrix fed (negated) to dtw.
    n_skip = len(tokenizer.sot_sequence); last row (eot) dropped."""
    w = chw[:, :, : num_frames // 2].float()
    w = (w * qk_scale).softmax(dim=-1)
    std, mean = torch.std_mean(w, dim=-2, keepdim=True, unbiased=False)
    w = median_filter((w - mean) / std, medfilt_width)
    return w.mean(dim=0)[n_skip:-1]
