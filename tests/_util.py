"""Shared helpers for the parity tests: seeded oracle weights exported to .b2w, error metrics."""
import os
import tempfile

import numpy as np
import torch

from oracle import model as om

_EXPORTED = {}


def rel(a, b) -> float:
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def exported(name: str, seed: int = 0, logit_scale: float = 1.0):
    """(dims, ckpt, folder): oracle weights for `name`, exported once per process."""
    key = (name, seed, logit_scale)
    if key not in _EXPORTED:
        from whisper_b200 import export
        dims = om.DIMS[name]
        ckpt = om.init_weights(dims, seed, logit_scale)
        folder = os.path.join(tempfile.gettempdir(), f"b200_weights_{os.getpid()}", f"{name}_{seed}_{logit_scale}")
        export.export_model(ckpt, dims, folder, fused=False)
        _EXPORTED[key] = (dims, ckpt, folder)
    return _EXPORTED[key]


def golden(name: str):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"ref_{name}.npz")
    return np.load(path)


def close_library():
    """The library holds ONE model per process (like the reference's plugin): a test that builds its own model while a
    module-scoped fixture still has another one loaded would silently run on the fixture's weights (loads are idempotent)."""
    from whisper_b200 import _lib
    lib = _lib.load()
    lib.closeDecoder1(); lib.closeDecoder256(); lib.closeCrossKV(); lib.closeEncoder()
    _lib.check_errors("close")
