"""median filter and DTW on the device: bit-exact against the oracle (which is pinned to the reference's
own known-answer tests, tests/test_timing.py of the reference) and against the committed goldens."""
import numpy as np
import pytest
import torch

from oracle import timing as ot
from tests._util import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(10,), (1, 15), (4, 5, 345), (3, 6, 40, 128), (2, 3, 1500)])
@pytest.mark.parametrize("width", [3, 5, 7, 13])
def test_median_filter(shape, width):
    from whisper_b200.timing import median_filter
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(7))
    assert torch.equal(median_filter(x, width), ot.median_filter(x, width))


def test_median_filter_golden_and_short_rows():
    from whisper_b200.timing import median_filter
    x = torch.randn(2, 9, 200, generator=torch.Generator().manual_seed(8))
    assert np.array_equal(median_filter(x, 7).numpy(), golden("timing")["median7"])
    short = torch.randn(4, 3)
    assert torch.equal(median_filter(short, 7), short)        # rows no longer than the padding are returned unchanged


@pytest.mark.parametrize("n,m", [(10, 20), (32, 16), (123, 1500), (234, 189)])
def test_dtw_known_answer(n, m):
    """tests/test_timing.py:22-52 of the reference: carve a monotone path into a random matrix."""
    from whisper_b200.timing import dtw
    rng = np.random.RandomState(42)
    steps = np.concatenate([np.zeros(n - 1), np.ones(m - 1)]); rng.shuffle(steps)
    x = rng.random((n, m)).astype(np.float32)
    i = j = 0; x[0, 0] -= 1
    for s in steps:
        if s == 0: i += 1
        else: j += 1
        x[i, j] -= 1
    ti, tj = dtw(x)
    oi, oj = ot.dtw(x)
    assert np.array_equal(ti, oi) and np.array_equal(tj, oj)
    assert ti[0] == 0 and tj[0] == 0 and ti[-1] == n - 1 and tj[-1] == m - 1


def test_dtw_goldens_ties_and_full_size():
    from whisper_b200.timing import dtw
    g = golden("timing")
    rng = np.random.RandomState(42)
    for n, m in ((10, 20), (32, 16), (123, 1500), (234, 189)):
        x = rng.randn(n, m).astype(np.float32)
        ti, tj = dtw(x)
        assert np.array_equal(ti, g[f"dtw_{n}_{m}_i"]) and np.array_equal(tj, g[f"dtw_{n}_{m}_j"])
    x = rng.randint(0, 3, size=(40, 60)).astype(np.float32)      # ties: strict-< fall-through rule (timing.py:95-100)
    ti, tj = dtw(x)
    assert np.array_equal(ti, g["dtw_ties_i"]) and np.array_equal(tj, g["dtw_ties_j"])
    for n, m in ((444, 1500), (1, 1), (1, 37), (29, 1), (700, 1500)):   # largest alignment problem; degenerate; global-memory trace
        x = np.random.RandomState(n).randn(n, m).astype(np.float32)
        ti, tj = dtw(x)
        oi, oj = ot.dtw(x)
        assert np.array_equal(ti, oi) and np.array_equal(tj, oj), (n, m)
