import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_library():
    """Makes sure librcn_cuda.so and the oracle exist (built in-tree by __graft_entry__.build())."""
    import __graft_entry__ as g
    from mercer_research_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        g.build()
    return _lib.load()


def assert_close(x, y, rtol=1e-9, what=""):
    """Parity bar for f64 tensors (north_star: 1e-9 relative because the reference computes in f64).
    Elementwise |x-y| <= rtol*|y| + rtol*1e-3*max|y| (the second term only guards exact zeros / denormals)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    assert x.shape == y.shape, (what, x.shape, y.shape)
    if y.size == 0:
        return
    scale = float(np.max(np.abs(y)))
    err = np.abs(x - y)
    bound = rtol * np.abs(y) + rtol * 1e-3 * scale
    bad = err > bound
    if bad.any() or not np.isfinite(x).all():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {int(bad.sum())}/{y.size} elements off; worst at {i}: got {x[i]!r} want {y[i]!r} "
                             f"(err {err[i]:.3e}, scale {scale:.3e}, max rel-to-scale {float(err.max()) / max(scale, 1e-300):.3e})")
