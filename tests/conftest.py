import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O

    return O.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O

    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return O.Ref()


def ir_frame(h, w, seed, n_bad_frac=1e-3):
    """IR-like frame of SURVEY.md 8d (same recipe as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    b = 8000 + 2000 * np.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
    f = np.clip(b + rng.normal(0, 3, (h, w)), 0, 16383).astype(np.uint16)
    nb = min(h * w, max(2, round(n_bad_frac * h * w)))
    idx = rng.choice(h * w, nb, replace=False)
    f.flat[idx[: nb // 2]] = 0
    f.flat[idx[nb // 2:]] = 16000
    return f


def ir_movie(t, h, w, seed=1234, drift=0.05):
    """Small IR-like movie: background + drifting hot spot + noise, stuck pixels fixed."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    b = 8000 + 2000 * np.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
    nb = max(2, round(1e-3 * h * w))
    idx = np.random.default_rng(4321).choice(h * w, nb, replace=False)
    mov = np.empty((t, h, w), dtype=np.uint16)
    for i in range(t):
        cx, cy = w * 0.3 + drift * i, h * 0.6 - drift * i
        g = 1500 * np.exp(-(((x - cx) / 6.0) ** 2) - ((y - cy) / 6.0) ** 2)
        f = np.clip(b + g + rng.normal(0, 3, (h, w)), 0, 16383).astype(np.uint16)
        f.flat[idx[: nb // 2]] = 0
        f.flat[idx[nb // 2:]] = 16000
        mov[i] = f
    return mov
