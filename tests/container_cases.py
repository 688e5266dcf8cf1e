"""Deterministic inputs of the file-format tests (no RNG stream involved): shared by
tests/golden/make_container_golden.py, which has the compiled reference write the golden files, and
tests/test_container.py, which checks the product and the oracle port against them."""
import numpy as np


def golden_movie():
    """6 frames of 32 x 24 uint16: a smooth ramp, a moving blob and a few isolated spikes."""
    t = np.arange(6).reshape(6, 1, 1)
    y = np.arange(24).reshape(1, 24, 1)
    x = np.arange(32).reshape(1, 1, 32)
    f = 8000 + 13 * x + 7 * y + (900 * np.exp(-(((x - 10 - 2 * t) / 4.0) ** 2) - ((y - 12) / 3.0) ** 2)).astype(np.int64)
    f = f + ((x * 31 + y * 17 + t * 5) % 11)
    f[:, 3, 5] = 0
    f[:, 20, 30] = 16383
    return f.astype(np.uint16)


def golden_times():
    return (np.arange(6, dtype=np.int64) * 20_000_000 + 1_700_000_000_000_000_000)


def golden_attrs():
    """(global, per-frame, payload) -- bytes keys and values; one value long enough to be zstd-compressed in the
    trailer, one long but incompressible (stays raw), one empty."""
    lcg = np.empty(1500, dtype=np.uint8)
    s = 12345
    for i in range(1500):
        s = (1103515245 * s + 12345) % (1 << 31)
        lcg[i] = (s >> 16) & 0xFF
    g = {
        b"Device": b"synthetic",
        b"MIN_T": b"273",
        b"empty": b"",
        b"long_text": b"librir attribute " * 120,
        b"long_noise": lcg.tobytes(),
        "unitµ".encode("utf8"): "µm".encode("utf8"),
    }
    frames = [{b"BackgroundError": str(6 + i).encode(), b"ForegroundError": b"2"} if i % 2 == 0 else {} for i in range(6)]
    payload = bytes(range(256)) * 2
    return g, frames, payload
