"""Frames whose spread around the median exceeds 46,340 counts -- saturated or dead pixels over an ordinary background,
full-range noise -- where the reference's global spread is an overflowing int product (Filters.h:145-156).  Deterministic
(no RNG stream): shared by tests/golden/make_bp_extreme_golden.py, which has the compiled reference answer them, and the tests."""
import numpy as np

from tests.ecc_cases import hash_noise


def frames():
    out = []
    for k, (h, w) in enumerate([(40, 64), (33, 17), (64, 136), (5, 3), (90, 40)]):
        n = hash_noise((h, w), 100 + k) + 0.5  # [0, 1)
        m = hash_noise((h, w), 200 + k) + 0.5
        if k == 0:      # ordinary background, 5 % saturated pixels
            f = 8000 + 600 * n
            f[m < 0.05] = 65535
        elif k == 1:    # full-range noise
            f = 65535 * n
        elif k == 2:    # bright scene, 10 % dead pixels
            f = 52000 + 13000 * n
            f[m < 0.10] = 0
        elif k == 3:    # a handful of pixels, half of them saturated
            f = 3000 * n
            f[m < 0.5] = 65535
        else:           # a third of the pixels near saturation: the wrapped sum decides the thresholds
            f = 2500 * n
            f[m < 0.33] = 61000 + 4000 * n[m < 0.33]
        out.append(np.clip(np.rint(f), 0, 65535).astype(np.uint16))
    return out


def second_frame(first, k):
    """Another frame of the same 'movie': same structure, different noise."""
    n = hash_noise(first.shape, 300 + k)
    return np.clip(first.astype(np.int64) + np.rint(400 * n).astype(np.int64), 0, 65535).astype(np.uint16)
