"""Cases for the rows of the path that live in the reference's video_io library: the lossless writer's byte-plane split
and key-frame rule (a-7), the lossy pre-conditioner in both of its doors (f-2) and the reader's post-decode chain
(a-3, a-6, f-1).  Shared by tests/golden/make_vio_golden.py -- which has the COMPILED REFERENCE (oracle/_ref/libs/
libvideo_io.so, the reference's own sources over oracle/libav_stub.c) answer them -- and by the tests.  Inputs are built
from a hash (no RNG stream), so the golden file only has to hold the reference's answers."""
import zlib

import numpy as np

from tests.ecc_cases import hash_noise


def movie(t, h, w, seed, drift=0.05, jump_at=None, jump=40, ramp_every=0, n_bad_frac=2e-3):
    """IR-like movie: smooth background + drifting hot spot + noise + stuck pixels, values < 2**14.
    ``jump_at`` / ``ramp_every``: a step and a slow ramp of the whole scene (what makes the lossy bounds move)."""
    y, x = np.mgrid[0:h, 0:w]
    b = 8000 + 2000 * np.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
    nb = max(2, round(n_bad_frac * h * w))
    order = np.argsort(hash_noise((1, h * w), seed * 7 + 1)[0], kind="stable")[:nb]
    mov = np.empty((t, h, w), dtype=np.uint16)
    for i in range(t):
        cx, cy = w * 0.3 + drift * i, h * 0.6 - drift * i
        g = 1500 * np.exp(-(((x - cx) / 6.0) ** 2) - ((y - cy) / 6.0) ** 2)
        f = b + g + 10.0 * hash_noise((h, w), seed * 1000 + i)  # uniform in [-5, 5)
        if ramp_every:
            f = f + i // ramp_every
        if jump_at is not None and i >= jump_at:
            f = f + jump
        f = np.clip(np.rint(f), 0, 16383).astype(np.uint16)
        f.flat[order[: nb // 2]] = 0
        f.flat[order[nb // 2:]] = 16000
        mov[i] = f
    return mov


def shifts(t, seed, amp=3.0):
    """Per-frame registration shifts, exactly representable in float32 (the .regfile is parsed as float)."""
    sx = (2 * amp * hash_noise((1, t), seed * 13 + 5)[0]).astype(np.float32)
    sy = (2 * amp * hash_noise((1, t), seed * 13 + 6)[0]).astype(np.float32)
    return sx, sy


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# ---- a-7: writer.  (name, frames, h, w, gop, codec) ------------------------------------------------------------------
SPLIT_CASES = [
    ("s444_a", 27, 35, 48, 5, "h264"),
    ("s444_b", 12, 17, 40, 50, "h264"),    # width not a multiple of 32, one key frame
    ("s444_c", 130, 10, 33, 50, "h264"),   # odd width, three GOPs
    ("s444_d", 7, 64, 64, 1, "h264"),      # every frame a key frame
    ("s420_a", 27, 35, 48, 5, "h265"),     # kvazaar layout: frame padded to (48, 80), no key-frame rule
    ("s420_b", 9, 19, 33, 50, "h265"),     # padded to (40, 48)
]

# ---- f-2: lossy pre-conditioner.  The saver's string parameters (h264.cpp:1709-1781) ---------------------------------
LOSSY_CONFIGS = [
    ("default", dict()),
    ("no_average", dict(runningAverage=0)),
    ("sub_min", dict(subtractMin=1)),
    ("bp_sub_min", dict(removeBadPixels=1, subtractMin=1)),
    ("tight", dict(lowValueError=10, highValueError=4, stdFactor=2.0, runningAverage=8)),
]
LOSSY_SHAPE = (220, 40, 48, 37)  # frames, h, w, stop_lossy_height
LOSSY_DOORS = ("add_image_lossy", "add_loss")  # addImageLossyNoCamera (h264.cpp:2253-2424) / addLoss (:2426-2607)
LOSSY_FULL_EVERY = 20  # frames stored in full; every frame is pinned by its CRC


def lossy_movie():
    t, h, w, _ = LOSSY_SHAPE
    return movie(t, h, w, seed=9, jump_at=t // 2, ramp_every=7)


def lossy_params(cfg):
    """kwargs of oracle.Port.lossy_open / librir_b200.video_io.LossyPreconditioner for a saver configuration."""
    return dict(low_error=cfg.get("lowValueError", 6), high_error=cfg.get("highValueError", 2), std_factor=cfg.get("stdFactor", 5.0),
                running_average=cfg.get("runningAverage", 32), subtract_min=bool(cfg.get("subtractMin", 0)),
                bp_enabled=bool(cfg.get("removeBadPixels", 0)))


# ---- a-3 / a-6 / f-1: reader.  (name, frames, h, w, codec, MIN_T, MIN_T_HEIGHT) ---------------------------------------
LOADER_CASES = [
    ("r_plain", 40, 35, 48, "h264", 0, 0),
    ("r_min_t", 40, 35, 48, "h264", 300, 0),        # MIN_T_HEIGHT absent -> height - 3 (IRFileLoader.cpp:918-921)
    ("r_420_min_t", 40, 35, 48, "h265", 120, 20),
    ("r_wrap", 25, 67, 80, "h264", 60000, 64),      # unsigned short += int wraps
    ("r_small", 30, 19, 24, "h265", 0, 0),
]
LOADER_MODES = [(0, 0), (0, 1), (1, 0), (1, 1)]  # (bad pixels, motion correction)


def loader_movie(case):
    name, t, h, w, codec, min_t, min_th = case
    return movie(t, h, w, seed=t + h)
