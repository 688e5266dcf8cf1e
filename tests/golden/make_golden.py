"""Generate tests/golden/ref_golden.npz from the REFERENCE ITSELF (oracle/_ref, i.e. the
reference's own sources compiled by oracle/build_ref.sh).  Run in the authoring container,
where /root/reference exists:

    python tests/golden/make_golden.py

The file holds inputs AND the reference's outputs, so the tests that read it depend on
neither /root/reference nor on numpy's RNG stream.  Cases are kept small (the whole file is
a few hundred KB) -- full-size parity runs live against oracle/_ref or the C port.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SHIFTS = [(1.3, -2.7), (0.0, 0.0), (-0.25, 0.5), (3.0, 2.0), (-17.5, 10.25), (0.99999997, 1e-30), (19.999, 15.999),
          (-1.25, 1.25)]
STRATEGIES = ["", "background", "wrap", "nearest"]
DTYPES = ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64"]


def ir_frame(h, w, seed, n_bad_frac=1e-3):
    """IR-like frame of SURVEY.md 8d: smooth background + noise + stuck pixels, < 2**14."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    b = 8000 + 2000 * np.exp(-(((x - w / 2) / (0.23 * w)) ** 2) - ((y - h / 2) / (0.23 * h)) ** 2)
    f = np.clip(b + rng.normal(0, 3, (h, w)), 0, 16383).astype(np.uint16)
    nb = max(2, round(n_bad_frac * h * w))
    idx = rng.choice(h * w, nb, replace=False)
    f.flat[idx[: nb // 2]] = 0
    f.flat[idx[nb // 2:]] = 16000
    return f


def typed_image(dt, h, w, rng):
    dt = np.dtype(dt)
    if dt == np.bool_:
        return rng.random((h, w)) > 0.5
    if dt.kind == "f":
        return (rng.random((h, w)) * 1000 - 300).astype(dt)
    info = np.iinfo(dt)
    lo, hi = max(info.min, -(2 ** 40)), min(info.max, 2 ** 40)
    return rng.integers(lo, hi, (h, w), dtype=np.int64).astype(dt)


def main():
    ref = O.Ref()
    rng = np.random.default_rng(20261018)
    out = {}
    # --- translate: every dtype x strategy x shift on a 16x20 image ------------------------
    for dt in DTYPES:
        img = typed_image(dt, 16, 20, rng)
        out[f"tr_{dt}_in"] = img
        for si, st in enumerate(STRATEGIES):
            for k, (dx, dy) in enumerate(SHIFTS):
                out[f"tr_{dt}_{si}_{k}"] = ref.translate(img, dx, dy, st, background=1)
    out["tr_shifts"] = np.array(SHIFTS, dtype=np.float64)
    # a realistic uint16 IR frame, the C1 shift, all strategies
    f = ir_frame(96, 128, 11)
    out["tr_ir_in"] = f
    for si, st in enumerate(STRATEGIES):
        out[f"tr_ir_{si}"] = ref.translate(f, 1.3, -2.7, st, background=7)
    # --- gaussian ------------------------------------------------------------------------
    g = ir_frame(40, 56, 12).astype(np.float32)
    out["ga_in"] = g
    sig = [0.3, 0.5, 1.0, 1.7, 2.0, 4.2]
    out["ga_sigmas"] = np.array(sig, dtype=np.float32)
    for k, s in enumerate(sig):
        out[f"ga_{k}"] = ref.gaussian_filter(g, s)
    # --- bad pixels: detection list, create/correct through the C facade ------------------
    for k, (h, w) in enumerate([(64, 80), (33, 47), (5, 5), (7, 4), (3, 9)]):
        first = ir_frame(h, w, 100 + k)
        other = ir_frame(h, w, 200 + k)
        out[f"bp_{k}_first"] = first
        out[f"bp_{k}_other"] = other
        out[f"bp_{k}_xy"] = ref.bad_pixels_list(first)
        hd = ref.bad_pixels_create(first)
        out[f"bp_{k}_first_out"] = ref.bad_pixels_correct(hd, first)
        out[f"bp_{k}_other_out"] = ref.bad_pixels_correct(hd, other)
        ref.bad_pixels_destroy(hd)
        out[f"bp_{k}_motion"] = ref.loader_remove_motion(first, 1.3, -2.7)
    # --- find_median_pixel ---------------------------------------------------------------
    f = ir_frame(48, 64, 13)
    m = (rng.random(f.shape) > 0.4).astype(np.uint8)
    pcs = [0.0, 0.1, 0.37, 0.5, 0.9, 1.0]
    out["mp_in"] = f
    out["mp_mask"] = m
    out["mp_percents"] = np.array(pcs, dtype=np.float32)
    out["mp_out"] = np.array([ref.find_median_pixel(f, p) for p in pcs], dtype=np.int32)
    out["mp_out_mask"] = np.array([ref.find_median_pixel(f, p, m) for p in pcs], dtype=np.int32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
