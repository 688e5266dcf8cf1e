"""tests/golden/bp_extreme_golden.npz: the COMPILED REFERENCE's bad-pixel list, clamp behaviour and corrected frames on
tests/bp_extreme_cases.py (oracle/_ref).  Run in the authoring container: python tests/golden/make_bp_extreme_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import bp_extreme_cases as bc  # noqa: E402


def main():
    ref = O.Ref()
    out = {}
    for k, first in enumerate(bc.frames()):
        out[f"xy_{k}"] = np.asarray(ref.bad_pixels_list(first), dtype=np.int32).reshape(-1, 2)
        h = ref.bad_pixels_create(first)
        out[f"first_out_{k}"] = ref.bad_pixels_correct(h, first)
        out[f"other_out_{k}"] = ref.bad_pixels_correct(h, bc.second_frame(first, k))
        ref.bad_pixels_destroy(h)
        print(k, first.shape, "bad pixels:", len(out[f"xy_{k}"]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bp_extreme_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
