#!/usr/bin/env python
"""scripts/percall_bench.py -- latency of the reference-shaped, one-frame-per-call entries (host numpy
in, host numpy out: the seam of INTEGRATION.md section 1), with the compiled reference (oracle/_ref)
timed beside them on the same frame when it is there.  Evidence for profiles/, not a bench.py line.

    python scripts/percall_bench.py [--reps 300] > gpurun_out/percall.jsonl
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, reps, warm=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6, min(ts) * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=300)
    args = ap.parse_args()
    from librir_b200 import signal_processing as sp
    from tests.conftest import ir_movie

    ref = None
    try:
        from oracle import oracle as orc  # checker only: times the compiled reference beside the product
        ref = orc.Ref() if orc.have_ref() else None
    except Exception as e:  # noqa: BLE001
        print(f"no compiled reference: {e}", file=sys.stderr)

    for (h, w) in [(512, 640), (1024, 1280)]:
        mov = ir_movie(4, h, w)
        img = mov[1]
        bp = sp.BadPixels(mov[0])
        cases = {
            "bad_pixels_correct": lambda: bp.correct(img),
            "gaussian_filter_s1": lambda: sp.gaussian_filter(img, 1.0),
            "translate_nearest": lambda: sp.translate(img, 1.3, -2.6, "nearest"),
            "find_median_pixel": lambda: sp.find_median_pixel(img, 0.5),
        }
        refcases = {}
        if ref is not None:
            rh = ref.bad_pixels_create(mov[0])
            refcases = {
                "bad_pixels_correct": lambda: ref.bad_pixels_correct(rh, img),
                "gaussian_filter_s1": lambda: ref.gaussian_filter(img, 1.0),
                "translate_nearest": lambda: ref.translate(img, 1.3, -2.6, "nearest"),
                "find_median_pixel": lambda: ref.find_median_pixel(img, 0.5),
            }
        for name, fn in cases.items():
            med, best = timeit(fn, args.reps)
            row = {"frame": [w, h], "call": name, "us_median": round(med, 1), "us_min": round(best, 1)}
            if name in refcases:
                rmed, rbest = timeit(refcases[name], max(20, args.reps // 10), warm=3)
                row.update({"reference_us_median": round(rmed, 1), "speedup": round(rmed / med, 2)})
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
