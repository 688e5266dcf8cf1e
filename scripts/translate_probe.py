#!/usr/bin/env python
"""scripts/translate_probe.py -- the uint16 translate kernel alone (a-5 / a-6): ms per launch and achieved GB/s on a
movie that is resident in HBM, per-frame shifts as in bench.py.  Evidence for profiles/ and the command ncu wraps.
usage: translate_probe.py [--w 640 --h 512 --frames 2000 --reps 5 --motion --shift random|fixed|zero|half]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from librir_b200 import _lib, signal_processing as sp, video_io as vio

    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=640)
    ap.add_argument("--h", type=int, default=512)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--motion", action="store_true")
    ap.add_argument("--shift", default="random")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(777)
    n, h, w = args.frames, args.h, args.w
    mov = torch.randint(7000, 9000, (n, h, w), generator=g, device=dev, dtype=torch.int32).to(torch.int16).view(torch.uint16)
    if args.shift == "random":
        dx = torch.rand(n, generator=g, device=dev) * 6 - 3
        dy = torch.rand(n, generator=g, device=dev) * 6 - 3
    elif args.shift == "fixed":
        dx = torch.full((n,), 1.3, device=dev)
        dy = torch.full((n,), -2.7, device=dev)
    elif args.shift == "half":
        dx = torch.full((n,), 0.5, device=dev)
        dy = torch.full((n,), -1.5, device=dev)
    else:
        dx = torch.zeros(n, device=dev)
        dy = torch.zeros(n, device=dev)
    out = torch.empty_like(mov)
    mx, my = (-dx).double().cpu().numpy(), (-dy).double().cpu().numpy()

    def run():
        if args.motion:
            vio.remove_motion(mov, mx, my, meta_rows=0, out=out)
        else:
            sp.translate_batch(mov, dx, dy, "nearest", 0, out=out)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    gbs = 4.0 * n * h * w / (ms * 1e-3) / 1e9
    peak = 6534.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"kernel": "translate_u16" + ("_motion" if args.motion else ""), "frame": [w, h], "frames": n, "shift": args.shift,
                      "ms": ms, "gbs": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8TBs": gbs / 8000.0,
                      "rows_kernel": os.environ.get("RIRB_TRANSLATE_ROWS", "1")}))


if __name__ == "__main__":
    main()
