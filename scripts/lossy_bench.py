#!/usr/bin/env python
"""scripts/lossy_bench.py -- frames/s of the lossy pre-conditioner on a 640x512 movie:
the GPU path (device-resident frames, one rirb_lossy_add_images call) next to the restated reference on one
host core.  Evidence for profiles/, not a bench.py line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    from bench import synth_movie_torch, W, H
    from librir_b200 import video_io as vio
    from oracle import oracle as O

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    dev = torch.device("cuda", 0)
    frames = synth_movie_torch(n, 0, dev)
    out = torch.empty_like(frames)
    rows = []
    for cfg in (dict(), dict(runningAverage=0), dict(removeBadPixels=True, subtractMin=True)):
        times = []
        for attempt in range(7):  # the first pass of a process pays one-time allocations: dropped; median of the other six
            pre = vio.LossyPreconditioner(W, H, H - 3, **cfg)
            pre.add_images(frames[:50], out=out[:50])  # the first-image branch and the window filling up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pre.add_images(frames[50:], out=out[50:])
            e1.record()
            torch.cuda.synchronize()
            if attempt:
                times.append(e0.elapsed_time(e1))
        times.sort()
        ms = 0.5 * (times[2] + times[3])
        frozen = float((out[50:].view(torch.int16) != frames[50:].view(torch.int16)).float().mean())
        # CPU: the restated reference, one core
        port = O.Port()
        mov = frames[:120].cpu().view(torch.int16).numpy().view(np.uint16)
        st = port.lossy_open(W, H, H - 3, cfg.get("lowValueError", 6), cfg.get("highValueError", 2), 5.0, cfg.get("runningAverage", 32),
                             cfg.get("subtractMin", False), cfg.get("removeBadPixels", False))
        for t in range(20):
            port.lossy_add(st, mov[t])
        t0 = time.perf_counter()
        for t in range(20, 120):
            port.lossy_add(st, mov[t])
        cpu = 100 / (time.perf_counter() - t0)
        port.lossy_close(st)
        rows.append({"config": cfg or "defaults", "frames": n - 50, "gpu_frames_per_s": (n - 50) / (ms * 1e-3), "us_per_frame": 1e3 * ms / (n - 50),
                     "us_per_frame_min_max": [round(1e3 * times[0] / (n - 50), 2), round(1e3 * times[-1] / (n - 50), 2)],
                     "cpu_port_frames_per_s_1core": cpu, "fraction_of_pixels_changed": frozen})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
