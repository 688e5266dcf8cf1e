"""Reader merge pass (loader_merge_kernel) with and without the bad-pixel medians and the min_T offset, 2,000 frames of
640x512: where the distance to the plain plane merge (decode_movie, 0.99 of peak) comes from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from bench import synth_movie_torch, W, H, hbm_peak
from librir_b200 import video_io as vio

n = 2000
dev = torch.device("cuda", 0)
frames = synth_movie_torch(n, 0, dev)
lo = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
hi = torch.empty_like(lo)
out = torch.empty_like(frames)
vio.precode_movie(frames, 50, False, 0, out=(lo, hi))
lbp = vio.LoaderBadPixels(frames[0].cpu().view(torch.int16).numpy().view(np.uint16))
peak, _ = hbm_peak()
for name, fn in (("merge only", lambda: vio.read_movie(lo, hi, None, 0, 0, None, None, out=out)),
                 ("merge + min_T", lambda: vio.read_movie(lo, hi, None, 273, H - 3, None, None, out=out)),
                 ("merge + medians", lambda: vio.read_movie(lo, hi, lbp, 0, 0, None, None, out=out)),
                 ("merge + min_T + medians", lambda: vio.read_movie(lo, hi, lbp, 273, H - 3, None, None, out=out)),
                 ("decode_movie (plain merge)", lambda: vio.decode_movie(lo, hi, 50, False, 0, out=out))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:28s} {ms:.3f} ms  {4 * W * H * n / ms / 1e6 / peak:.2f} of peak")
